"""Generates ``tests/golden/reference_grads.pt``: gradients of the REFERENCE's own ``DrivingModel.forward_loss`` (run here
exactly as ``make_golden.py`` builds it: the reference's driving.py / adaptors.py / internvl2_model.py / llm.py / utils.py
on top of ``transformers.Qwen2ForCausalLM``), obtained with torch autograd.  They pin the *backward* of the oracle, which
in turn is what the CUDA backward is compared with (tests/test_training_gpu.py):

  * every parameter of ``adaptors.driving.*`` and ``wp_encoder.*``                      (reference modules, direct)
  * LoRA A / B of every adapted Qwen2 linear, from the gradient of the merged HF weight the reference run sees:
    W = W0 + s B A  =>  dA = s B^T dW,  dB = s dW A^T                                   (chain rule, exact)
  * d loss / d vit_embeds, the gradient that enters the ViT backward through ``replace_placeholder_tokens``
    (internvl2_model.py:119-131) — the ViT itself has no independent implementation offline

Full tensors would be megabytes; each gradient is stored as (L2 norm, projection on a seeded Gaussian vector, first 4
values).  Inputs and weights are regenerated from seeds by ``simlingo_b200.spec``.

    python tests/golden/make_golden_grads.py
"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec_ = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
G = importlib.util.module_from_spec(spec_)
spec_.loader.exec_module(G)          # sets up the reference imports / stubs; does not run main()

from simlingo_b200.spec import LLM_PREFIX, init_state_dict, synth_labels, tiny_spec  # noqa: E402


def summary(key: str, g: torch.Tensor) -> dict:
    gen = torch.Generator().manual_seed(abs(hash_key(key)) % (2 ** 31))
    r = torch.randn(g.numel(), generator=gen)
    return dict(norm=float(g.norm()), proj=float((g.flatten().double() * r.double()).sum()), head=g.flatten()[:4].clone())


def hash_key(key: str) -> int:
    h = 0
    for ch in key:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    spec = tiny_spec(2, 2, 4096)
    sd = init_state_dict(spec, seed=0)
    model = G.build_reference_model(spec, sd)
    out = {"weights_seed": 0, "cases": []}
    for name, B, seed, pads in [("loss_b2", 2, 5, ()), ("loss_b2_padded", 2, 5, ((1, 9),))]:
        di = G.make_input(spec, B, seed, None, answer_len=16, pad_rows=pads)
        wps, path = synth_labels(spec, B, seed)
        ex = G.DrivingExample(di, G.DrivingLabel(wps, path, di.prompt, torch.zeros(1)), ["x"] * B)
        with torch.no_grad():
            feats = G.O.extract_feature(sd, spec, di.camera_images.flatten(0, 2))
        leaf = feats.clone().requires_grad_(True)
        chat = model.vision_model.image_encoder.model
        chat.extract_feature = lambda px, leaf=leaf: leaf
        model.zero_grad()
        to, _ = model.forward_loss(ex)
        to.loss.backward()
        grads = {"dvit_embeds": summary("dvit_embeds", leaf.grad)}
        for k, p in model.named_parameters():
            if k.startswith("adaptors.driving.") or k.startswith("wp_encoder."):
                grads[k] = summary(k, p.grad)
        hf = model.language_model.model
        for k, p in hf.named_parameters():
            full = LLM_PREFIX + k
            stem, kind = full.rsplit(".", 1)
            if kind != "weight" or f"{stem}.lora_A.default.weight" not in sd:
                continue
            A, Bm = sd[f"{stem}.lora_A.default.weight"], sd[f"{stem}.lora_B.default.weight"]
            grads[f"{stem}.lora_A.default.weight"] = summary(f"{stem}.lora_A.default.weight", spec.lora_scale * Bm.t() @ p.grad)
            grads[f"{stem}.lora_B.default.weight"] = summary(f"{stem}.lora_B.default.weight", spec.lora_scale * p.grad @ A.t())
        out["cases"].append(dict(name=name, B=B, seed=seed, pads=list(pads), loss=to.loss.detach().clone(), grads=grads))
        print(name, float(to.loss), len(grads), "gradients; |dvit| =", grads["dvit_embeds"]["norm"])
    torch.save(out, os.path.join(HERE, "reference_grads.pt"))
    print("written")


if __name__ == "__main__":
    main()
