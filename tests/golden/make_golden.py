"""Generates ``tests/golden/reference_run.pt`` by running the REFERENCE's own code in this container.

What runs unmodified from /root/reference (imported, never copied):
  * ``simlingo_training/models/driving.py``      DrivingModel.__init__/forward/forward_model/forward_loss
  * ``simlingo_training/models/adaptors/adaptors.py``  every adaptor class
  * ``simlingo_training/models/encoder/internvl2_model.py``  LingoInternVLModel.replace_placeholder_tokens
  * ``simlingo_training/models/language_model/llm.py``  LLM.forward / greedy_sample / sample_categorical
  * ``simlingo_training/models/utils.py``  summarise_losses
on top of ``transformers.Qwen2ForCausalLM`` (eager attention, the class the reference calls; LoRA folded into
its weights because ``peft`` is not installed - identical math in eval mode).  The only part that cannot come
from the reference or an installed package is UPSTREAM InternViT/mlp1 (HF-Hub remote code): the oracle's
restatement supplies ``extract_feature``.  ``hydra`` / ``pytorch_lightning`` are absent and are stubbed just
enough for ``driving.py`` to import (they do no arithmetic).

Usage (in the authoring container, where /root/reference exists):
    python tests/golden/make_golden.py
The fixture stores only seeds + outputs; inputs are regenerated from the seeds by ``simlingo_b200.spec``.
"""
import os
import sys
import types
from types import SimpleNamespace

import torch
from torch import nn

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
# 1) the repo's own test helpers first (oracle + spec; neither imports `simlingo_training`) ...
sys.path.insert(0, REPO)
from oracle import model as O  # noqa: E402
from simlingo_b200.spec import (LLM_PREFIX, LMHEAD_SHIFT, init_state_dict, synth_frames, synth_labels,  # noqa: E402
                                synth_placeholders, synth_prompt_ids, tiny_spec)
# 2) ... then make `simlingo_training` resolve to the REFERENCE (a namespace package there, so the repo's
#    regular package of the same name must not be on the path at all)
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
assert "simlingo_training" not in sys.modules
sys.path.insert(0, REF)

hydra = types.ModuleType("hydra")
hydra.utils = types.ModuleType("hydra.utils")
hydra.utils.instantiate = lambda cfg, **kw: cfg          # we pass ready-made sub-modules as "configs"
hydra.utils.get_original_cwd = lambda: os.getcwd()
sys.modules["hydra"], sys.modules["hydra.utils"] = hydra, hydra.utils
pl = types.ModuleType("pytorch_lightning")


class _LM(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


pl.LightningModule = _LM
sys.modules["pytorch_lightning"] = pl

import simlingo_training  # noqa: E402
assert list(simlingo_training.__path__)[0].startswith(REF), simlingo_training.__path__
from simlingo_training.models import driving as R_driving  # noqa: E402
from simlingo_training.models.encoder import internvl2_model as R_ivl  # noqa: E402
from simlingo_training.models.language_model import llm as R_llm  # noqa: E402
from simlingo_training.utils.custom_types import DrivingExample, DrivingInput, DrivingLabel, LanguageLabel  # noqa: E402
from transformers import Qwen2Config, Qwen2ForCausalLM  # noqa: E402



class StubTokenizer:
    """Duck-typed tokenizer facts the model code touches."""

    def __init__(self, spec):
        self.spec = spec
        self.eos_token_id = spec.eos_id
        self.additional_special_tokens_ids = list(range(spec.first_added_id, spec.first_added_id + 8))
        self.added_tokens_encoder = {"<|im_end|>": spec.eos_id}

    def convert_tokens_to_ids(self, tok):
        return {"<IMG_CONTEXT>": self.spec.img_context_id}[tok]

    def batch_decode(self, tokens, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row) for row in tokens]


def build_reference_model(spec, sd):
    cfg = Qwen2Config(vocab_size=spec.vocab, hidden_size=spec.llm_hidden, intermediate_size=spec.llm_mlp,
                      num_hidden_layers=spec.llm_layers, num_attention_heads=spec.llm_heads,
                      num_key_value_heads=spec.llm_kv_heads, rope_theta=spec.rope_theta, rms_norm_eps=spec.rms_eps,
                      max_position_embeddings=32768, tie_word_embeddings=False, attn_implementation="eager")
    hf = Qwen2ForCausalLM(cfg).eval().float()
    hsd = {}
    for k in hf.state_dict():
        full = LLM_PREFIX + k
        if full in sd:
            hsd[k] = sd[full]
            continue
        stem, kind = full.rsplit(".", 1)
        w = sd[f"{stem}.base_layer.{kind}"]
        if kind == "weight":
            w = w + spec.lora_scale * sd[f"{stem}.lora_B.default.weight"] @ sd[f"{stem}.lora_A.default.weight"]
        hsd[k] = w
    hf.load_state_dict(hsd)

    llm = R_llm.LLM.__new__(R_llm.LLM)
    nn.Module.__init__(llm)
    llm.variant = "OpenGVLab/InternVL2-1B"
    llm.model = hf
    llm.model.embed_tokens = llm.model.base_model.embed_tokens   # reference llm.py:91
    llm.hidden_size, llm.vocab_size = spec.llm_hidden, spec.vocab

    class Chat(nn.Module):
        config = SimpleNamespace(output_attentions=False, output_hidden_states=False, use_return_dict=True)

        def extract_feature(self, px):
            return O.extract_feature(sd, spec, px)

    enc = R_ivl.LingoInternVLModel.__new__(R_ivl.LingoInternVLModel)
    nn.Module.__init__(enc)
    enc.model, enc.processor, enc.use_global_img, enc.num_embeddings = Chat(), StubTokenizer(spec), False, spec.vocab

    class Vision(nn.Module):
        def __init__(self):
            super().__init__()
            self.image_encoder = enc

    model = R_driving.DrivingModel(cfg_data_module={}, processor=StubTokenizer(spec), cache_dir=None, vision_model=Vision(),
                                   language_model=llm, speed_wps_mode="2d", predict_route_as_wps=True, lr=3e-5,
                                   weight_decay=0.1, betas=(0.9, 0.999), pct_start=0.05)
    own = {k: v for k, v in sd.items() if k.startswith("adaptors.driving.") or k.startswith("wp_encoder.")}
    missing = model.load_state_dict(own, strict=False)
    assert not [k for k in missing.unexpected_keys], missing
    return model.eval()


def make_input(spec, B, seed, G_list=None, answer_len=0, pad_rows=()):
    ids = synth_prompt_ids(spec, B, seed, answer_len=answer_len)
    if G_list is not None:
        for b, G in enumerate(G_list):
            ids[b, -1] = (spec.eos_id - G * LMHEAD_SHIFT) % spec.vocab
    valid = torch.ones_like(ids, dtype=torch.bool)
    for b, n in pad_rows:
        valid[b, :n] = False
    lm = torch.zeros_like(valid)
    if answer_len:
        lm[:, -answer_len:] = True
    label = LanguageLabel(ids, valid, valid.clone(), synth_placeholders(spec, B, seed), [""] * B, lm)
    z = torch.zeros(B, 1)
    return DrivingInput(synth_frames(spec, B, seed), z, z, z, z, z, label, label)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    spec = tiny_spec(2, 2, 4096)
    sd = init_state_dict(spec, seed=0)
    model = build_reference_model(spec, sd)
    out = {"spec": dict(vit_layers=2, llm_layers=2, vocab=4096), "weights_seed": 0, "cases": []}
    with torch.no_grad():
        # ---- DrivingModel.forward (inference) ----
        for name, B, seed, G_list, pads in [("agent_b1_g1", 1, 11, [1], ()), ("agent_b1_g4", 1, 11, [4], ()),
                                            ("ragged_b3", 3, 21, [2, 3, 2], ()), ("padded_b3", 3, 21, [2, 3, 2], ((1, 6),))]:
            di = make_input(spec, B, seed, G_list, pad_rows=pads)
            sp, rt, lang = model.forward(di)
            out["cases"].append(dict(kind="forward", name=name, B=B, seed=seed, G_list=G_list, pads=list(pads),
                                     speed_wps=sp.clone(), route=rt.clone(), language=list(lang)))
            print(name, lang)
        # ---- DrivingModel.forward_loss (teacher forced, eval mode) ----
        for name, B, seed, pads in [("loss_b2", 2, 5, ()), ("loss_b2_padded", 2, 5, ((1, 9),))]:
            di = make_input(spec, B, seed, None, answer_len=16, pad_rows=pads)
            wps, path = synth_labels(spec, B, seed)
            ex = DrivingExample(di, DrivingLabel(wps, path, di.prompt, torch.zeros(1)), ["x"] * B)
            to, _ = model.forward_loss(ex)
            out["cases"].append(dict(kind="loss", name=name, B=B, seed=seed, pads=list(pads), loss=to.loss.clone(),
                                     averages={k: v.clone() for k, v in to.loss_averages.items()}))
            print(name, float(to.loss), {k: float(v) for k, v in to.loss_averages.items()})
    torch.save(out, os.path.join(REPO, "tests", "golden", "reference_run.pt"))
    print("written")


if __name__ == "__main__":
    main()
