"""Pins the CPU oracle (oracle/model.py) before anything is compared against it.

1. ``reference_run.pt``: outputs of the REFERENCE's own driving.py / adaptors.py / internvl2_model.py / llm.py /
   utils.py executed in the authoring container on transformers' Qwen2 (tests/golden/make_golden.py) - the
   oracle's restatement of that logic must reproduce them (tokens identical, floats to fp32 round-off).
2. ``transformers.Qwen2ForCausalLM`` (the class the reference calls) vs the oracle's Qwen2 restatement.
3. Upstream pixel_shuffle algorithm vs the closed form; LoRA vs merged weights; torch AdamW / OneCycleLR."""
import os

import pytest
import torch

from oracle import model as O
from simlingo_b200.spec import LLM_PREFIX, init_state_dict, tiny_spec
from tests.helpers import make_case_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_run.pt")


@pytest.fixture(scope="module")
def tiny():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    spec = tiny_spec(2, 2, 4096)
    return spec, init_state_dict(spec, seed=0)


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_golden_fixture_describes_tiny_spec(golden):
    assert golden["spec"] == dict(vit_layers=2, llm_layers=2, vocab=4096) and golden["weights_seed"] == 0


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_oracle_forward_matches_reference_run(tiny, golden, idx):
    spec, sd = tiny
    case = [c for c in golden["cases"] if c["kind"] == "forward"][idx]
    inp = make_case_inputs(spec, case["B"], case["seed"], case["G_list"], pad_rows=[tuple(p) for p in case["pads"]])
    with torch.no_grad():
        sp, rt, toks = O.driving_forward(sd, spec, inp["frames"], inp["ids"], inp["valid"], inp["placeholders"],
                                         max_new_tokens=100, eos_token_id=spec.eos_id)
    assert [" ".join(str(int(t)) for t in row) for row in toks] == case["language"]
    assert torch.allclose(sp, case["speed_wps"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(rt, case["route"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("idx", [0, 1])
def test_oracle_loss_matches_reference_run(tiny, golden, idx):
    spec, sd = tiny
    case = [c for c in golden["cases"] if c["kind"] == "loss"][idx]
    inp = make_case_inputs(spec, case["B"], case["seed"], None, answer_len=16, pad_rows=[tuple(p) for p in case["pads"]])
    wps, path = inp["labels"]
    with torch.no_grad():
        loss, avgs, _ = O.forward_loss(sd, spec, inp["frames"], inp["ids"], inp["valid"], inp["loss_masking"],
                                       inp["placeholders"], wps, path)
    assert torch.allclose(loss, case["loss"], rtol=1e-5)
    for k, v in case["averages"].items():
        assert torch.allclose(avgs[k], v, rtol=1e-5), k


def test_llm_matches_hf_qwen2(tiny):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    spec, sd = tiny
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
    x = torch.randn(2, 40, spec.llm_hidden, generator=torch.Generator().manual_seed(0))
    mask = torch.ones(2, 40, dtype=torch.bool)
    mask[1, :7] = False
    with torch.no_grad():
        out = hf(inputs_embeds=x, attention_mask=mask, output_hidden_states=True)
        feats, logits = O.llm_forward(sd, spec, x, mask)
    # hidden_states[-1] is the post-final-norm state (SURVEY 8a note 3)
    assert torch.allclose(out.hidden_states[-1], feats, atol=2e-4, rtol=1e-4)
    assert (out.logits - logits).abs().max() / logits.abs().max() < 1e-5


def test_pixel_shuffle_closed_form_equals_upstream_algorithm():
    x = torch.randn(3, 1024, 64)
    up = O.pixel_shuffle_upstream(x.reshape(3, 32, 32, 64), 0.5).reshape(3, 256, 256)
    assert torch.equal(up, O.pixel_shuffle_closed_form(x, 32))


def test_lora_equals_merged_weight(tiny):
    spec, sd = tiny
    p = f"{LLM_PREFIX}model.layers.0.self_attn.q_proj."
    x = torch.randn(5, spec.llm_hidden)
    merged = sd[p + "base_layer.weight"] + spec.lora_scale * sd[p + "lora_B.default.weight"] @ sd[p + "lora_A.default.weight"]
    ref = torch.nn.functional.linear(x, merged, sd[p + "base_layer.bias"])
    assert torch.allclose(O.lora_linear(sd, p, x, spec.lora_scale), ref, atol=1e-5)


def test_adamw_and_onecycle_match_torch():
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(64))
    opt = torch.optim.AdamW([p], lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999))
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=3e-5, total_steps=50, pct_start=0.05)
    q, m, v = p.detach().clone(), torch.zeros(64), torch.zeros(64)
    for step in range(1, 6):
        g = torch.randn(64)
        lr, beta1 = O.one_cycle(step - 1, 50, 3e-5, 0.05)
        assert abs(lr - opt.param_groups[0]["lr"]) < 1e-12
        assert abs(beta1 - opt.param_groups[0]["betas"][0]) < 1e-12  # OneCycleLR cycles AdamW's beta1 (cycle_momentum)
        p.grad = g.clone()
        opt.step()
        sched.step()
        q, m, v = O.adamw_step(q, g, m, v, step, lr, beta1=beta1)
        assert torch.allclose(q, p.detach(), atol=2e-6, rtol=0)


def _grad_summary(key, g):
    h = 0
    for ch in key:
        h = (h * 131 + ord(ch)) % 1000003
    gen = torch.Generator().manual_seed(h % (2 ** 31))
    r = torch.randn(g.numel(), generator=gen)
    return float(g.norm()), float((g.flatten().double() * r.double()).sum()), g.flatten()[:4]


@pytest.mark.parametrize("idx", [0, 1])
def test_oracle_backward_matches_reference_run(tiny, idx, monkeypatch):
    """autograd through the oracle vs autograd through the reference's own forward_loss (tests/golden/
    make_golden_grads.py): driving heads / queries / waypoint encoder, every LoRA A / B, and the gradient entering the
    ViT.  Each gradient is compared through its norm, a seeded random projection and its first values (fp32, 1e-3)."""
    from simlingo_b200.spec import trainable
    spec, sd0 = tiny
    case = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_grads.pt"), weights_only=False)["cases"][idx]
    sd = {k: v.clone().requires_grad_(trainable(k)) for k, v in sd0.items()}
    inp = make_case_inputs(spec, case["B"], case["seed"], None, answer_len=16, pad_rows=[tuple(p) for p in case["pads"]])
    wps, path = inp["labels"]
    captured = []
    real = O.extract_feature

    def capture(*a, **k):
        out = real(*a, **k)
        out.retain_grad()
        captured.append(out)
        return out

    monkeypatch.setattr(O, "extract_feature", capture)
    loss, _, _ = O.forward_loss(sd, spec, inp["frames"], inp["ids"], inp["valid"], inp["loss_masking"], inp["placeholders"], wps, path)
    loss.backward()
    assert torch.allclose(loss.detach(), case["loss"], rtol=1e-5) and len(captured) == 1
    got = {k: v.grad for k, v in sd.items() if v.grad is not None}
    got["dvit_embeds"] = captured[0].grad
    assert len(case["grads"]) == 45
    for key, ref in case["grads"].items():
        norm, proj, head = _grad_summary(key, got[key])
        assert norm == pytest.approx(ref["norm"], rel=1e-3), key
        assert abs(proj - ref["proj"]) <= 2e-3 * ref["norm"] + 1e-9, key
        assert torch.allclose(head, ref["head"], rtol=1e-3, atol=1e-3 * ref["norm"] / max(1.0, got[key].numel() ** 0.5) + 1e-12), key


def _hf_internvl(spec, sd):
    """transformers' HF-native InternVL vision tower + projector carrying the synthetic weights (mapped key by key);
    returns (vision model, projector, {reference key: HF parameter})"""
    from transformers.models.internvl import modeling_internvl as M
    from transformers.models.internvl.configuration_internvl import InternVLConfig, InternVLVisionConfig
    from simlingo_b200.spec import MLP1_PREFIX, VIT_PREFIX
    vcfg = InternVLVisionConfig(hidden_size=spec.vit_hidden, num_hidden_layers=spec.vit_layers, num_attention_heads=spec.vit_heads,
                                intermediate_size=spec.vit_mlp, attention_bias=True, use_qk_norm=False, norm_type="layer_norm",
                                layer_norm_eps=spec.vit_eps, image_size=(448, 448), patch_size=(14, 14), hidden_act="gelu",
                                use_mean_pooling=True, attn_implementation="eager")
    hf = M.InternVLVisionModel(vcfg).eval().float()
    E, L = VIT_PREFIX + "embeddings.", VIT_PREFIX + "encoder.layers."
    m = {"embeddings.cls_token": E + "class_embedding", "embeddings.position_embeddings": E + "position_embedding",
         "embeddings.patch_embeddings.projection.weight": E + "patch_embedding.weight",
         "embeddings.patch_embeddings.projection.bias": E + "patch_embedding.bias"}
    H = spec.vit_hidden
    split = {}
    for i in range(spec.vit_layers):
        s_, d_ = f"{L}{i}.", f"encoder.layer.{i}."
        for j, n in enumerate("qkv"):
            split[d_ + f"attention.{n}_proj.weight"] = (s_ + "attn.qkv.weight", j)
            split[d_ + f"attention.{n}_proj.bias"] = (s_ + "attn.qkv.bias", j)
        for a, b in (("attn.proj", "attention.projection_layer"), ("mlp.fc1", "mlp.fc1"), ("mlp.fc2", "mlp.fc2"),
                     ("norm1", "layernorm_before"), ("norm2", "layernorm_after")):
            m[d_ + b + ".weight"], m[d_ + b + ".bias"] = s_ + a + ".weight", s_ + a + ".bias"
        m[d_ + "lambda_1"], m[d_ + "lambda_2"] = s_ + "ls1", s_ + "ls2"
    state = {k: sd[v].detach().float() for k, v in m.items()}
    state.update({k: sd[v].detach().float()[j * H:(j + 1) * H] for k, (v, j) in split.items()})
    missing = hf.load_state_dict(state, strict=False)
    assert not missing.unexpected_keys and all(k.startswith("layernorm.") for k in missing.missing_keys), missing
    assert isinstance(hf.layernorm, torch.nn.Identity)        # the InternVL2-1B conversion: no final LayerNorm
    pcfg = InternVLConfig(vision_config=vcfg.to_dict(), text_config=dict(model_type="qwen2", hidden_size=spec.llm_hidden, num_hidden_layers=1,
                                                                          num_attention_heads=spec.llm_heads, vocab_size=128),
                          downsample_ratio=0.5, projector_hidden_act="gelu")
    proj = M.InternVLMultiModalProjector(pcfg).eval().float()
    pm = {"layer_norm.weight": "0.weight", "layer_norm.bias": "0.bias", "linear_1.weight": "1.weight", "linear_1.bias": "1.bias",
          "linear_2.weight": "3.weight", "linear_2.bias": "3.bias"}
    proj.load_state_dict({k: sd[MLP1_PREFIX + v].detach().float() for k, v in pm.items()})
    named = dict(hf.named_parameters())
    back = {v: named[k] for k, v in m.items()}
    back.update({MLP1_PREFIX + v: p for (k, v), p in zip(pm.items(), (dict(proj.named_parameters())[k] for k in pm))})
    qkv = {k: (named[k], v, j) for k, (v, j) in split.items()}

    def features(px):
        last = hf(pixel_values=px).last_hidden_state
        f = M.InternVLModel.pixel_shuffle(None, last[:, 1:].reshape(px.shape[0], 32, 32, -1), scale_factor=0.5)
        return last, proj(f.reshape(px.shape[0], -1, f.shape[-1]))

    return features, back, qkv


def test_oracle_vit_and_projector_match_hf_internvl_port(tiny):
    """InternViT + pixel-shuffle + mlp1 have no reference-side implementation offline (HF-Hub remote code).  An
    independent one exists in this image: transformers' HF-native InternVL port (``InternVLVisionModel``,
    ``InternVLModel.pixel_shuffle``, ``InternVLMultiModalProjector`` — different module tree and key names, same
    architecture; the conversion of OpenGVLab/InternVL2-1B uses ``use_mean_pooling=True`` = no final LayerNorm,
    ``attention_bias=True``, ``use_qk_norm=False``).  The synthetic weights are mapped key by key and both are run,
    forward and backward (gradient of a fixed random functional of the projected tokens w.r.t. every parameter)."""
    from simlingo_b200.spec import synth_frames, trainable
    spec, sd0 = tiny
    features, back, qkv = _hf_internvl(spec, sd0)
    px = synth_frames(spec, 1, 3).flatten(0, 2).float()                   # [2, 3, 448, 448]
    sd = {k: v.clone().requires_grad_(trainable(k) and ("vision_model" in k or "mlp1" in k)) for k, v in sd0.items()}
    want_last, want_proj = features(px)
    got_last = O.vit_forward(sd, spec, px)
    got_proj = O.extract_feature(sd, spec, px)
    assert got_last.shape == want_last.shape == (2, 1025, spec.vit_hidden)
    assert got_proj.shape == want_proj.shape == (2, spec.tokens_per_tile, spec.llm_hidden)
    assert (got_last - want_last).abs().max().item() <= 2e-4 * want_last.abs().max().item()
    assert (got_proj - want_proj).abs().max().item() <= 2e-4 * want_proj.abs().max().item()
    r = torch.randn(want_proj.shape, generator=torch.Generator().manual_seed(9))
    (want_proj * r).sum().backward()
    (got_proj * r).sum().backward()
    checked, scales = 0, []
    for key, p in back.items():
        g, w = sd[key].grad, p.grad
        assert g is not None and w is not None, key
        assert (g - w.view_as(g)).abs().max().item() <= 1e-3 * w.abs().max().item() + 1e-6, key
        scales.append(w.abs().max().item())
        checked += 1
    H = spec.vit_hidden
    for k, (p, key, j) in qkv.items():
        g = sd[key].grad[j * H:(j + 1) * H]
        # the key bias has a mathematically zero gradient (softmax is shift invariant): both sides hold ~1e-8 of round-off
        assert (g - p.grad).abs().max().item() <= 1e-3 * p.grad.abs().max().item() + 1e-6, k
        checked += 1
    assert checked == 4 + spec.vit_layers * (12 + 6) + 6 and min(scales) > 1e-4      # the 1e-6 floor is far below every real gradient
