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
