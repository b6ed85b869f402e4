"""The drop-in boundary on a B200: simlingo_training.models.driving.DrivingModel (same constructor / forward
signatures / state_dict keys as the reference) against the committed golden fixture produced by the reference's
own code (tests/golden/make_golden.py).  Tolerance: tokens identical; waypoints / route / loss max rel err 2e-2."""
import os

import numpy as np
import pytest
import torch

from simlingo_b200.spec import init_state_dict, state_dict_schema, tiny_spec
from tests.helpers import StubTokenizer, make_case_inputs, to_driving_input

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_run.pt")
TOL = 2e-2


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def model():
    from simlingo_b200.modules import register_variant
    from simlingo_training.models.driving import DrivingModel
    spec = tiny_spec(2, 2, 4096)
    name = "internvl2-tiny-test"
    register_variant(name, spec)
    cfg = dict(vision_model=dict(_target_="simlingo_training.models.encoder.vlm.VLMEncoderModel", variant=name, embed_dim=512, freeze=False),
               language_model=dict(_target_="simlingo_training.models.language_model.llm.LLM", variant=name, lora=True, lora_alpha=64,
                                   lora_r=32, lora_dropout=0.1),
               lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999), pct_start=0.05, speed_wps_mode="2d", predict_route_as_wps=True)
    torch.set_default_dtype(torch.bfloat16)
    try:
        m = DrivingModel(cfg_data_module={"use_global_img": False}, processor=StubTokenizer(spec), cache_dir=None, **cfg)
    finally:
        torch.set_default_dtype(torch.float32)
    sd = init_state_dict(spec, seed=0, with_aliases=True)
    assert set(m.state_dict().keys()) == set(state_dict_schema(spec).keys())
    m.load_state_dict(sd, strict=True)
    return spec, m.to("cuda").eval()


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_forward_matches_reference_run(model, golden, idx):
    spec, m = model
    case = [c for c in golden["cases"] if c["kind"] == "forward"][idx]
    inp = make_case_inputs(spec, case["B"], case["seed"], case["G_list"], pad_rows=[tuple(p) for p in case["pads"]])
    sp, rt, lang = m(to_driving_input(inp, "cuda", torch.bfloat16))
    assert lang == case["language"]
    assert relerr(sp, case["speed_wps"]) < TOL and relerr(rt, case["route"]) < TOL


@pytest.mark.parametrize("idx", [0, 1])
def test_forward_loss_matches_reference_run(model, golden, idx):
    from simlingo_training.utils.custom_types import DrivingExample, DrivingLabel
    spec, m = model
    case = [c for c in golden["cases"] if c["kind"] == "loss"][idx]
    inp = make_case_inputs(spec, case["B"], case["seed"], None, answer_len=16, pad_rows=[tuple(p) for p in case["pads"]])
    di = to_driving_input(inp, "cuda", torch.bfloat16)
    wps, path = inp["labels"]
    ex = DrivingExample(di, DrivingLabel(wps.cuda(), path.cuda(), di.prompt, torch.zeros(1)), ["x"] * case["B"])
    with torch.no_grad():
        out, _ = m.forward_loss(ex)
    assert relerr(out.loss, case["loss"]) < TOL
    for k, v in case["averages"].items():
        assert relerr(out.loss_averages[k], v) < TOL, k


def test_submodule_entry_points(model):
    """LLM.forward / greedy_sample and extract_feature are callable on their own like in the reference."""
    from oracle import model as O
    spec, m = model
    sd = init_state_dict(spec, seed=0)
    x = torch.randn(1, 40, spec.llm_hidden, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    with torch.no_grad():
        f_ref, lg_ref = O.llm_forward(sd, spec, x.float(), None)
        feats, logits = m.language_model.forward(x.cuda())
    assert relerr(feats, f_ref) < TOL and relerr(logits, lg_ref) < TOL
    emb = m.adaptors.language.embed_tokens(torch.tensor([[5, 9]], device="cuda"))
    assert torch.equal(emb.cpu().float(), sd["language_model.model.base_model.model.model.embed_tokens.weight"][[5, 9]][None])
    toks, grown = m.language_model.greedy_sample(x.cuda(), eos_token_id=spec.eos_id, max_new_tokens=3,
                                                 input_embed_matrix=m.adaptors.language.embed_tokens.weight,
                                                 logit_matrix=m.adaptors.language.lm_head.weight,
                                                 attention_mask=torch.ones(1, 40, dtype=torch.bool, device="cuda"))
    with torch.no_grad():
        t_ref, g_ref = O.greedy_sample(sd, spec, x.float(), 3, spec.eos_id, torch.ones(1, 40, dtype=torch.bool))
    assert toks.cpu().tolist() == t_ref.tolist() and grown.shape == g_ref.shape


def test_predict_step_resamples_route_like_the_reference(model, golden):
    """predict_step (reference driving.py:285-328): forward + per-item ``equal_spacing_route`` — here one kernel for the
    batch (``slb_equal_spacing_route``); the record it keeps must equal the numpy restatement applied to forward's route"""
    from oracle.postprocess import equal_spacing_route
    from simlingo_training.utils.custom_types import DrivingExample, DrivingLabel
    spec, m = model
    case = [c for c in golden["cases"] if c["kind"] == "forward"][2]   # ragged B=3
    inp = make_case_inputs(spec, case["B"], case["seed"], case["G_list"], pad_rows=[tuple(p) for p in case["pads"]])
    di = to_driving_input(inp, "cuda", torch.bfloat16)
    B = case["B"]
    run_id = torch.zeros(B, 8, dtype=torch.uint8)
    run_id[:, :3] = torch.tensor(list(b"r01"), dtype=torch.uint8)
    label = DrivingLabel(torch.zeros(B, 10, 2), torch.zeros(B, 20, 2), di.prompt, torch.zeros(1), {"k": [0] * B})
    m.prediction = {}
    sp, rt, lang, sp_gt, rt_gt, lang_gt = m.predict_step(DrivingExample(di, label, run_id, None))
    assert rt.dtype == torch.float64 and rt.shape == (B, 20, 2) and rt.is_cuda and lang == case["language"]
    raw = m.route.float().cpu().numpy()                                  # what forward produced (fp32)
    want = np.stack([equal_spacing_route(r) for r in raw])
    np.testing.assert_allclose(rt.cpu().numpy(), want, rtol=0, atol=1e-12)
    assert m.prediction["path"] == ["r01"] * B and len(m.prediction["route"]) == 1
