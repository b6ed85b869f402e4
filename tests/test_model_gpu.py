"""Model-level parity on a B200: the CUDA engine (bf16) against the fp32 CPU oracle on identical
random-init weights and synthetic frames + prompts (tiny_spec: real widths, 2+2 layers, 4096 vocab,
so the oracle finishes in seconds).  Tolerance from BASELINE.json north_star: max rel err 2e-2 on
waypoints / route / logits (max-abs error over the tensor's max magnitude), greedy tokens identical."""
import pytest
import torch

from simlingo_b200.spec import (LLM_PREFIX, LMHEAD_SHIFT, init_state_dict, synth_frames, synth_placeholders,
                                synth_prompt_ids, tiny_spec)

pytestmark = pytest.mark.gpu
TOL = 2e-2


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def setup():
    from simlingo_b200.engine import Engine
    spec = tiny_spec(2, 2, 4096)
    sd = init_state_dict(spec, seed=0)
    sd_gpu = {k: v.to("cuda", torch.bfloat16) for k, v in sd.items()}
    return spec, sd, Engine(sd_gpu, spec)


def test_vit_and_projector(setup):
    from oracle import model as O
    spec, sd, eng = setup
    px = synth_frames(spec, 1, seed=3).reshape(2, 3, 448, 448)
    ref_layers = []
    with torch.no_grad():
        ref = O.extract_feature(sd, spec, px, ref_layers)
    got_layers = []
    eng.vit(px.to("cuda", torch.bfloat16), got_layers)
    for i, (g, r) in enumerate(zip(got_layers, ref_layers)):
        assert relerr(g.view(2, 1025, 1024), r) < TOL, f"vit layer {i}"
    got = eng.extract_feature(px.to("cuda", torch.bfloat16))
    assert relerr(got.view(2, 256, 896), ref) < TOL


def test_teacher_forced_pass_with_padding(setup):
    """forward_model semantics: [valid language | 30 queries | pads], key-padding mask, arange positions."""
    from oracle import model as O
    spec, sd, eng = setup
    B = 2
    ids = synth_prompt_ids(spec, B, seed=5, answer_len=16)
    valid = torch.ones_like(ids, dtype=torch.bool)
    valid[1, :9] = False
    lm = torch.zeros_like(valid)
    lm[:, -16:] = True
    fr, ph = synth_frames(spec, B, 5), synth_placeholders(spec, B, 5)
    with torch.no_grad():
        ad = O.adaptor_list_forward(sd, spec, ids, valid, lm)
        ad, feats, logits = O.forward_model(sd, spec, ad, fr, ph)
    emb = eng.embed_prompt(ids.cuda(), fr.to("cuda", torch.bfloat16), ph, ids)
    assert relerr(emb, ad["language_inputs"]) < TOL
    inputs = ad["inputs"].to("cuda", torch.bfloat16)
    rows = torch.arange(0, inputs.shape[1], 7, device="cuda")
    f, lg = eng.forward_model(inputs, ad["inputs_mask"].cuda(), want_logits_rows=None)
    m = ad["inputs_mask"]
    assert relerr(f.cpu()[m], feats[m]) < TOL
    lg = eng.logits(f[0, rows].contiguous())
    assert relerr(lg, logits[0, rows.cpu()]) < TOL


@pytest.mark.parametrize("G", [1, 4])
def test_driving_forward_tokens_and_waypoints(setup, G):
    """DrivingModel.forward: planted next-token walk reaches EOS after G tokens; tokens must be identical to the
    oracle's (no-KV-cache, full re-forward) greedy loop and waypoints / route within tolerance."""
    from oracle import model as O
    spec, sd, eng = setup
    eos = spec.eos_id
    last = (eos - G * LMHEAD_SHIFT) % spec.vocab
    ids = synth_prompt_ids(spec, 1, seed=11, last_token=last)
    valid = torch.ones_like(ids, dtype=torch.bool)
    fr, ph = synth_frames(spec, 1, 11), synth_placeholders(spec, 1, 11)
    margins = []
    with torch.no_grad():
        sp_ref, rt_ref, tok_ref = O.driving_forward(sd, spec, fr, ids, valid, ph, max_new_tokens=8, eos_token_id=eos, margins=margins)
    assert len(tok_ref[0]) == G and int(tok_ref[0][-1]) == eos
    assert min(margins) > 50.0, margins  # planted margin >> bf16 noise
    sp, rt, tok = eng.driving_forward(fr.to("cuda", torch.bfloat16), ids.cuda(), valid.cuda(), ph, max_new_tokens=8,
                                      eos_token_id=eos, ids_cpu=ids)
    assert tok[0].cpu().tolist() == tok_ref[0].tolist()
    assert relerr(sp, sp_ref) < TOL and relerr(rt, rt_ref) < TOL


def test_driving_forward_batched_ragged_and_padded(setup):
    """B=3 with different generation lengths (ragged EOS) and one left-padded prompt: must match the reference's
    per-item loop, including its no-mask final pass for the padded row."""
    from oracle import model as O
    spec, sd, eng = setup
    eos = spec.eos_id
    ids = synth_prompt_ids(spec, 3, seed=21)
    for b, G in enumerate([2, 3, 2]):
        ids[b, -1] = (eos - G * LMHEAD_SHIFT) % spec.vocab
    valid = torch.ones_like(ids, dtype=torch.bool)
    fr, ph = synth_frames(spec, 3, 21), synth_placeholders(spec, 3, 21)
    with torch.no_grad():
        sp_ref, rt_ref, tok_ref = O.driving_forward(sd, spec, fr, ids, valid, ph, max_new_tokens=6, eos_token_id=eos)
    sp, rt, tok = eng.driving_forward(fr.to("cuda", torch.bfloat16), ids.cuda(), valid.cuda(), ph, max_new_tokens=6,
                                      eos_token_id=eos, ids_cpu=ids)
    assert [t.cpu().tolist() for t in tok] == [t.tolist() for t in tok_ref]
    assert relerr(sp, sp_ref) < TOL and relerr(rt, rt_ref) < TOL
    # left-padded variant (eval path, datamodule.py:138): first 6 tokens of row 1 are pads
    valid[1, :6] = False
    with torch.no_grad():
        sp_ref, rt_ref, tok_ref = O.driving_forward(sd, spec, fr, ids, valid, ph, max_new_tokens=6, eos_token_id=eos)
    sp, rt, tok = eng.driving_forward(fr.to("cuda", torch.bfloat16), ids.cuda(), valid.cuda(), ph, max_new_tokens=6,
                                      eos_token_id=eos, ids_cpu=ids)
    assert [t.cpu().tolist() for t in tok] == [t.tolist() for t in tok_ref]
    assert relerr(sp, sp_ref) < TOL and relerr(rt, rt_ref) < TOL


def test_cuda_graph_generation_equals_eager_launches(setup):
    """From the second call with a shape the engine replays CUDA graphs (ViT on few tiles, prefill, one decode-step graph
    per token with the position counter on the device, query append): tokens identical and waypoints equal to the
    eagerly launched kernels, for the agent case (batch 1, stops at EOS) and the language case (batch 3, EOS suppressed).
    This pins the per-kernel chain (bit-equal to its eager launches); the opt-in persistent decode kernel is compared with the
    chain in test_decode_gpu.py."""
    spec, sd, eng = setup
    eos = spec.eos_id
    mega_default, eng.decode_mega = eng.decode_mega, False
    for B, G, use_eos, max_new in [(1, 5, True, 9), (3, 6, False, 6)]:
        ids = synth_prompt_ids(spec, B, seed=31)
        ids[:, -1] = (eos - G * LMHEAD_SHIFT) % spec.vocab
        valid = torch.ones_like(ids, dtype=torch.bool)
        fr, ph = synth_frames(spec, B, 31).to("cuda", torch.bfloat16), synth_placeholders(spec, B, 31)
        run = lambda: eng.driving_forward(fr, ids.cuda(), valid.cuda(), ph, max_new_tokens=max_new, eos_token_id=eos if use_eos else None,
                                          ids_cpu=ids)
        eng.graphs_enabled = False
        sp0, rt0, tok0 = run()
        eng.graphs_enabled = True
        r0 = eng.graph_replays
        for _ in range(3):   # eager sighting, capture + replay, replay
            sp, rt, tok = run()
        assert eng.graph_replays - r0 >= 2 * (2 + (G if use_eos else max_new) - 1)
        assert [t.cpu().tolist() for t in tok] == [t.cpu().tolist() for t in tok0]
        assert len(tok[0]) == (G if use_eos else max_new)
        assert torch.equal(sp, sp0) and torch.equal(rt, rt0)
    eng.decode_mega = mega_default
