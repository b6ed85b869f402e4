"""Decode path on a B200: (1) the fused RoPE + KV-write + attention launch of the per-kernel chain against the two launches it
replaces, (2) the opt-in persistent decode kernel (``slb_decode_loop``: token loop, grid barriers, device-side EOS) against the
per-kernel chain on the same engine - tokens identical, caches / waypoints within bf16 noise - for batch 1 (key segments over many CTAs),
a ragged batch (5) and the largest batch (32, two 16-row MMA tiles).  Both paths are pinned against the fp32 oracle elsewhere
(test_model_gpu, test_dropin_gpu, test_fullscale_gpu); this file pins them against each other at sizes the oracle would not finish."""
import pytest
import torch

from simlingo_b200.spec import LMHEAD_SHIFT, init_state_dict, synth_frames, synth_placeholders, synth_prompt_ids, tiny_spec

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("batch,masked", [(1, False), (2, True), (8, False), (8, True)])
def test_fused_rope_attention_step_matches_two_launches(batch, masked):
    """slb_attn_decode_rope == slb_rope_kv_write + slb_attn_gqa(lq = 1): same cache rows, same attention output (the fused kernel
    keeps the rotated q in fp32 where the two-launch path rounds it to bf16: tolerance of one bf16 step on q)."""
    from simlingo_b200 import lib
    torch.manual_seed(batch * 7 + masked)
    hq, hkv, lmax, past = 14, 2, 640, 577
    qkv = (torch.randn(batch, (hq + 2 * hkv) * 64, device="cuda") * 1.5).to(torch.bfloat16)
    kc = torch.randn(batch, hkv, lmax, 64, device="cuda").to(torch.bfloat16)
    vc = torch.randn(batch, hkv, lmax, 64, device="cuda").to(torch.bfloat16)
    key_valid = None
    if masked:
        key_valid = torch.ones(batch, lmax, device="cuda", dtype=torch.uint8)
        key_valid[-1, :37] = 0
    pos = torch.tensor([past], device="cuda", dtype=torch.int32)
    q2, k2, v2 = qkv.clone(), kc.clone(), vc.clone()
    lib.rope_kv_write(q2, k2, v2, batch, 1, past, hq, hkv, 1.0e6)
    ref = lib.attn_gqa(q2, q2.shape[1], k2, v2, batch, 1, past, hq, hkv, key_valid=key_valid)
    for past_dev in (None, pos):
        k1, v1 = kc.clone(), vc.clone()
        got = lib.attn_decode_rope(qkv.clone(), k1, v1, batch, past if past_dev is None else lmax - 1, hq, hkv, 1.0e6, key_valid=key_valid,
                                   past_dev=past_dev)
        torch.cuda.synchronize()
        assert torch.equal(v1, v2)
        assert relerr(k1[:, :, past], k2[:, :, past]) < 1e-2 and torch.equal(k1[:, :, :past], k2[:, :, :past])
        assert torch.equal(k1[:, :, past + 1:], kc[:, :, past + 1:])
        assert relerr(got, ref) < 2e-2, relerr(got, ref)


@pytest.mark.parametrize("B,G,use_eos,max_new", [(1, 5, True, 9), (1, 12, False, 12), (5, 6, False, 6), (32, 7, False, 7)])
def test_persistent_decode_kernel_matches_kernel_chain(setup, B, G, use_eos, max_new):
    spec, sd, eng = setup
    eos = spec.eos_id
    ids = synth_prompt_ids(spec, B, seed=41 + B)
    ids[:, -1] = (eos - G * LMHEAD_SHIFT) % spec.vocab          # planted walk: EOS would be the G-th generated token
    valid = torch.ones_like(ids, dtype=torch.bool)
    fr, ph = synth_frames(spec, B, 41).to("cuda", torch.bfloat16), synth_placeholders(spec, B, 41)
    run = lambda: eng.driving_forward(fr, ids.cuda(), valid.cuda(), ph, max_new_tokens=max_new, eos_token_id=eos if use_eos else None,
                                      ids_cpu=ids)
    key = ("gen", B, ids.shape[1], max_new, eos if use_eos else None)
    res = {}
    old = eng.decode_mega
    try:
        for mega in (False, True):
            eng.decode_mega = mega
            l0 = eng.launches
            for _ in range(3):   # eager sighting, capture + replay, replay
                sp, rt, tok = run()
            rec = eng._graphs[key]
            assert rec["mega"] == mega
            torch.cuda.synchronize()
            res[mega] = dict(sp=sp.clone(), rt=rt.clone(), tok=[t.cpu().tolist() for t in tok], kc=rec["cache"][0].clone(),
                             vc=rec["cache"][1].clone(), pos=int(rec["pos"]), step=int(rec["step"]), n_gen=rec["n_gen"].cpu().tolist(),
                             sampled=rec["sampled"].cpu().clone(), launches=eng.launches - l0)
    finally:
        eng.decode_mega = old
    a, b = res[False], res[True]
    assert b["tok"] == a["tok"]
    assert len(b["tok"][0]) == (G if use_eos else max_new)
    assert b["pos"] == a["pos"] and b["step"] == a["step"] and b["n_gen"] == a["n_gen"]
    assert torch.equal(b["sampled"], a["sampled"])
    L = ids.shape[1]
    n_cached = a["pos"]   # prompt + generated tokens whose K/V were written by the decode steps (the query pass writes beyond)
    assert n_cached == L + (G if use_eos else max_new) - 1
    assert torch.equal(b["kc"][:, :, :, :L], a["kc"][:, :, :, :L])                     # prefill rows: the same graph in both runs
    assert relerr(b["kc"][:, :, :, L:n_cached], a["kc"][:, :, :, L:n_cached]) < 2e-2   # decode rows: other summation order
    assert relerr(b["vc"][:, :, :, L:n_cached], a["vc"][:, :, :, L:n_cached]) < 2e-2
    assert relerr(b["sp"], a["sp"]) < 1e-2 and relerr(b["rt"], a["rt"]) < 1e-2, (relerr(b["sp"], a["sp"]), relerr(b["rt"], a["rt"]))
    assert b["launches"] < a["launches"]


def test_persistent_decode_kernel_against_oracle(setup):
    """agent case end to end through the persistent kernel: tokens identical to the fp32 oracle's no-cache loop, waypoints 2e-2"""
    from oracle import model as O
    spec, sd, eng = setup
    eos = spec.eos_id
    ids = synth_prompt_ids(spec, 1, seed=77)
    ids[:, -1] = (eos - 4 * LMHEAD_SHIFT) % spec.vocab
    valid = torch.ones_like(ids, dtype=torch.bool)
    fr, ph = synth_frames(spec, 1, 77), synth_placeholders(spec, 1, 77)
    with torch.no_grad():
        sp_ref, rt_ref, tok_ref = O.driving_forward(sd, spec, fr, ids, valid, ph, max_new_tokens=8, eos_token_id=eos)
    old, eng.decode_mega = eng.decode_mega, True
    try:
        for _ in range(2):
            sp, rt, tok = eng.driving_forward(fr.to("cuda", torch.bfloat16), ids.cuda(), valid.cuda(), ph, max_new_tokens=8, eos_token_id=eos,
                                              ids_cpu=ids)
        assert eng._graphs[("gen", 1, ids.shape[1], 8, eos)]["mega"]
    finally:
        eng.decode_mega = old
    assert [t.cpu().tolist() for t in tok] == [t.tolist() for t in tok_ref]
    assert relerr(sp, sp_ref) < 2e-2 and relerr(rt, rt_ref) < 2e-2
