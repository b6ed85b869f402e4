"""CPU-only checks (run with -m "not gpu"): the C-ABI library loads and exports every symbol the header declares,
state_dict schema / drop-in module tree, host-side glue (adaptors, losses, sharding), world_size-2 gloo path."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from simlingo_b200 import build
    lib_path = build.build()
    so = ctypes.CDLL(str(lib_path))
    header = open(os.path.join(ROOT, "include", "simlingo_b200.h")).read()
    names = sorted(set(re.findall(r"\b(slb_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(so, n), n
    so.slb_last_error.restype = ctypes.c_char_p
    assert so.slb_version() >= 100


def test_invalid_arguments_return_error_codes_not_crashes():
    from simlingo_b200 import lib
    L = lib.load()
    L.slb_last_error.restype = ctypes.c_char_p
    g = lib.GemmArgs(M=0, N=8, K=8)
    assert L.slb_gemm_bf16(ctypes.byref(g), None) != 0
    assert b"bad shape" in L.slb_last_error()
    assert L.slb_layernorm_fwd(None, None, None, None, 4, 7, ctypes.c_float(1e-6), None, None, None) != 0  # cols % 8
    # post-processing entry points: argument validation happens before any CUDA call
    from simlingo_b200.postprocess import ControlParams
    buf = (ctypes.c_double * 64)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    prm = ControlParams(0, 2, 1.0, 2.0, 24.0, 105.0, 0.1)
    assert L.slb_control_inputs(ptr, ptr, ptr, 1, 1, 10, ctypes.byref(prm), ptr, None) != 0      # a route needs >= 2 points
    assert b"route needs" in L.slb_last_error()
    assert L.slb_control_inputs(ptr, ptr, ptr, 1, 20, 2, ctypes.byref(prm), ptr, None) != 0      # wp_b = 2 outside 2 waypoints
    assert b"indices out of range" in L.slb_last_error()
    bad = ControlParams(0, 2, 1.0, 2.0, 24.0, 105.0, 0.0)
    assert L.slb_control_inputs(ptr, ptr, ptr, 1, 20, 10, ctypes.byref(bad), ptr, None) != 0     # zero sampling step
    assert L.slb_equal_spacing_route(ptr, 0, 20, 20, ptr, None) != 0 and L.slb_equal_spacing_route(ptr, 1, 64, 20, ptr, None) != 0
    assert L.slb_attn_delta(ctypes.c_void_p(ptr.value + 2), ptr, ptr, 1, 1, 1, None) != 0          # operands must be 16-byte aligned
    assert b"16-byte" in L.slb_last_error()
    # persistent decode kernel: argument validation precedes any CUDA call
    assert L.slb_decode_loop(None, None) != 0
    d = lib.DecodeArgs(n_layers=2, batch=33)
    assert L.slb_decode_loop(ctypes.byref(d), None) != 0 and b"decode_loop" in L.slb_last_error()


def test_state_dict_schema_matches_dropin_module_tree():
    from simlingo_b200.modules import register_variant
    from simlingo_b200.spec import INTERNVL2_1B, init_state_dict, state_dict_schema, tiny_spec, trainable
    from simlingo_training.models.driving import DrivingModel
    from tests.helpers import StubTokenizer
    full = state_dict_schema(INTERNVL2_1B)
    assert 980 <= len(full) <= 1000  # "about 990 keys"
    spec = tiny_spec(2, 2, 4096)
    register_variant("internvl2-tiny-cpu", spec)
    cfg = dict(vision_model=dict(_target_="simlingo_training.models.encoder.vlm.VLMEncoderModel", variant="internvl2-tiny-cpu", embed_dim=512, freeze=False),
               language_model=dict(_target_="simlingo_training.models.language_model.llm.LLM", variant="internvl2-tiny-cpu", lora=True, lora_alpha=64,
                                   lora_r=32, lora_dropout=0.1),
               lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999), pct_start=0.05, speed_wps_mode="2d", predict_route_as_wps=True)
    m = DrivingModel(cfg_data_module={"use_global_img": False}, processor=StubTokenizer(spec), cache_dir=None, **cfg)
    schema = state_dict_schema(spec)
    sd = m.state_dict()
    assert set(sd.keys()) == set(schema.keys())
    for k, (shape, _) in schema.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    m.load_state_dict(init_state_dict(spec, with_aliases=True), strict=True)
    # aliases share storage (llm.py:91, adaptors.py:227-229)
    assert m.adaptors.language.lm_head.weight is m.language_model.model.base_model.model.lm_head.weight
    assert m.adaptors.language.embed_tokens.weight.data_ptr() == sd["language_model.model.base_model.model.embed_tokens.weight"].data_ptr()
    # trainability: ViT + mlp1 + LoRA + heads + wp_encoder train, Qwen2 base / embeddings / lm_head are frozen
    for n, p in m.named_parameters():
        assert p.requires_grad == trainable(n), n
    # no CPU fallback: using the model without a GPU fails loudly
    with pytest.raises(RuntimeError):
        m.adaptors.language.embed_tokens(torch.zeros(1, 2, dtype=torch.long))


def test_adaptor_list_matches_oracle_glue():
    """AdaptorList permutation / split (host index glue) against the oracle's restatement, with stub adaptors."""
    from oracle import model as O
    from simlingo_b200.spec import init_state_dict, tiny_spec
    from simlingo_training.models.adaptors.adaptors import AdaptorList, DrivingAdaptor
    spec = tiny_spec(1, 1, 512)
    sd = init_state_dict(spec)
    B, L = 3, 9
    ids = torch.randint(0, 500, (B, L))
    valid = torch.ones(B, L, dtype=torch.bool)
    valid[1, :3] = False
    valid[2, :5] = False

    class Lang(torch.nn.Module):
        def forward(self, example, **kw):
            return {"inputs": O.language_embed(sd, spec, ids), "inputs_mask": valid, "_ids": ids, "_ids_mask": torch.zeros_like(valid)}

    drv = DrivingAdaptor(spec.llm_hidden, speed_wps_mode="2d", predict_route_as_wps=True)
    with torch.no_grad():
        drv.query_embeds_wps.copy_(sd["adaptors.driving.query_embeds_wps"])
        drv.query_embeds_speed.copy_(sd["adaptors.driving.query_embeds_speed"])
    al = AdaptorList(language=Lang(), driving=drv)

    class Ex:
        camera_images = torch.zeros(B, 1)
    out = al(Ex())
    ref = O.adaptor_list_forward(sd, spec, ids, valid, torch.zeros_like(valid))
    assert torch.equal(out["perm"], ref["perm"]) and torch.equal(out["inputs_mask"], ref["inputs_mask"])
    assert torch.allclose(out["inputs"], ref["inputs"])
    feats = torch.randn(B, L + 30, 8)
    a, b = O.split_outputs(ref, feats)
    sp = al.split_outputs_by_adaptor(out, feats)
    assert torch.equal(sp["language"], a) and torch.equal(sp["driving"], b)


def test_summarise_losses_and_helpers():
    from simlingo_training.models.adaptors.adaptors import cross_track_error
    from simlingo_training.models.utils import summarise_losses
    v1, n1 = torch.tensor([2.0, 4.0]), torch.tensor([1, 3])
    v2, n2 = torch.tensor([0.0, 0.0]), torch.tensor([0, 0])
    out = summarise_losses({"a_loss": (v1, n1), "b_loss": (v2, n2)})
    assert torch.isclose(out.loss, torch.tensor(1.5)) and float(out.loss_averages["b_loss"]) == 0.0
    path = torch.stack([torch.arange(5.0), torch.zeros(5)], -1)[None]
    pts = torch.tensor([[[2.0, 1.0], [3.0, -2.0]]])
    assert torch.allclose(cross_track_error(pts, path), torch.tensor([[1.0, 2.0]]))


def test_shard_ranges_cover_everything():
    from simlingo_b200.dist import shard_range
    for n, w in [(64, 8), (64, 3), (5, 8), (7, 2)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from simlingo_b200.dist import env_rank_world, gather_predictions, max_over_ranks, shard_range
rank, world, _ = env_rank_world()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
n_items = 7 if world <= 7 else 2 * world - 1      # ragged shards at every world size
lo, hi = shard_range(n_items, rank, world)
local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 2, 2).contiguous()
ms = max_over_ranks(10.0 + rank)
assert ms == 10.0 + world - 1, ms
out = gather_predictions(local, [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)])
if rank == 0:
    full = torch.cat(out)
    assert full[:, 0, 0].tolist() == [float(i) for i in range(n_items)], full
# gradient all-reduce semantics of the training path (sum then / world), bf16 buckets
g = torch.full((16,), float(rank + 1), dtype=torch.bfloat16)
dist.all_reduce(g)
g /= world
assert torch.allclose(g.float(), torch.full((16,), (world + 1) / 2.0))
# flat parameter / gradient store: layout in backward-completion order, bucketed all-reduce as groups finish
from simlingo_b200.spec import tiny_spec, LLM_PREFIX, VIT_PREFIX
from simlingo_b200.training import ParamStore
from tests.helpers import build_drop_in_model
spec = tiny_spec(2, 2, 512)
model = build_drop_in_model(spec, "internvl2-tiny-dp", device="cpu")
store = ParamStore(model, "", spec, bucket_bytes=8 << 20, allow_cpu=True)
names = [g.name for g in store.groups]
assert names == ["llm1", "llm0", "mlp1", "vit1", "vit0", "vit_emb", "other"], names
assert all(p.data_ptr() == store.flat_param[o:o + n].data_ptr() for k, p in store.params.items() for o, n in [store.offsets[k]])
assert all(o % 8 == 0 for o, _ in store.offsets.values())
assert len(store._buckets) >= 3 and store._buckets[0][0] == 0 and store._buckets[-1][1] == store.numel
store.enable_data_parallel()
store.begin_backward()
store.flat_grad.fill_(float(rank + 1))
order = []
for gi, g in enumerate(store.groups[:-1]):
    before = store.n_allreduce
    store.group_ready(gi)
    order.append(store.n_allreduce - before)
assert sum(order) == len(store._buckets) - 1, (order, store._buckets)   # the tail bucket waits for the end of backward
store.finish_backward()
store.wait_exchange()
assert store.n_allreduce == len(store._buckets)
assert torch.equal(store.flat_grad.float(), torch.full((store.numel,), float(sum(range(1, world + 1)))))
assert all(p.grad.data_ptr() == store.grad_view[k].data_ptr() for k, p in store.params.items())
# gradient accumulation under data parallelism: a second backward on already-reduced sums must be refused ...
try:
    store.begin_backward()
    raise AssertionError("second backward after an exchange was accepted")
except RuntimeError as e:
    assert "no_sync" in str(e)
# ... and under no_sync() the exchange is deferred to the last micro-batch: world * (g1 + g2) is never formed
store.zero_grad()
assert float(store.flat_grad.abs().max()) == 0.0
n0 = store.n_allreduce
with store.no_sync():
    store.begin_backward()
    store.flat_grad.add_(float(rank + 1))
    for gi in range(len(store.groups) - 1):
        store.group_ready(gi)
    store.finish_backward()
assert store.n_allreduce == n0 and store.accumulate
store.begin_backward()
store.flat_grad.add_(float(rank + 1))
for gi in range(len(store.groups) - 1):
    store.group_ready(gi)
store.finish_backward()
store.wait_exchange()
assert store.n_allreduce == n0 + len(store._buckets)
assert torch.equal(store.flat_grad.float(), torch.full((store.numel,), 2.0 * sum(range(1, world + 1))))
assert store.update_ranges() == [(0, store.numel)]
dist.destroy_process_group()
print("ok", rank)
"""


def _run_gloo_world(world: int) -> None:
    port = 29500 + (os.getpid() % 400) + 10 * world
    code = _WORKER.format(root=ROOT, port=port)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    try:
        for p in procs:
            out, _ = p.communicate(timeout=600)
            assert p.returncode == 0, out
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()


def test_world_size_2_gloo():
    _run_gloo_world(2)


def test_world_size_8_gloo():
    """the same shard / max-over-ranks / bucket-protocol checks at the reference's training world size (8 ranks, train_simlingo_seed1.sh:3-5):
    every range of the flat gradient buffer reduced exactly once, sum over 8 ranks exact, no_sync accumulation, double-reduction refusal"""
    _run_gloo_world(8)


def test_store_tracks_which_parameters_received_gradients():
    """torch.optim.AdamW (the reference's optimizer) skips parameters whose .grad is None; the flat store therefore reports
    to the fused optimizer only the ranges of groups a backward reached since zero_grad, minus torch-managed parameters
    autograd left without a gradient.  Host logic on a CPU store."""
    from simlingo_b200.spec import tiny_spec
    from simlingo_b200.training import ParamStore
    from tests.helpers import build_drop_in_model
    spec = tiny_spec(2, 2, 512)
    model = build_drop_in_model(spec, "internvl2-tiny-touch", device="cpu")
    store = ParamStore(model, "", spec, bucket_bytes=8 << 20, allow_cpu=True)
    store.zero_grad()
    assert store.update_ranges() == []
    # an LLM-only backward: the two decoder groups and the torch-managed tail, where only the route head got a gradient
    store.begin_backward()
    store._touched.update({store.group_index["llm1"], store.group_index["llm0"]})
    for k, p in store.params.items():
        if k.startswith("adaptors.driving.route_head"):
            p.grad = torch.ones_like(p)
    store.finish_backward()
    rng = store.update_ranges()
    g0, g1 = store.groups[store.group_index["llm1"]], store.groups[store.group_index["llm0"]]
    assert rng[0] == (g1.start if g1.start < g0.start else g0.start, max(g0.end, g1.end))
    covered = lambda o: any(a <= o < b for a, b in rng)
    for k in store.params:
        o, _ = store.offsets[k]
        want = k.startswith("language_model.") or k.startswith("adaptors.driving.route_head")
        assert covered(o) == want, k
    assert float(store.grad_view["adaptors.driving.route_head.0.bias"].min()) == 1.0
    store.zero_grad()   # everything is cleared again, stale gradients cannot survive a step
    assert float(store.flat_grad.abs().max()) == 0.0 and store.update_ranges() == []


def test_fused_adamw_state_dict_interoperates_with_torch_adamw():
    """Checkpoint / resume: FusedAdamW.state_dict() has torch.optim.AdamW's layout over the reference's parameter list
    (all of ``self.parameters()``, frozen ones included: driving.py:718), so either optimizer can resume from the other.
    Host logic only (CPU store, no kernels)."""
    from simlingo_b200.optim import FusedAdamW
    from simlingo_b200.spec import tiny_spec
    from simlingo_b200.training import ParamStore
    from tests.helpers import build_drop_in_model
    spec = tiny_spec(2, 2, 512)
    model = build_drop_in_model(spec, "internvl2-tiny-optim", device="cpu")
    store = ParamStore(model, "", spec, bucket_bytes=8 << 20, allow_cpu=True)
    params = list(model.parameters())
    trainable = [i for i, p in enumerate(params) if p.requires_grad]
    assert 0 < len(trainable) < len(params)
    opt = FusedAdamW(params, store, lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999), max_grad_norm=0.3)
    assert opt.state_dict()["state"] == {}                       # like torch before the first step

    # a torch AdamW the reference's way (fp32 copies of the same parameter list), two real steps on the trainable ones
    g = torch.Generator().manual_seed(0)
    ref_params = [torch.nn.Parameter(p.detach().float().clone(), requires_grad=p.requires_grad) for p in params]
    ref = torch.optim.AdamW(ref_params, lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999))
    for _ in range(2):
        for p in ref_params:
            p.grad = torch.randn(p.shape, generator=g) if p.requires_grad else None
        ref.step()
    sd = ref.state_dict()
    assert sorted(sd["state"].keys()) == trainable

    opt.load_state_dict(sd)                                      # torch -> fused
    assert opt.step_count == 2
    for i, o, n, shape in opt._slices:
        assert torch.equal(opt.exp_avg[o:o + n].view(shape), sd["state"][i]["exp_avg"])
        assert torch.equal(opt.exp_avg_sq[o:o + n].view(shape), sd["state"][i]["exp_avg_sq"])
        assert torch.equal(opt.master[o:o + n].view(shape), params[i].detach().float())   # no master in a torch checkpoint

    out = opt.state_dict()                                       # fused -> torch
    assert sorted(out["state"].keys()) == trainable and out["param_groups"][0]["params"] == list(range(len(params)))
    ref2 = torch.optim.AdamW([torch.nn.Parameter(p.detach().float().clone(), requires_grad=p.requires_grad) for p in params], lr=1.0)
    ref2.load_state_dict(out)
    assert ref2.param_groups[0]["lr"] == 3e-5 and ref2.param_groups[0]["weight_decay"] == 0.1
    for i in trainable:
        st = ref2.state[ref2.param_groups[0]["params"][i]]
        assert float(st["step"]) == 2.0 and torch.equal(st["exp_avg"], sd["state"][i]["exp_avg"])

    # fused -> fused round trip keeps the fp32 master weights (they differ from the bf16 parameters after real steps)
    opt.master.add_(1e-4)
    snap = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state_dict()["state"][trainable[0]].items()}
    opt2_model = build_drop_in_model(spec, "internvl2-tiny-optim", device="cpu")
    opt2 = FusedAdamW(list(opt2_model.parameters()), ParamStore(opt2_model, "", spec, bucket_bytes=8 << 20, allow_cpu=True), lr=1.0)
    opt2.load_state_dict(opt.state_dict())
    assert opt2.step_count == 2 and torch.equal(opt2.master, opt.master) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)
    assert torch.equal(opt2.store.flat_param, opt.master.bfloat16()) and opt2.param_groups[0]["lr"] == 3e-5
    assert torch.equal(snap["master"], opt2.state_dict()["state"][trainable[0]]["master"])

    # an optimizer whose store no longer backs the parameters (the model was re-flattened behind it) must say so
    opt.check_attached()
    ParamStore(model, "", spec, bucket_bytes=8 << 20, allow_cpu=True)          # what a re-load does: a new flat store takes the parameters
    with pytest.raises(RuntimeError, match="no longer live"):
        opt.check_attached()

    with pytest.raises(ValueError, match="does not match"):
        bad = ref.state_dict()
        bad["param_groups"][0]["params"] = bad["param_groups"][0]["params"][:-1]
        opt.load_state_dict(bad)


def test_every_entry_point_survives_null_arguments():
    """C-ABI robustness: each declared function called with all-zero arguments must return (an error code, or 0 for an
    empty job) instead of dereferencing; run in a child process so a crash is a test failure, not a dead test run"""
    code = r'''
import ctypes, os, re
root = {root!r}
header = open(os.path.join(root, "include", "simlingo_b200.h")).read()
so = ctypes.CDLL(os.path.join(root, "simlingo_b200", "libsimlingo_b200.so"))
names = sorted(set(re.findall(r"\b(slb_[a-z0-9_]+)\s*\(", header)) - {{"slb_last_error", "slb_version", "slb_num_sms", "slb_comm_version"}})
zeros = [ctypes.c_void_p(0)] * 16
bad = []
for n in names:
    f = getattr(so, n); f.restype = ctypes.c_int32
    rc = f(*zeros)
    if rc > 0: bad.append((n, rc))
print("OK", len(names), bad)
'''.format(root=ROOT)
    from simlingo_b200 import build
    build.build()
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK") and r.stdout.strip().endswith("[]"), (r.returncode, r.stdout[-300:], r.stderr[-300:])


def test_flop_model_reproduces_survey_figures():
    """``roofline.achieved`` and the model-TFLOP/s figures of bench.py are computed from simlingo_b200.spec's FLOP model;
    it has to reproduce the ALGORITHMIC figures the scope table states: ViT 1.447 TF per frame (attention
    0.207), projector 0.0046, LLM pass 0.422 / 0.446 / 0.459 TF at L = 545 / 575 / 591, LM head 0.272 GF per row,
    frame 1.898 TF, offline config 121.5 TF, training step 42.4 TF per GPU"""
    from simlingo_b200 import spec as S
    sp = S.INTERNVL2_1B
    tf = lambda x: x / 1e12
    assert tf(S.flops_vit(sp, 2)) == pytest.approx(1.447, abs=2e-3)
    assert tf(2 * sp.vit_layers * sp.vit_heads * 4 * sp.vit_tokens ** 2 * 64) == pytest.approx(0.207, abs=1e-3)
    assert tf(S.flops_proj(sp, 2)) == pytest.approx(0.0046, abs=2e-4)
    for L, want in ((545, 0.422), (575, 0.446), (591, 0.459)):
        assert tf(S.flops_llm(sp, L)) == pytest.approx(want, abs=1.5e-3), L
    assert 2 * sp.llm_hidden * sp.vocab / 1e9 == pytest.approx(0.272, abs=1e-3)
    assert tf(S.flops_frame(sp, 575)) == pytest.approx(1.898, abs=3e-3)
    assert tf(64 * S.flops_frame(sp, 575)) == pytest.approx(121.5, abs=0.2)
    sys.path.insert(0, ROOT)
    import bench
    assert tf(bench.train_flops(sp, 8, 591)) == pytest.approx(42.4, abs=0.3)


def test_bench_contract_on_a_cpu_only_host():
    """bench.py without a GPU: the product arm must refuse (no CPU fallback), the reference arm must print exactly ONE
    JSON line carrying the contract's keys, with the oracle as a bounded sample on the host cores"""
    import json
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    bench = os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, bench, "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in (r.stdout + r.stderr)
    r = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "vla_forward_frames_per_s" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the default run carries the other BASELINE configurations as sub-objects (training step, agent step, language mode)
    for sub, metric, unit in (("train", "vla_train_samples_per_s", "samples/s"), ("agent", "agent_step_latency_ms_p50", "ms"),
                              ("language", "language_generated_tokens_per_s", "tokens/s")):
        assert d[sub]["metric"] == metric and d[sub]["unit"] == unit and d[sub]["value"] > 0 and d[sub]["cpu_baseline"]["kind"] == "port", sub
        assert "workload" in d[sub]["config"]
    assert d["agent"]["higher_is_better"] is False


def test_zero_checkpoint_directory_round_trip(tmp_path):
    """train.py:104-111: a DeepSpeed ZeRO-2 checkpoint DIRECTORY (flat fp32 partitions per rank, padded to 2 * world; frozen
    parameters as fragments; tied tensors through shared_params; Lightning's ``_forward_module.`` prefix) or a single file is
    turned into the state_dict ``DrivingModel.load_state_dict`` takes.  Layout per deepspeed/utils/zero_to_fp32.py (0.16.2)."""
    from simlingo_b200.checkpoint import load_checkpoint, load_zero_checkpoint, write_zero2_checkpoint
    from simlingo_b200.spec import tiny_spec
    from tests.helpers import build_drop_in_model
    spec = tiny_spec(1, 1, 512)
    model = build_drop_in_model(spec, "internvl2-tiny-ckpt", device="cpu")
    want = model.state_dict()
    for world in (1, 3, 8):
        d = str(tmp_path / f"zero_w{world}")
        write_zero2_checkpoint(model, d, world)
        files = sorted(os.listdir(os.path.join(d, "checkpoint")))
        assert len([f for f in files if f.endswith("_optim_states.pt")]) == world and open(os.path.join(d, "latest")).read() == "checkpoint"
        part = torch.load(os.path.join(d, "checkpoint", "zero_pp_rank_0_mp_rank_00_optim_states.pt"), weights_only=False)
        n = part["optimizer_state_dict"]["single_partition_of_fp32_groups"][0].numel()
        assert (n * world) % (2 * world) == 0 and n * world >= sum(p.numel() for p in model.parameters() if p.requires_grad)
        got = load_zero_checkpoint(d)
        assert set(got) == set(want), set(got) ^ set(want)
        for k, v in want.items():
            assert got[k].dtype == torch.float32 and torch.equal(got[k], v.float()), k
        # the aliases the reference's module tree produces stay tied (llm.py:91, adaptors.py:227-229)
        assert got["adaptors.language.lm_head.weight"].data_ptr() == got["language_model.model.base_model.model.lm_head.weight"].data_ptr()
        fresh = build_drop_in_model(spec, "internvl2-tiny-ckpt", seed=7, device="cpu")
        fresh.load_state_dict(load_checkpoint(d), strict=True)
        assert all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), want.values()))
    one = str(tmp_path / "single.pt")
    torch.save({"state_dict": {"_forward_module." + k: v for k, v in want.items()}, "epoch": 3}, one)
    got = load_checkpoint(one)
    assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)
    # damaged directories fail loudly
    os.remove(os.path.join(str(tmp_path / "zero_w3"), "checkpoint", "zero_pp_rank_2_mp_rank_00_optim_states.pt"))
    with pytest.raises(ValueError, match="Expected 3"):
        load_zero_checkpoint(str(tmp_path / "zero_w3"))
    os.remove(os.path.join(str(tmp_path / "zero_w1"), "latest"))
    with pytest.raises(ValueError, match="latest"):
        load_zero_checkpoint(str(tmp_path / "zero_w1"))


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in lib.py / postprocess.py must have the layout the C compiler gives the header's structs: compile a
    probe against include/simlingo_b200.h with gcc and compare sizeof and every field offset."""
    import ctypes as C
    from simlingo_b200 import lib
    from simlingo_b200.postprocess import ControlParams
    pairs = [("slb_gemm_args", lib.GemmArgs), ("slb_decode_args", lib.DecodeArgs), ("slb_wgrad_problem", lib.WgradProblem),
             ("slb_heads_weights", lib.HeadsWeights), ("slb_wp_weights", lib.WpWeights), ("slb_control_params", ControlParams)]
    header = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "simlingo_b200.h")).read(), flags=re.S)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "simlingo_b200.h"', 'int main(void) {']
    expect = []
    for cname, cls in pairs:
        body = re.search(r"typedef struct\s*\{([^}]*)\}\s*" + cname + r"\s*;", header, re.S)
        assert body, cname
        text = body.group(1)
        cfields = []
        for decl in text.split(";"):
            for part in decl.split(","):
                m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())
                if m and part.strip():
                    cfields.append(m.group(1))
        pyfields = [f[0] for f in cls._fields_]
        assert len(cfields) == len(pyfields), (cname, cfields, pyfields)
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for cf in cfields:
            lines.append(f'  printf(" %zu", offsetof({cname}, {cf}));')
        lines.append('  printf("\\n");')
        expect.append((cname, [C.sizeof(cls)] + [getattr(cls, f).offset for f in pyfields]))
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    got = {ln.split()[0]: [int(v) for v in ln.split()[1:]] for ln in out}
    for cname, vals in expect:
        assert got[cname] == vals, (cname, got[cname], vals)
