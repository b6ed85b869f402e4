#!/usr/bin/env python
"""Benchmark of the SimLingo (InternVL2-1B) VLA hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|offline64|train|agent|language] [--impl ours|reference]

The default (``--workload all``) prints ONE JSON line whose top-level keys are the offline forward (below) and which carries the
other BASELINE configurations as sub-objects measured in the same run, each with its own device-timed value, end-to-end value,
CPU-oracle baseline and parity-vs-oracle figures: ``train`` (configs[3]: training step incl. the NCCL gradient all-reduce at N > 1,
its exposed time and an in-bench data-parallel check), and at N = 1 ``agent`` (configs[1]) and ``language`` (configs[4]).

Headline workload (BASELINE.json configs[2], the configuration the frames/s metric is quoted on):
"offline batched forward": 64 synthetic frames per GPU per step = 128 InternViT tiles -> pixel-shuffle/mlp1 ->
prompt assembly (L=545) + 30 driving queries -> one teacher-forced Qwen2 pass (L+30=575) -> route / speed
waypoint heads.  Batch-sharded data parallel: every rank processes its own 64 frames, no data-path collective
("scaling": "weak").  One JSON line is printed by rank 0 (contract in the task statement):

  value     frames/s with inputs resident in HBM (whole job, all ranks; max-over-ranks device time)
  e2e       the same step from pinned HOST inputs: H2D of the uint8 camera frames + prompt ids, GPU pre-processing
            (resize / tile / normalise), the drop-in ``DrivingModel.forward_model`` + heads, D2H of the predicted
            waypoints - all inside the timed region
  roofline  tcgen05 GEMM kernel: algorithmic FLOPs of all its launches in a step / their summed CUDA-event time,
            against the measured cuBLAS bf16 peak (sustained figure: the kernel runs inside a long step)
  cpu_baseline  the fp32 oracle (the reference's PyTorch path restated, oracle/model.py) on the host cores,
            bounded sample of the same workload (1 frame)

``--impl reference`` times the oracle alone (rank 0 only) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from simlingo_b200 import spec as S  # noqa: E402

FRAMES_PER_GPU = 64
# (0.275+0.760 + 0.540+0.230 + 0.279+1.030 + 1.640+0.258) GB / 4 launches, measured once with `ncu --set full` (profiles/)
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = 1.526e9   # profiles/r02_ncu_gemm2_offline64_summary.txt (ncu --set full, fp32 residual streams)
PROMPT_LEN = 545


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else that writes to file descriptor 1 (constructor banners, NCCL's
    version line printed from C) is sent to stderr; emit() writes the line to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _pin(t):
    """pinned host memory for the H2D legs (plain memory on a box without a driver: the reference arm runs there too)"""
    return t.pin_memory() if torch.cuda.is_available() else t


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", 1394.0)), hbm=float(d.get("hbm_gbs", 6545.9)), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc = gpu_index, None

    def start(self):
        if os.environ.get("SLB_BENCH_NO_SAMPLER") == "1":   # diagnostic switch: is the nvidia-smi poll itself perturbing the run?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], 0.0, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# model + weights: the same deterministic synthetic weights (spec.init_state_dict, "planted walk" LM head) on the GPU
# and in the CPU oracle, so that every workload can state its parity against the oracle in the JSON line
# --------------------------------------------------------------------------------------------------
_SD32 = None


def oracle_state_dict(spec):
    """fp32 copy of the synthetic weights for the CPU oracle (every value is bf16-representable), built once per process."""
    global _SD32
    if _SD32 is None:
        _SD32 = S.init_state_dict(spec, seed=0, with_aliases=True)
    return _SD32


def build_model(spec, device):
    """The drop-in DrivingModel built exactly as the agent does (bf16 default dtype, then .to(cuda)), with the deterministic
    synthetic weights of the InternVL2-1B architecture (no checkpoint offline): LoRA B non-zero, layer scales in [0.05, 0.2]."""
    import contextlib
    from simlingo_b200.modules import register_variant
    from simlingo_training.models.driving import DrivingModel
    from tests.helpers import StubTokenizer
    name = "OpenGVLab/InternVL2-1B"
    register_variant(name, spec)
    cfg = dict(
        vision_model=dict(_target_="simlingo_training.models.encoder.vlm.VLMEncoderModel", variant=name, embed_dim=512, freeze=False),
        language_model=dict(_target_="simlingo_training.models.language_model.llm.LLM", variant=name, lora=True, lora_alpha=64,
                            lora_r=32, lora_dropout=0.1),
        lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999), pct_start=0.05, speed_wps_mode="2d", predict_route_as_wps=True,
    )
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with contextlib.redirect_stdout(sys.stderr):   # the reference's constructors print banners; stdout carries the JSON line only
            model = DrivingModel(cfg_data_module={"use_global_img": False}, processor=StubTokenizer(spec), cache_dir=None, **cfg)
    finally:
        torch.set_default_dtype(prev)
    model.load_state_dict(oracle_state_dict(spec), strict=True)   # fp32 -> bf16 parameter copy, exact
    return model.to(device).eval()


def release(*objs):
    import gc
    del objs
    gc.collect()
    torch.cuda.empty_cache()


class Ctx:
    def __init__(self, rank, world, local):
        self.rank, self.world, self.local = rank, world, local
        self.device = torch.device("cuda", local)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms: float) -> float:
        if self.world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], device=self.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_region(self, fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize, CUDA events on the launch stream, MAX over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trace = os.environ.get("SLB_BENCH_TRACE") == "1"   # diagnostic: host-side wall time of every step (adds a sync per step)
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            if trace:
                t1 = time.perf_counter()
                torch.cuda.synchronize()
                print(f"[rank {self.rank}] step: issue {1e3 * (t1 - t0):.1f} ms, done {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr, flush=True)
        e1.record()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1))


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def host_batch(spec, batch, seed):
    """Pinned host tensors of one step's inputs: `batch` DISTINCT frames and prompts."""
    ids = S.synth_prompt_ids(spec, batch, seed)
    frames = S.synth_frames(spec, batch, seed, dtype=torch.bfloat16)
    valid = torch.ones_like(ids, dtype=torch.bool)
    ph = S.synth_placeholders(spec, batch, seed)
    return dict(ids=_pin(ids), frames=_pin(frames), valid=_pin(valid), placeholders=ph)


def make_example(hb, device):
    from simlingo_training.utils.custom_types import DrivingInput, LanguageLabel
    ids = hb["ids"].to(device, non_blocking=True)
    valid = hb["valid"].to(device, non_blocking=True)
    frames = hb["frames"].to(device, non_blocking=True)
    lab = LanguageLabel(ids, valid, valid, hb["placeholders"], [""] * ids.shape[0], torch.zeros_like(valid))
    z = torch.zeros((ids.shape[0], 1), device=device)
    return DrivingInput(frames, z, z, z, z, z, lab, lab)


@torch.no_grad()
def offline_step(model, example):
    """Teacher-forced forward of a batch: DrivingModel.forward_model + driving heads (BASELINE config 3)."""
    ad = model.adaptors(example)
    feats, _ = model.forward_model(example, ad, want_logits=False)
    drv = model.adaptors.split_outputs_by_adaptor(ad, feats)["driving"]
    pred = model.adaptors.driving.get_predictions(drv)
    return pred["route"], pred["speed_wps"]


def _gemm_kernel_name(M, N, K, kw):
    """Which kernel slb_gemm_bf16 dispatches this problem to (mirror of the rule in csrc/gemm.cu / gemv.cu)"""
    kmajor = not kw.get("a_t") and not kw.get("b_t")
    bn = kw.get("block_n", 0)
    if bn in (2256, 2224, 2192):
        return f"gemm2_bf16_kernel<{bn - 2000}>"
    if bn == 0 and kmajor and kw.get("aux") is None:
        if M <= 4:
            return "gemv_bf16_kernel"
        if M <= 32 and K % 32 == 0:
            return "skinny_gemm_kernel"
    if bn == 0:   # tile shape by estimated waves x tile area / kernel efficiency (slb_gemm_bf16)
        if kmajor and N <= 64 and M >= 1024:
            return "gemm_bf16_kernel<narrow>"
        cd = lambda a, b: -(-a // b)
        sms, mt, mt2 = 148, cd(M, 128), cd(M, 256)
        cost = lambda tiles, units, area, eff: cd(tiles, units) * area / eff
        best, name = cost(mt * cd(N, 256), sms, 128 * 256, 1.0), "gemm_bf16_kernel<1-CTA>"
        if not kw.get("swiglu"):
            best = min(best, cost(mt * cd(N, 128), sms, 128 * 128, 0.9))
        min_m2 = 1024 if (kw.get("a_t") and kw.get("b_t") and M % 256 == 0) else 2048
        if M >= min_m2 and not (kw.get("a_t") and not kw.get("b_t")):
            c = cost(mt2 * cd(N, 256), sms // 2, 128 * 256, 1.1)
            if c <= best:
                best, name = c, "gemm2_bf16_kernel<256>"
            if kmajor and not kw.get("swiglu") and N % 224 == 0:
                c = cost(mt2 * (N // 224), sms // 2, 128 * 224, 1.08)
                if c < best:
                    best, name = c, "gemm2_bf16_kernel<224>"
        return name
    return "gemm_bf16_kernel<1-CTA>"


def gemm_roofline(peaks, step_fn, traffic=None):
    """Times every launch of the tcgen05 GEMM inside one step with CUDA events (on the launch stream).  `roofline` describes the
    DOMINANT kernel (the one with the largest share of the step's GEMM time: gemm2_bf16_kernel<256>, the cta_group::2 256 x 256
    tile kernel): achieved = sum of 2MNK over its launches / sum of their durations; `all_gemm_launches` keeps the aggregate over every
    slb_gemm_bf16 call of the step (narrow LoRA down-projections, 1-CTA tiles, weight-streaming kernels included)."""
    from simlingo_b200 import lib
    orig = lib.gemm
    rec = []

    def timed(a, b, out=None, **kw):
        M, K = (a.shape[1], a.shape[0]) if kw.get("a_t") else (a.shape[0], a.shape[1])
        N = b.shape[1] if kw.get("b_t") else b.shape[0]
        if kw.get("a2") is not None:
            K += kw["a2"].shape[1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(a, b, out, **kw)
        e1.record()
        rec.append((2.0 * M * N * K, e0, e1, _gemm_kernel_name(M, N, K, kw)))
        return r

    lib.gemm = timed
    try:
        step_fn()
        torch.cuda.synchronize()
    finally:
        lib.gemm = orig
    per = {}
    for f, a, b, name in rec:
        t = per.setdefault(name, [0.0, 0.0, 0])
        t[0] += f
        t[1] += a.elapsed_time(b) * 1e-3
        t[2] += 1
    flops = sum(t[0] for t in per.values())
    secs = sum(t[1] for t in per.values())
    dom = max(per, key=lambda k: per[k][1])
    dflops, dsecs, dn = per[dom]
    ach = dflops / dsecs / 1e12
    return {"bound": "tensor", "kernel": dom + " (tcgen05, TMA-fed, TMEM accumulators)", "achieved": round(ach, 1), "peak": peaks["tflops"],
            "unit": "TFLOP/s", "frac": round(ach / peaks["tflops"], 4), "traffic": traffic,
            "traffic_note": None if traffic is None else
            "dram read+write bytes per launch, mean over the 4 GEMMs of one InternViT layer at M=131200 (fc2, qkv, proj, fc1) from "
            "profiles/r02_ncu_gemm2_offline64_summary.txt (ncu --set full); algorithmic bytes of the same 4 launches with the fp32 residual "
            "stream (proj / fc2 read and write it): 1.479e9 mean",
            "launches_per_step": dn, "gemm_flops_per_step": dflops, "gemm_ms_per_step": round(dsecs * 1e3, 3),
            "peak_source": peaks["src"] + " (sustained cuBLAS bf16)",
            "all_gemm_launches": {"launches_per_step": len(rec), "flops_per_step": flops, "ms_per_step": round(secs * 1e3, 3),
                                  "achieved": round(flops / secs / 1e12, 1), "frac": round(flops / secs / 1e12 / peaks["tflops"], 4),
                                  "per_kernel": {k: {"launches": v[2], "ms": round(v[1] * 1e3, 3), "tflops": round(v[0] / v[1] / 1e12, 1)} for k, v in per.items()}}}


# --------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline of each workload and the --impl reference arm): bounded samples of the same workloads,
# returning the oracle's outputs as well so the GPU arm can state its parity on the same inputs
# --------------------------------------------------------------------------------------------------
def _threads():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return cores


def cpu_offline(spec, hb, reps: int = 2, warmup: int = 1):
    """fp32 oracle on the host cores: teacher-forced forward of frame 0 of `hb` -> (frames/s, cores, seconds, route, speed)."""
    from oracle import model as O
    cores = _threads()
    sd = oracle_state_dict(spec)
    ids, valid = hb["ids"][:1], hb["valid"][:1]
    fr, ph = hb["frames"][:1].float(), hb["placeholders"][:1]
    ts = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            ad = O.adaptor_list_forward(sd, spec, ids, valid, torch.zeros_like(valid))
            ad, feats, _ = O.forward_model(sd, spec, ad, fr, ph, logits=False)
            _, drv = O.split_outputs(ad, feats)
            pred = O.driving_predictions(sd, spec, drv)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return 1.0 / t, cores, t, pred["route"], pred["speed_wps"]


def cpu_train(spec, hb):
    """one sample (sample 0 of `hb`) forward + backward of the fp32 oracle under torch autograd, eval mode (no dropout)"""
    from oracle import model as O
    cores = _threads()
    sd = {k: v.detach().clone().requires_grad_(S.trainable(k)) for k, v in S.init_state_dict(spec, seed=0).items()}
    t0 = time.perf_counter()
    loss, _, _ = O.forward_loss(sd, spec, hb["frames"][:1].float(), hb["ids"][:1], hb["valid"][:1], hb["loss_masking"][:1], hb["placeholders"][:1],
                                hb["wps"][:1], hb["path"][:1], training=False)
    loss.backward()
    t = time.perf_counter() - t0
    return 1.0 / t, cores, t, float(loss.detach())


def cpu_agent(spec, hb, max_new_tokens=100, eos=True):
    """one DrivingModel.forward of the oracle in the reference's own formulation (no KV cache: every token re-forwards the sequence)"""
    from oracle import model as O
    cores = _threads()
    sd = oracle_state_dict(spec)
    t0 = time.perf_counter()
    with torch.no_grad():
        sp, rt, toks = O.driving_forward(sd, spec, hb["frames"][:1].float(), hb["ids"][:1], hb["valid"][:1], hb["placeholders"][:1],
                                         max_new_tokens=max_new_tokens, eos_token_id=spec.eos_id if eos else None)
    return time.perf_counter() - t0, cores, sp, rt, toks[0]


# --------------------------------------------------------------------------------------------------
# BASELINE configs[2]: offline batched forward (the headline workload)
# --------------------------------------------------------------------------------------------------
def offline_cfg(args, world):
    return {"workload": f"offline batched forward: {args.frames} distinct frames/GPU/step (2x448^2 tiles each), prompt L={PROMPT_LEN}+30 queries, "
                        "teacher-forced Qwen2 pass + route/speed heads (BASELINE configs[2])",
            "frames_per_gpu": args.frames, "parallelism": f"dp{world} batch-sharded, no data-path collective",
            "l2": "inputs and activations (>= 1 GB per step) far exceed the 126 MB L2; no explicit flush needed"}


def ref_offline(args):
    spec = S.INTERNVL2_1B
    steps = max(1, min(args.steps, 3))
    hb = host_batch(spec, 1, 1234)
    fps, cores, t, _, _ = cpu_offline(spec, hb, steps, 1)
    return {"impl": "reference", "metric": "vla_forward_frames_per_s", "value": round(fps, 4), "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": round(t * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": offline_cfg(args, args.gpus),
            "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "1 frame per step (2 tiles + 575-token Qwen2 pass), fp32 oracle, median"},
            "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_offline(args, ctx):
    spec = S.INTERNVL2_1B
    rank, world, device = ctx.rank, ctx.world, ctx.device
    model = build_model(spec, device)
    eng = model._engine()
    B = args.frames
    hb = host_batch(spec, B, 1234 + rank)
    example = make_example(hb, device)
    torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        offline_step(model, example)
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ms = ctx.timed_region(lambda: offline_step(model, example), args.steps)
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)
    route, speed = offline_step(model, example)
    route, speed = route.float().cpu(), speed.float().cpu()

    # ---- end to end: pinned host inputs -> H2D -> pre-processing -> step -> D2H of the predictions ----
    # As a deployment would run it: every step's inputs are the uint8 camera frames (359 x 1024 after the agent's crop) and
    # the prompt ids in pinned host memory; they are copied to the device, resized / tiled / normalised there (slb_preprocess_frames) and
    # pushed through DrivingModel.forward_model + heads; the predicted waypoints return to pinned host memory.
    import numpy as np
    from simlingo_b200.preprocess import preprocess_frames
    from simlingo_training.utils.custom_types import DrivingInput, LanguageLabel
    out_host = (torch.empty((B, 20, 2), dtype=torch.float32).pin_memory(), torch.empty((B, 10, 2), dtype=torch.float32).pin_memory())
    cam_host = _pin(torch.from_numpy(np.stack([S.synth_camera(359, 1024, 77 + 1000 * rank + b) for b in range(B)])))   # B distinct camera frames

    def e2e_step():
        cam = cam_host.to(device, non_blocking=True)
        ids, valid = hb["ids"].to(device, non_blocking=True), hb["valid"].to(device, non_blocking=True)
        frames = preprocess_frames(cam).view(B, 1, 2, 3, 448, 448)
        lab = LanguageLabel(ids, valid, valid, hb["placeholders"], [""] * B, torch.zeros_like(valid))
        z = torch.zeros((B, 1), device=device)
        r, s = offline_step(model, DrivingInput(frames, z, z, z, z, z, lab, lab))
        out_host[0].copy_(r.float(), non_blocking=True)
        out_host[1].copy_(s.float(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    ms_e2e = ctx.timed_region(e2e_step, args.steps)
    e2e_val = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = cam_host.numel() + hb["ids"].numel() * 8 + hb["valid"].numel()
    d2h = (out_host[0].numel() + out_host[1].numel()) * 4

    line = None
    if rank == 0:
        peaks = measured_peaks()
        roof = gemm_roofline(peaks, lambda: offline_step(model, example), traffic=NCU_GEMM_DRAM_BYTES_PER_LAUNCH if args.frames == FRAMES_PER_GPU else None)
        flops_step = B * S.flops_frame(spec, PROMPT_LEN + 30)
        line = {"metric": "vla_forward_frames_per_s", "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": offline_cfg(args, world),
                "e2e": {"value": round(e2e_val, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(ms_e2e / args.steps, 3)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof,
                "model_tflops_per_gpu": round(flops_step * args.steps / (ms * 1e-3) / 1e12, 1),
                "algorithmic_tflop_per_step_per_gpu": round(flops_step / 1e12, 2),
                "checksum": {"what": "rank 0's device-resident step: sums over the 64 predicted routes [64,20,2] / speed waypoints [64,10,2]",
                             "route_sum": round(float(route.double().sum()), 4), "speed_wps_sum": round(float(speed.double().sum()), 4),
                             "route_abs_max": round(float(route.abs().max()), 4), "distinct_frames": int(torch.unique(route.reshape(B, -1), dim=0).shape[0])}}
        if not args.no_cpu_baseline and world == 1:
            fps, cores, t, rt_ref, sp_ref = cpu_offline(spec, hb)
            line["cpu_baseline"] = {"value": round(fps, 4), "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "frame 0 of the same batch (2 tiles + 575-token Qwen2 pass), fp32 oracle, median of 2 after 1 warm-up"}
            line["parity_vs_oracle"] = {"what": "frame 0 of the timed batch (computed inside the B=64 step) vs the fp32 CPU oracle on the same weights / inputs; "
                                                "max-abs error over the tensor's largest magnitude, tolerance 2e-2",
                                        "route_relerr": round(relerr(route[:1], rt_ref), 5), "speed_wps_relerr": round(relerr(speed[:1], sp_ref), 5)}
    release(model, eng, example)
    return line


# --------------------------------------------------------------------------------------------------
# BASELINE configs[3]: training step (full ViT + LoRA Qwen2, fused AdamW), batch 8 / GPU, DP gradient all-reduce
# --------------------------------------------------------------------------------------------------
TRAIN_BATCH = 8
ANSWER_LEN = 16


def host_train_batch(spec, batch, seed):
    ids = S.synth_prompt_ids(spec, batch, seed, answer_len=ANSWER_LEN)
    g = torch.Generator().manual_seed(seed + 9)
    frames = torch.randn((batch, 1, spec.tiles_per_frame, 3, spec.image, spec.image), generator=g).clamp_(-2.2, 2.7).to(torch.bfloat16)
    valid = torch.ones_like(ids, dtype=torch.bool)
    lm = torch.zeros_like(valid)
    lm[:, -ANSWER_LEN:] = True
    wps, path = S.synth_labels(spec, batch, seed)
    return dict(ids=_pin(ids), frames=_pin(frames), valid=_pin(valid), loss_masking=_pin(lm),
                wps=_pin(wps), path=_pin(path), placeholders=S.synth_placeholders(spec, batch, seed))


def make_train_example(hb, device, rows=None):
    from simlingo_training.utils.custom_types import DrivingExample, DrivingInput, DrivingLabel, LanguageLabel
    sl = (lambda t: t) if rows is None else (lambda t: t[rows])
    mv = lambda t: sl(t).to(device, non_blocking=True)
    ids, valid = mv(hb["ids"]), mv(hb["valid"])
    ph = hb["placeholders"] if rows is None else hb["placeholders"][rows]
    lab = LanguageLabel(ids, valid, valid, ph, [""] * ids.shape[0], mv(hb["loss_masking"]))
    z = torch.zeros((ids.shape[0], 1), device=device)
    di = DrivingInput(mv(hb["frames"]), z, z, z, z, z, lab, lab)
    return DrivingExample(di, DrivingLabel(mv(hb["wps"]), mv(hb["path"]), lab, z), ["x"] * ids.shape[0])


def train_flops(spec, B, L):
    """Algorithmic FLOPs of one training step per GPU: ViT and projector x3 (fprop + dgrad + wgrad), frozen
    Qwen2 base x2 (fprop + dgrad), LoRA x3, LM head + CE on the ANSWER_LEN labelled rows per sample x2."""
    vit = S.flops_vit(spec, 2 * B) + S.flops_proj(spec, 2 * B)
    base = S.flops_llm(spec, L, B, lora=False)
    lora = S.flops_llm(spec, L, B, lora=True) - base
    head = 2.0 * B * ANSWER_LEN * spec.llm_hidden * spec.vocab
    return 3 * vit + 2 * base + 3 * lora + 2 * head


def train_cfg(args, world):
    B = args.batch
    L = PROMPT_LEN + ANSWER_LEN
    return {"workload": f"training step: batch {B}/GPU, 2x448^2 tiles/frame, L={L}+30, full ViT + mlp1 + LoRA(r32, dropout 0.1) Qwen2 fwd+bwd, "
                        "global-norm clip 0.3 + fused AdamW, bf16 gradient all-reduce overlapped with backward (BASELINE configs[3])",
            "batch_per_gpu": B, "parallelism": f"dp{world} batch-sharded, NCCL all-reduce of the flat bf16 gradient buffer",
            "l2": "activations + gradients (> 20 GB per step) far exceed the 126 MB L2; no explicit flush needed"}


def ref_train(args):
    spec = S.INTERNVL2_1B
    hb = host_train_batch(spec, 1, 4321)
    sps, cores, t, _ = cpu_train(spec, hb)
    return {"impl": "reference", "metric": "vla_train_samples_per_s", "value": round(sps, 5), "unit": "samples/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": round(t * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": train_cfg(args, args.gpus),
            "cpu_baseline": {"value": round(sps, 5), "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "1 sample fwd+bwd (fp32 oracle under torch autograd), single run"},
            "e2e": {"value": round(sps, 5), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_train(args, ctx):
    spec = S.INTERNVL2_1B
    rank, world, device = ctx.rank, ctx.world, ctx.device
    B = args.batch
    L = PROMPT_LEN + ANSWER_LEN
    model = build_model(spec, device)
    model.train()  # LoRA dropout active, as the reference trains
    from simlingo_b200.optim import FusedAdamW
    store = model.param_store()
    if world > 1:
        import torch.distributed as dist
        store.enable_data_parallel()
        dist.broadcast(store.flat_param, src=0)
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], store, lr=3e-5, weight_decay=0.1, betas=(0.9, 0.999), max_grad_norm=0.3)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=3e-5, total_steps=10000, pct_start=0.05)
    teng = model.__dict__["_slb_train_engine"]
    hb = host_train_batch(spec, B, 4321 + rank)
    example = make_train_example(hb, device)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step(ex):
        opt.zero_grad()
        out = model.training_step(ex)
        out["loss"].backward()
        opt.step()
        sched.step()
        return out["loss"]

    import warnings
    warnings.filterwarnings("ignore", message=".*lr_scheduler.step.*")
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(example)

    # ---- data-parallel check (N > 1): the SAME batch on every rank => the exchanged gradient of every bucket must be
    # world x the local one (bf16 sums of identical values are exact; what remains is the run-to-run noise of the fp32
    # atomics in the small-parameter reductions).  A bucket reduced twice / not at all would read 2x / (1/world)x. ----
    dp_check = None
    if world > 1:
        model.eval()    # no dropout: both passes see the same masks (none)
        ex_same = make_train_example(host_train_batch(spec, B, 999), device)
        opt.zero_grad()
        with store.no_sync():
            model.training_step(ex_same)["loss"].backward()
        torch.cuda.synchronize()
        local = store.flat_grad.float().clone()
        opt.zero_grad()
        model.training_step(ex_same)["loss"].backward()
        store.wait_exchange()
        torch.cuda.synchronize()
        red = store.flat_grad.float()
        worst = 0.0
        for a, b, _ in store._buckets:
            den = float(local[a:b].abs().max()) * world
            worst = max(worst, float((red[a:b] - world * local[a:b]).abs().max()) / max(den, 1e-30))
        t = torch.tensor([worst], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dp_check = {"what": "same batch on all ranks: max over buckets and ranks of |reduced - world*local| / (world*max|local|)",
                    "buckets": len(store._buckets), "max_relerr": round(float(t.item()), 6), "ok": bool(t.item() < 2e-2)}
        opt.zero_grad()
        model.train()
        del local, red

    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    l0, n_ar0 = teng.launches + opt.launches, store.n_allreduce
    ms = ctx.timed_region(lambda: step(example), args.steps)
    launches = teng.launches + opt.launches - l0
    n_ar = (store.n_allreduce - n_ar0) // max(1, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- exposed all-reduce time: the same steps with the exchange switched off (gradients stay local: timing only) ----
    exposed = None
    if world > 1:
        def step_nosync():
            with store.no_sync():
                step(example)
        step_nosync()
        ms_ns = ctx.timed_region(step_nosync, args.steps)
        exposed = round((ms - ms_ns) / args.steps, 3)
        dist.broadcast(store.flat_param, src=0)   # ranks drifted apart during the unsynchronised steps
        opt.resync_master()

    def e2e_step():
        ex = make_train_example(hb, device)
        loss = step(ex)
        loss_host.copy_(loss.detach().float(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    ms_e2e = ctx.timed_region(e2e_step, args.steps)
    e2e_val = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = sum(hb[k].numel() * hb[k].element_size() for k in ("frames", "ids", "valid", "loss_masking", "wps", "path"))
    dp_backend = store.dp_backend if world > 1 else None
    line = None
    if rank == 0:
        peaks = measured_peaks()
        teng.graphs_enabled = False  # the per-launch event timing needs the eager launches (the timed region above replays graphs)
        store.pg, store.world = None, 1  # rank 0 alone from here on: no collectives in the roofline pass
        store.layout_version += 1
        step(example)
        roof = gemm_roofline(peaks, lambda: step(example))
        teng.graphs_enabled = True
        flops_step = train_flops(spec, B, L + 30)
        line = {"metric": "vla_train_samples_per_s", "value": round(value, 2), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": train_cfg(args, world),
                "e2e": {"value": round(e2e_val, 2), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": round(ms_e2e / args.steps, 3)},
                "gpu_launches": launches, "cuda_graph_replays": teng.graph_replays, "clocks": clocks, "roofline": roof, "loss": round(float(loss_host), 4),
                "n_allreduce": n_ar, "allreduce_bytes_per_step": store.numel * 2 if world > 1 else 0, "allreduce_exposed_ms": exposed, "dp_backend": dp_backend, "dp_check": dp_check,
                "model_tflops_per_gpu": round(flops_step * args.steps / (ms * 1e-3) / 1e12, 1),
                "model_flops_frac_of_peak": round(flops_step * args.steps / (ms * 1e-3) / 1e12 / peaks["tflops"], 4),
                "algorithmic_tflop_per_step_per_gpu": round(flops_step / 1e12, 2)}
        if not args.no_cpu_baseline and world == 1:
            sps, cores, t, loss_ref = cpu_train(spec, hb)
            line["cpu_baseline"] = {"value": round(sps, 5), "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": "sample 0 of the same batch, fwd+bwd (fp32 oracle under torch autograd), single run"}
            model.load_state_dict(oracle_state_dict(spec), strict=True)   # back to the initial weights (in place: the store stays)
            model.eval()
            with torch.no_grad():
                out, _ = model.forward_loss(make_train_example(hb, device, rows=slice(0, 1)))
            line["parity_vs_oracle"] = {"what": "forward_loss of sample 0 at the initial weights (eval mode: LoRA dropout off) vs the fp32 CPU oracle, tolerance 2e-2",
                                        "loss": round(float(out.loss), 4), "loss_oracle": round(loss_ref, 4),
                                        "loss_relerr": round(abs(float(out.loss) - loss_ref) / abs(loss_ref), 5)}
    if store.comm is not None:
        store.comm.destroy()
    release(model, opt, teng, store, example)
    return line


# --------------------------------------------------------------------------------------------------
# BASELINE configs[1] (closed-loop agent step, batch 1, p50 latency) and configs[4] (language mode: prefill + 64 greedy
# tokens, batch 32).  The "planted walk" synthetic weights (spec.init_state_dict) control the number of greedy tokens
# before EOS through the last prompt token.
# --------------------------------------------------------------------------------------------------
def host_agent_batch(spec, batch, seed, n_gen):
    ids = S.synth_prompt_ids(spec, batch, seed)
    if n_gen is not None:
        ids[:, -1] = (spec.eos_id - n_gen * S.LMHEAD_SHIFT) % spec.vocab   # greedy decoding emits n_gen tokens, the last one EOS
    frames = S.synth_frames(spec, batch, seed, dtype=torch.bfloat16)
    valid = torch.ones_like(ids, dtype=torch.bool)
    return dict(ids=_pin(ids), frames=_pin(frames), valid=_pin(valid), placeholders=S.synth_placeholders(spec, batch, seed))


def agent_cfg(world):
    return {"workload": "closed-loop agent step: DrivingModel.forward, batch 1, prompt L=545, greedy decode G tokens (KV cache) + 30-query pass "
                        "+ heads; latency percentiles over the timed steps (BASELINE configs[1])",
            "parallelism": f"replicas only ({world} independent agents)", "l2": "weights (1.9 GB bf16) exceed the 126 MB L2"}


def ref_agent(args):
    spec = S.INTERNVL2_1B
    hb = host_agent_batch(spec, 1, 99, 1)
    t, cores, *_ = cpu_agent(spec, hb)
    t *= 1e3
    return {"impl": "reference", "metric": "agent_step_latency_ms_p50", "value": round(t, 1), "unit": "ms", "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": round(t, 1), "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": agent_cfg(args.gpus),
            "cpu_baseline": {"value": round(t, 1), "unit": "ms", "cores": cores, "kind": "port", "sample": "one agent step with G=1, fp32 oracle (no KV cache, as the reference)"},
            "e2e": {"value": round(t, 1), "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_agent(args, ctx):
    """One LingoAgent.run_step model call (reference team_code/agent_simlingo.py:796-797): DrivingModel.forward on one
    frame = ViT (2 tiles) + projector + Qwen2 prefill + greedy decode until EOS + 30-query pass + heads."""
    spec = S.INTERNVL2_1B
    rank, world, device = ctx.rank, ctx.world, ctx.device
    model = build_model(spec, device)
    from simlingo_b200 import lib
    steps = max(args.steps, 20)
    out = {}
    keep = {}
    for G in (1, 25):
        hb = host_agent_batch(spec, 1, 99 + rank, G)
        ex = make_example(hb, device)
        for _ in range(max(args.warmup, 3)):
            sp, rt, lang = model(ex)
        assert len(model.sampled_tokens[0]) == G, (G, len(model.sampled_tokens[0]))
        torch.cuda.synchronize()
        keep[G] = (hb, sp.float().cpu(), rt.float().cpu(), model.sampled_tokens[0].cpu().tolist())
        dev_ms, e2e_ms = [], []
        eng = model._engine()
        l0 = eng.launches   # kernels of the library launched by the engine, those inside replayed CUDA graphs included
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sp, rt, lang = model(ex)
            e1.record()
            torch.cuda.synchronize()
            dev_ms.append(e0.elapsed_time(e1))
        launches = (eng.launches - l0) // steps
        # end to end as the agent sees it (agent_simlingo.py:470-502, 796-797, 878): the cropped uint8 camera frame (359 x 1024)
        # leaves pinned host memory, is resized / tiled / normalised on the GPU (slb_preprocess_frames), goes through
        # DrivingModel.forward, and the predictions are turned into (steer, throttle, brake) by control_pid: geometry on the
        # GPU (slb_control_inputs), one 64-byte row back to the host, PID windows on the host
        from simlingo_b200.postprocess import ControlPID
        from simlingo_b200.preprocess import preprocess_frames
        cam_host = _pin(torch.from_numpy(S.synth_camera(359, 1024, 3 + rank))[None])
        speed_host = _pin(torch.tensor([4.0]))
        pid = ControlPID()
        for _ in range(steps + 3):
            t0 = time.perf_counter()
            ex2 = make_example(hb, device)
            cam = preprocess_frames(cam_host.to(device, non_blocking=True)).view(1, 1, 2, 3, 448, 448)
            ex2 = ex2._replace(camera_images=cam)
            sp, rt, lang = model(ex2)
            steer, throttle, brake = pid.control_pid(rt, speed_host, sp)   # ends with the read-back + stream synchronize
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        e2e_ms = e2e_ms[3:]
        q = lambda v, p: sorted(v)[min(len(v) - 1, int(p * len(v)))]
        out[G] = dict(p50=round(q(dev_ms, 0.5), 3), p90=round(q(dev_ms, 0.9), 3), e2e_p50=round(q(e2e_ms, 0.5), 3), e2e_p90=round(q(e2e_ms, 0.9), 3),
                      launches=launches, tflops=round((S.flops_frame(spec, PROMPT_LEN + 30) + G * 2 * 0.494e9 + (G + 1) * 2 * spec.llm_hidden * spec.vocab) / 1e12, 3))
    line = None
    if rank == 0:
        hb = keep[1][0]
        h2d = 3 * 359 * 1024 + hb["ids"].numel() * 8 + hb["valid"].numel() + 4
        line = {"metric": "agent_step_latency_ms_p50", "value": out[1]["p50"], "unit": "ms", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
                "ms_per_step": out[1]["p50"], "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": agent_cfg(world), "e2e": {"value": out[1]["e2e_p50"], "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 64,
                                                   "ends_at": "steer / throttle / brake (control_pid)"},
                "gpu_launches": out[1]["launches"] * steps, "latency": {"G=1": out[1], "G=25": out[25]},
                "realtime_budget_ms": 50.0}
        if not args.no_cpu_baseline and world == 1:
            t, cores, sp_ref, rt_ref, tok_ref = cpu_agent(spec, keep[1][0])
            line["cpu_baseline"] = {"value": round(t * 1e3, 1), "unit": "ms", "cores": cores, "kind": "port",
                                    "sample": "the same agent step with G=1, fp32 oracle (no KV cache, as the reference), single run"}
            line["parity_vs_oracle"] = {"what": "the G=1 agent step vs the fp32 CPU oracle: greedy tokens identical, waypoints / route max rel err (tolerance 2e-2)",
                                        "tokens_identical": keep[1][3] == tok_ref.tolist(), "speed_wps_relerr": round(relerr(keep[1][1], sp_ref), 5),
                                        "route_relerr": round(relerr(keep[1][2], rt_ref), 5)}
    release(model)
    return line


def language_cfg(args, world):
    B, G = args.lang_batch, 64
    return {"workload": f"language mode: batch {B}/GPU, ViT (2 tiles/frame) + prefill L=545 + {G} greedy tokens (KV cache, EOS suppressed) "
                        "+ 30-query pass (BASELINE configs[4])", "batch_per_gpu": B, "parallelism": f"dp{world} batch-sharded, no data-path collective",
            "l2": "weights + KV cache exceed the 126 MB L2"}


LANG_ORACLE_TOKENS = 4


def ref_language(args):
    # bounded sample of the same workload: 1 sample, 4 greedy tokens, in the reference's own formulation (no KV cache:
    # every token re-forwards the whole sequence, llm.py:217-235)
    spec = S.INTERNVL2_1B
    hb = host_agent_batch(spec, 1, 500, None)
    t, cores, *_ = cpu_agent(spec, hb, max_new_tokens=LANG_ORACLE_TOKENS, eos=False)
    tps = LANG_ORACLE_TOKENS / t
    return {"impl": "reference", "metric": "language_generated_tokens_per_s", "value": round(tps, 4), "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": round(t * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": language_cfg(args, args.gpus),
            "cpu_baseline": {"value": round(tps, 4), "unit": "tokens/s", "cores": cores, "kind": "port",
                             "sample": "1 sample, ViT + 4 greedy tokens (no KV cache, as the reference) + 30-query pass, fp32 oracle, single run"},
            "e2e": {"value": round(tps, 4), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_language(args, ctx):
    """Language mode (eval.py -> predict_step): ViT + prefill + 64 greedy tokens with EOS suppressed, batch 32 / GPU."""
    spec = S.INTERNVL2_1B
    rank, world, device = ctx.rank, ctx.world, ctx.device
    B, G = args.lang_batch, 64
    model = build_model(spec, device)
    eng = model._engine()
    from simlingo_b200 import lib
    hb = host_agent_batch(spec, B, 500 + rank, None)

    def step(from_host):
        ids = hb["ids"].to(device, non_blocking=True)
        fr = hb["frames"].to(device, non_blocking=True)
        vd = hb["valid"].to(device, non_blocking=True)
        sp, rt, toks = eng.driving_forward(fr, ids, vd, hb["placeholders"], max_new_tokens=G, eos_token_id=None, ids_cpu=hb["ids"])
        return torch.stack(toks).cpu() if from_host else toks

    for _ in range(max(1, args.warmup)):
        step(False)
    l0 = eng.launches   # kernels of the library launched by the engine, those inside replayed CUDA graphs included
    ms = ctx.timed_region(lambda: step(False), args.steps)
    launches = eng.launches - l0
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        toks = step(True)
    ctx.barrier()
    ms2 = ctx.max_ms((time.perf_counter() - t0) * 1e3)
    line = None
    if rank == 0:
        val = world * B * G * args.steps / (ms * 1e-3)
        val2 = world * B * G * args.steps / (ms2 * 1e-3)
        h2d = hb["frames"].numel() * 2 + hb["ids"].numel() * 8 + hb["valid"].numel()
        line = {"metric": "language_generated_tokens_per_s", "value": round(val, 1), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(1, args.warmup), "ms_per_step": round(ms / args.steps, 2), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": language_cfg(args, world),
                "e2e": {"value": round(val2, 1), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * G * 8,
                        "ms_per_step": round(ms2 / args.steps, 2)},
                "gpu_launches": launches, "samples_per_s": round(world * B * args.steps / (ms * 1e-3), 2)}
        if not args.no_cpu_baseline and world == 1:
            one = {k: (v[:1] if torch.is_tensor(v) else v[:1]) for k, v in hb.items()}
            t, cores, _, _, tok_ref = cpu_agent(spec, one, max_new_tokens=LANG_ORACLE_TOKENS, eos=False)
            line["cpu_baseline"] = {"value": round(LANG_ORACLE_TOKENS / t, 4), "unit": "tokens/s", "cores": cores, "kind": "port",
                                    "sample": f"sample 0 of the same batch, ViT + {LANG_ORACLE_TOKENS} greedy tokens (no KV cache, as the reference) + 30-query pass, fp32 oracle, single run"}
            line["parity_vs_oracle"] = {"what": f"first {LANG_ORACLE_TOKENS} greedy tokens of sample 0 (decoded inside the batch of {B}) vs the fp32 CPU oracle",
                                        "tokens_identical": toks[0, :LANG_ORACLE_TOKENS].tolist() == tok_ref.tolist()}
    release(model, eng)
    return line


# --------------------------------------------------------------------------------------------------
SUB_KEYS = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "higher_is_better", "dtype", "config", "e2e", "gpu_launches", "roofline",
            "cpu_baseline", "parity_vs_oracle", "loss", "n_allreduce", "allreduce_bytes_per_step", "allreduce_exposed_ms", "dp_backend", "dp_check",
            "model_tflops_per_gpu", "model_flops_frac_of_peak", "algorithmic_tflop_per_step_per_gpu", "cuda_graph_replays", "latency",
            "realtime_budget_ms", "samples_per_s", "clocks", "impl")


def as_sub(line):
    return None if line is None else {k: line[k] for k in SUB_KEYS if k in line}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "offline64", "train", "agent", "language"])
    ap.add_argument("--lang-batch", type=int, default=32, help="samples per GPU (--workload language)")
    ap.add_argument("--batch", type=int, default=TRAIN_BATCH, help="training samples per GPU per step (--workload train)")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload

    if args.impl == "reference":
        # the reference's own CPU implementation of the path (oracle port: the reference stack cannot be installed offline,
        # DESIGN.md section 2) on the host cores; rank 0 alone runs and prints, the other ranks exit 0 without work
        if rank != 0:
            return
        if wl == "all":
            line = ref_offline(args)
            line["train"], line["agent"], line["language"] = as_sub(ref_train(args)), as_sub(ref_agent(args)), as_sub(ref_language(args))
        else:
            line = {"offline64": ref_offline, "train": ref_train, "agent": ref_agent, "language": ref_language}[wl](args)
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    ctx = Ctx(rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=ctx.device)
    t_start = time.perf_counter()
    if wl == "all":
        line = run_offline(args, ctx)
        train = run_train(args, ctx)
        agent = language = None
        if world == 1:   # batch-1 agent step: replicas only; language mode: same collective-free sharding as the headline
            agent = run_agent(args, ctx)
            language = run_language(args, ctx)
        if rank == 0:
            line["train"], line["agent"], line["language"] = as_sub(train), as_sub(agent), as_sub(language)
    else:
        line = {"offline64": run_offline, "train": run_train, "agent": run_agent, "language": run_language}[wl](args, ctx)
    if rank == 0:
        line["bench_wall_s"] = round(time.perf_counter() - t_start, 1)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
