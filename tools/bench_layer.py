"""Same-box micro-benchmarks of one InternViT layer / one Qwen2 layer at the offline-64 shapes, each kernel with its real
epilogue, next to the library bar (torch.matmul = cuBLASLt, F.scaled_dot_product_attention, flash-attn when importable).
CUDA events on the launch stream, L2 flushed between timed iterations, median of 9.
    python tools/bench_layer.py > gpurun_out/bench_layer.log"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib  # noqa: E402

lib.load()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=9, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us


bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)
ONLY = set(sys.argv[1:])


def row(name, us, flops=None, note=""):
    tf = f"{flops / us / 1e6:8.1f} TF/s" if flops else " " * 13
    print(f"{name:46s} {us:9.1f} us {tf} {note}", flush=True)


if not ONLY or "vit" in ONLY:
    M, D, I = 131200, 1024, 4096
    h, att, f = bf(M, D), bf(M, D), bf(M, I)
    wqkv, bqkv, wproj, bproj, ls = bf(3 * D, D), bf(3 * D), bf(D, D), bf(D), bf(D)
    w1, b1, w2, b2 = bf(I, D), bf(I), bf(D, I), bf(D)
    x32 = torch.randn(M, D, device=dev)
    x16 = x32.to(torch.bfloat16)
    qkv, fo = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16), torch.empty(M, I, device=dev, dtype=torch.bfloat16)
    row("vit.qkv  +bias", timeit(lambda: lib.gemm(h, wqkv, out=qkv, bias=bqkv)), 2.0 * M * 3 * D * D)
    row("  torch.matmul", timeit(lambda: torch.matmul(h, wqkv.t())), 2.0 * M * 3 * D * D)
    row("vit.proj +bias*ls+res fp32 stream", timeit(lambda: lib.gemm(att, wproj, out=x32, bias=bproj, scale_n=ls, residual=x32, out_fp32=True)), 2.0 * M * D * D)
    row("vit.proj +bias*ls+res bf16 stream", timeit(lambda: lib.gemm(att, wproj, out=x16, bias=bproj, scale_n=ls, residual=x16)), 2.0 * M * D * D)
    row("vit.proj plain (no epilogue terms)", timeit(lambda: lib.gemm(att, wproj, out=x16)), 2.0 * M * D * D)
    row("  torch.matmul", timeit(lambda: torch.matmul(att, wproj.t())), 2.0 * M * D * D)
    row("vit.fc1  +bias+GELU", timeit(lambda: lib.gemm(h, w1, out=fo, bias=b1, act=lib.ACT_GELU)), 2.0 * M * I * D)
    row("vit.fc1  +bias (no activation)", timeit(lambda: lib.gemm(h, w1, out=fo, bias=b1)), 2.0 * M * I * D)
    row("vit.fc1  +bias+ReLU", timeit(lambda: lib.gemm(h, w1, out=fo, bias=b1, act=lib.ACT_RELU)), 2.0 * M * I * D)
    row("  torch.matmul", timeit(lambda: torch.matmul(h, w1.t())), 2.0 * M * I * D)
    row("  torch F.gelu(F.linear)", timeit(lambda: F.gelu(F.linear(h, w1, b1))), 2.0 * M * I * D)
    row("vit.fc2  +bias*ls+res fp32 stream", timeit(lambda: lib.gemm(f, w2, out=x32, bias=b2, scale_n=ls, residual=x32, out_fp32=True)), 2.0 * M * D * I)
    row("vit.fc2  +bias*ls+res bf16 stream", timeit(lambda: lib.gemm(f, w2, out=x16, bias=b2, scale_n=ls, residual=x16)), 2.0 * M * D * I)
    row("  torch.matmul", timeit(lambda: torch.matmul(f, w2.t())), 2.0 * M * D * I)
    lnw, lnb = bf(D), bf(D)
    row("layernorm fp32 in -> bf16", timeit(lambda: lib.layernorm(x32, lnw, lnb, 1e-6, out=h)), note=f"{M * D * 6 / 1e3:.0f} KB")
    row("layernorm bf16 in -> bf16", timeit(lambda: lib.layernorm(x16, lnw, lnb, 1e-6, out=h)))
    del h, att, f, x32, x16, qkv, fo

if not ONLY or "attn" in ONLY:
    tiles = 128
    qkv = bf(tiles * 1025, 3072)
    out = lib.attn_vit(qkv, tiles, 1025)
    fl = tiles * 16 * 4 * 1025 * 1025 * 64
    row("attn_vit2 + cls (128 tiles)", timeit(lambda: lib.attn_vit(qkv, tiles, 1025, out=out)), fl)
    q, k, v = qkv.view(tiles, 1025, 3, 16, 64).permute(2, 0, 3, 1, 4)
    row("  torch SDPA", timeit(lambda: F.scaled_dot_product_attention(q, k, v)), fl)
    try:
        from flash_attn import flash_attn_qkvpacked_func
        q5 = qkv.view(tiles, 1025, 3, 16, 64)
        row("  flash-attn 2 (qkvpacked)", timeit(lambda: flash_attn_qkvpacked_func(q5, causal=False)), fl)
    except Exception as e:  # noqa: BLE001
        print("  flash-attn unavailable:", repr(e)[:80])
    B, L, Hq, Hkv = 64, 575, 14, 2
    qq = bf(B * L, 1152)
    kc, vc = bf(B, Hkv, 640, 64), bf(B, Hkv, 640, 64)
    o2 = lib.attn_gqa(qq, 1152, kc, vc, B, L, 0)
    fl = B * Hq * 4 * L * L * 64 * 0.5
    row("attn_gqa2 causal GQA (B=64, L=575)", timeit(lambda: lib.attn_gqa(qq, 1152, kc, vc, B, L, 0, out=o2)), fl, "(causal-half FLOPs)")
    q4 = qq.view(B, L, 18, 64)[:, :, :14].transpose(1, 2)
    k4, v4 = kc[:, :, :L], vc[:, :, :L]
    row("  torch SDPA (causal, enable_gqa)", timeit(lambda: F.scaled_dot_product_attention(q4, k4, v4, is_causal=True, enable_gqa=True)), fl)
    try:
        from flash_attn import flash_attn_func
        qf, kf, vf = q4.transpose(1, 2).contiguous(), k4.transpose(1, 2).contiguous(), v4.transpose(1, 2).contiguous()
        row("  flash-attn 2 (causal GQA)", timeit(lambda: flash_attn_func(qf, kf, vf, causal=True)), fl)
    except Exception as e:  # noqa: BLE001
        print("  flash-attn unavailable:", repr(e)[:80])

if not ONLY or "llm" in ONLY:
    M, D, I = 36800, 896, 4864
    h, att, act = bf(M, D), bf(M, D), bf(M, I)
    x32 = torch.randn(M, D, device=dev)
    for name, a, n, k in (("t_qkv", h, 96, D), ("t_o", att, 32, D), ("t_gu", h, 64, D), ("t_d", act, 32, I)):
        w = bf(n, k)
        row(f"LoRA down {name} N={n} K={k}", timeit(lambda: lib.gemm(a, w)), 2.0 * M * n * k, f"A read {M * k * 2 / 1e6:.0f} MB")
    wq, t = bf(1152, 992), bf(M, 96)
    row("llm.qkv K=896+96", timeit(lambda: lib.gemm(h, wq, a2=t, bias=bf(1152))), 2.0 * M * 1152 * 992)
    row("llm.qkv K=896 (folded)", timeit(lambda: lib.gemm(h, wq[:, :896].contiguous())), 2.0 * M * 1152 * 896)
    wo, t = bf(896, 928), bf(M, 32)
    row("llm.o K=896+32 fp32 stream", timeit(lambda: lib.gemm(att, wo, a2=t, out=x32, residual=x32, out_fp32=True)), 2.0 * M * 896 * 928)
    wg, t = bf(9728, 960), bf(M, 64)
    row("llm.gate|up SwiGLU K=896+64", timeit(lambda: lib.gemm(h, wg, a2=t, swiglu=True)), 2.0 * M * 9728 * 960)
    wd, t = bf(896, 4896), bf(M, 32)
    row("llm.down K=4864+32 fp32 stream", timeit(lambda: lib.gemm(act, wd, a2=t, out=x32, residual=x32, out_fp32=True)), 2.0 * M * 896 * 4896)

if "lora" in ONLY:
    # training-path LoRA side kernels at the B=8 training shapes (M = 8 * 591)
    M, r = 4728, 32
    sd = torch.zeros(1, device=dev, dtype=torch.int64)
    for name, K, n in (("qkv", 896, 3), ("o", 896, 1), ("gate|up", 896, 2), ("down", 4864, 1)):
        cat = bf(M, K + r * n)
        A = [bf(r, K) for _ in range(n)]
        seeds = list(range(100, 100 + n))
        out = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        us = timeit(lambda: lib.lora_dx(cat, K, A, p=0.1, seeds=seeds, seed_dev=sd, out=out))
        row(f"lora_dx {name} K={K} n={n} (masked)", us, note=f"{(cat.numel() + out.numel()) * 2 / us / 1e3:.0f} GB/s")
        us = timeit(lambda: lib.lora_dx(cat, K, A, out=out))
        row(f"lora_dx {name} K={K} n={n} (no mask)", us)
    x = bf(M, 896)
    row("dropout_multi x3 [4728, 896]", timeit(lambda: lib.dropout_multi(x, 0.1, [1, 2, 3], seed_dev=sd)))
    dy, t = bf(M, 896), bf(M, 32)
    g = torch.zeros(896, 32, device=dev, dtype=torch.bfloat16)
    row("LoRA dB wgrad [896 x 32] K=4728", timeit(lambda: lib.gemm(dy, t, out=g, a_t=True, b_t=True)))
    g2 = torch.zeros(32, 896, device=dev, dtype=torch.bfloat16)
    row("LoRA dA wgrad [32 x 896] K=4728", timeit(lambda: lib.gemm(t, dy, out=g2, a_t=True, b_t=True)))
    wx = bf(896, 4864 + 32)
    dx = bf(M, 896)
    row("llm down dgrad [4728 x 4896] K=896 (b_t)", timeit(lambda: lib.gemm(dx, wx, b_t=True)), 2.0 * M * 4896 * 896)
    wgu = bf(9728, 896 + 64)
    dgu = bf(M, 9728)
    row("llm gate|up dgrad [4728 x 960] K=9728 (b_t)", timeit(lambda: lib.gemm(dgu, wgu, b_t=True)), 2.0 * M * 960 * 9728)

if "decode" in ONLY:
    # batched decode step (language mode, B = 32, ~600 cached positions): weight-streaming kernels, HBM-bound by design
    B, D, I, V, L = 32, 896, 4864, 151655, 608
    x = bf(B, D)
    act = bf(B, I)
    xf = torch.randn(B, D, device=dev)
    for name, w, a, kw in (("qkv   [32 x 1152] K=896 +bias", bf(1152, D), x, dict(bias=bf(1152))),
                           ("o     [32 x 896] K=896 fp32 += ", bf(D, D), x, dict(residual=xf, out=xf, out_fp32=True)),
                           ("gu    [32 x 9728] K=896 SwiGLU", bf(2 * I, D), x, dict(swiglu=True)),
                           ("down  [32 x 896] K=4864 fp32 +=", bf(D, I), act, dict(residual=xf, out=xf, out_fp32=True)),
                           ("lm_head [32 x 151655] K=896 fp32", bf(V, D), x, dict(out_fp32=True))):
        us = timeit(lambda: lib.gemm(a, w, **kw))
        row("skinny " + name, us, note=f"{w.numel() * 2 / us / 1e3:.0f} GB/s of weights")
    kc, vc = bf(B, 2, 640, 64), bf(B, 2, 640, 64)
    qkv = bf(B, 1152)
    att = torch.empty(B, 896, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: lib.attn_gqa(qkv, 1152, kc, vc, B, 1, L, 14, 2, out=att))
    row("attn_decode_group B=32 L=608", us, note=f"{B * 2 * L * 64 * 2 * 2 / us / 1e3:.0f} GB/s of K/V")
    row("rope_kv_write B=32", timeit(lambda: lib.rope_kv_write(qkv, kc, vc, B, 1, L)))
    w = bf(D)
    h = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
    row("rmsnorm fp32 -> bf16 [32 x 896]", timeit(lambda: lib.rmsnorm(xf, w, 1e-6, out=h)))
    lg = torch.randn(B, V, device=dev)
    row("argmax [32 x 151655]", timeit(lambda: lib.argmax(lg)))
    kc1, vc1 = bf(1, 2, 640, 64), bf(1, 2, 640, 64)
    q1 = bf(1, 1152)
    a1 = torch.empty(1, 896, device=dev, dtype=torch.bfloat16)
    row("attn_small B=1 L=608", timeit(lambda: lib.attn_gqa(q1, 1152, kc1, vc1, 1, 1, L, 14, 2, out=a1)))
    x1 = torch.randn(1, D, device=dev)
    for name, w_, kw in (("qkv+norm", bf(1152, D), dict(bias=bf(1152), rms_weight=w, rms_eps=1e-6)), ("gu+norm SwiGLU", bf(2 * I, D), dict(swiglu=True, rms_weight=w, rms_eps=1e-6))):
        row("gemv B=1 " + name, timeit(lambda: lib.gemm(x1, w_, **kw)), note=f"{w_.numel() * 2 / 1e6:.1f} MB")
