"""Micro-benchmark of the backward kernels at the training-step shapes (B=8: 16 ViT tiles, 8 x 591 LLM tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simlingo_b200 import lib

lib.load()
dev = "cuda"


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def rnd(*s):
    return torch.randn(*s, device=dev).to(torch.bfloat16)


T, N, H = 16, 1025, 16
qkv, dout = rnd(T * N, 3 * H * 64), rnd(T * N, H * 64)
lse = torch.empty(T, H, N, device=dev)
out = lib.attn_vit(qkv, T, N, H, lse=lse)
delta = lib.attn_delta(out, dout, T, N, H)
us = timeit(lambda: lib.attn_vit_bwd(qkv, dout, lse, delta, T, N, H))
fl = T * H * 4 * N * N * 64 * 2.5
print(f"attn_vit_bwd T={T}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
us = timeit(lambda: lib.attn_vit(qkv, T, N, H, lse=lse, out=out))
print(f"attn_vit_fwd T={T}: {us:8.1f} us  {fl / 2.5 / us / 1e6:7.1f} TFLOP/s")

B, L, Hq, Hkv, lmax = 8, 591, 14, 2, 640
qkv2, dout2 = rnd(B * L, (Hq + 2 * Hkv) * 64), rnd(B * L, Hq * 64)
kc = torch.zeros(B, Hkv, lmax, 64, device=dev, dtype=torch.bfloat16); vc = torch.zeros_like(kc)
lib.rope_kv_write(qkv2, kc, vc, B, L, 0)
lse2 = torch.empty(B, Hq, L, device=dev)
out2 = lib.attn_gqa(qkv2, qkv2.stride(0), kc, vc, B, L, 0, lse=lse2)
delta2 = lib.attn_delta(out2, dout2, B, L, Hq)
us = timeit(lambda: lib.attn_gqa_bwd(qkv2, qkv2.stride(0), kc, vc, dout2, lse2, delta2, B, L))
fl2 = B * Hq * 4 * L * L * 64 * 0.5 * 2.5
print(f"attn_gqa_bwd B={B} L={L}: {us:8.1f} us  {fl2 / us / 1e6:7.1f} TFLOP/s")
us = timeit(lambda: lib.attn_gqa(qkv2, qkv2.stride(0), kc, vc, B, L, 0, lse=lse2, out=out2))
print(f"attn_gqa_fwd B={B} L={L}: {us:8.1f} us  {fl2 / 2.5 / us / 1e6:7.1f} TFLOP/s")

rows, cols = T * N, 1024
x, w, b, dy = rnd(rows, cols), rnd(cols), rnd(cols), rnd(rows, cols)
mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
lib.layernorm(x, w, b, 1e-6, stats=(mean, rstd))
dw, db = torch.zeros(cols, device=dev), torch.zeros(cols, device=dev)
dx = torch.empty_like(x)
us = timeit(lambda: lib.layernorm_bwd(dy, x, w, mean, rstd, dw, db, dx=dx))
print(f"layernorm_bwd {rows}x{cols}: {us:8.1f} us  {rows * cols * 2 * 3 / us / 1e3:7.1f} GB/s")
us = timeit(lambda: lib.layernorm(x, w, b, 1e-6, out=dx, stats=(mean, rstd)))
print(f"layernorm_fwd {rows}x{cols}: {us:8.1f} us  {rows * cols * 2 * 2 / us / 1e3:7.1f} GB/s")
acc = torch.zeros(cols, device=dev)
us = timeit(lambda: lib.col_reduce(dy, acc, x))
print(f"col_reduce(a*b) {rows}x{cols}: {us:8.1f} us  {rows * cols * 2 * 2 / us / 1e3:7.1f} GB/s")
f = rnd(rows, 4096); df = rnd(rows, 4096)
us = timeit(lambda: lib.gelu_bwd(f, df, out=df))
print(f"gelu_bwd {rows}x4096: {us:8.1f} us  {rows * 4096 * 2 * 3 / us / 1e3:7.1f} GB/s")
# LoRA-shaped GEMMs (M = 4728)
M = B * L
for (m, n, k, at, bt, what) in [(M, 32, 896, 0, 0, "t = xd A^T"), (M, 4864, 32, 0, 0, "y += t B^T (gate)"), (4864, 32, M, 1, 1, "dB = dy^T t"),
                                 (M, 32, 4864, 0, 1, "dt = dy B"), (32, 896, M, 1, 1, "dA = dt^T xd"), (M, 896, 32, 0, 1, "dx += dt A"),
                                 (M, 4864, 896, 0, 0, "gate base"), (M, 896, 4864, 0, 1, "dgrad gate")]:
    a = rnd(k, m) if at else rnd(m, k)
    bb = rnd(k, n) if bt else rnd(n, k)
    o = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: lib.gemm(a, bb, out=o, a_t=bool(at), b_t=bool(bt)))
    print(f"gemm {what:20s} M={m} N={n} K={k}: {us:7.1f} us {2 * m * n * k / us / 1e6:7.1f} TF/s")
