"""Micro-benchmarks of the hot kernels on the shape catalogue: TFLOP/s per GEMM shape, attention,
GB/s for the norm kernels.  CUDA events on the launch stream, L2 flushed between timed iterations."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib  # noqa: E402

lib.load()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
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
    return ts[len(ts) // 2] * 1e-3


def bf(*s):
    return torch.randn(*s, device=dev).to(torch.bfloat16)


rows = []
shapes = [
    ("vit.qkv", 131200, 3072, 1024, {}), ("vit.proj", 131200, 1024, 1024, {}), ("vit.fc1", 131200, 4096, 1024, {}),
    ("vit.fc2", 131200, 1024, 4096, {}), ("vit.qkv.agent", 2050, 3072, 1024, {}), ("vit.fc1.agent", 2050, 4096, 1024, {}),
    ("vit.fc2.agent", 2050, 1024, 4096, {}), ("vit.proj.agent", 2050, 1024, 1024, {}),
    ("llm.qkv", 36800, 1152, 896, {}), ("llm.o", 36800, 896, 896, {}), ("llm.gateup", 36800, 9728, 896, {"swiglu": True}),
    ("llm.down", 36800, 896, 4864, {}), ("llm.qkv.agent", 545, 1152, 896, {}), ("llm.gateup.agent", 545, 9728, 896, {"swiglu": True}),
    ("llm.down.agent", 545, 896, 4864, {}), ("lm_head.m1", 1, 151655, 896, {"out_fp32": True}),
    ("lm_head.m32", 32, 151655, 896, {"out_fp32": True}), ("cublas.ref.8192", 8192, 8192, 8192, {}),
]
for name, M, N, K, kw in shapes:
    a, b = bf(M, K), bf(N, K)
    for bn in ([256, 2256] if kw.get("swiglu") else [128, 256, 2256] + ([2224] if N % 224 == 0 else []) + ([2192] if N % 192 == 0 else [])):
        out = lib.gemm(a, b, block_n=bn, **kw)
        t = timeit(lambda: lib.gemm(a, b, out=out, block_n=bn, **kw))
        rows.append(dict(op="gemm", name=name, M=M, N=N, K=K, bn=bn, us=t * 1e6, tflops=2.0 * M * N * K / t / 1e12))
        print(rows[-1], flush=True)
    if not kw:
        t = timeit(lambda: torch.matmul(a, b.t()))
        rows.append(dict(op="torch.matmul", name=name, M=M, N=N, K=K, us=t * 1e6, tflops=2.0 * M * N * K / t / 1e12))
        print(rows[-1], flush=True)
    del a, b

for tiles in (2, 128):
    qkv = bf(tiles * 1025, 3072)
    out = lib.attn_vit(qkv, tiles, 1025)
    t = timeit(lambda: lib.attn_vit(qkv, tiles, 1025, out=out))
    fl = tiles * 16 * 4 * 1025 * 1025 * 64
    rows.append(dict(op="attn_vit", tiles=tiles, us=t * 1e6, tflops=fl / t / 1e12))
    print(rows[-1], flush=True)
    q, k, v = qkv.view(tiles, 1025, 3, 16, 64).permute(2, 0, 3, 1, 4)
    t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    rows.append(dict(op="torch.sdpa(vit)", tiles=tiles, us=t * 1e6, tflops=fl / t / 1e12))
    print(rows[-1], flush=True)

for rowsn, cols in ((131200, 1024), (2050, 1024)):
    x, w, b = bf(rowsn, cols), bf(cols), bf(cols)
    y = lib.layernorm(x, w, b, 1e-6)
    t = timeit(lambda: lib.layernorm(x, w, b, 1e-6, out=y))
    rows.append(dict(op="layernorm", rows=rowsn, cols=cols, us=t * 1e6, gbs=rowsn * cols * 4 / t / 1e9))
    print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_ops.json", "w"), indent=1)
