import os, sys, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
lib.load()
M, N, K = 131200, 4096, 1024
a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = lib.gemm(a, b)
def run(fn, secs=3.0):
    torch.cuda.synchronize(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return 2.0 * M * N * K * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
print("ours sustained 3s:", run(lambda: lib.gemm(a, b, out=out)))
print("cublas sustained 3s:", run(lambda: torch.matmul(a, b.t(), out=out)))
print("ours sustained 3s:", run(lambda: lib.gemm(a, b, out=out)))
bias = torch.randn(N, device="cuda").to(torch.bfloat16)
print("ours bias+gelu sustained:", run(lambda: lib.gemm(a, b, out=out, bias=bias, act=lib.ACT_GELU)))
x = torch.randn(M, 1024, device="cuda").to(torch.bfloat16); w2 = torch.randn(1024, 4096, device="cuda").to(torch.bfloat16); ls = torch.randn(1024, device="cuda").to(torch.bfloat16)
M, N, K = 131200, 1024, 4096
o2 = torch.empty(131200, 1024, device="cuda", dtype=torch.bfloat16)
print("ours fc2 (bias, ls, residual) sustained:", run(lambda: lib.gemm(out, w2, out=x, bias=ls, scale_n=ls, residual=x)))
print("cublas fc2 plain sustained:", run(lambda: torch.matmul(out, w2.t(), out=o2)))
