import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
lib.load()
M = 131200
for (N, K, what) in [(1024, 1024, "proj+ls+res"), (1024, 4096, "fc2+ls+res"), (896, 4864, "down+res (M=36800)")]:
    m = 36800 if "down" in what else M
    a = torch.randn(m, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda").to(torch.bfloat16); ls = torch.rand(N, device="cuda").to(torch.bfloat16)
    x = torch.randn(m, N, device="cuda").to(torch.bfloat16)
    f = lambda: lib.gemm(a, w, out=x, bias=bias, scale_n=ls, residual=x)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    print(f"{what:22s} {us:8.1f} us  {2*m*N*K/us/1e6:7.1f} TF/s")
