import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
lib.load()
M, N, K = [int(x) for x in sys.argv[1:4]]
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = lib.gemm(a, b, block_n=bn)
for _ in range(3): lib.gemm(a, b, out=out, block_n=bn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lib.gemm(a, b, out=out, block_n=bn); e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3
print(f"gemm {M}x{N}x{K} bn={bn}: {t*1e6:.1f} us {2.0*M*N*K/t/1e12:.1f} TF")
