import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
L = lib.load()
for (M, N, K, bn) in [(131200, 4096, 1024, 2256), (131200, 1024, 4096, 2256), (36800, 9728, 896, 2256), (36800, 896, 896, 2224), (131200, 3072, 1024, 2256)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    kw = dict(swiglu=True) if N == 9728 else {}
    out = lib.gemm(a, b, block_n=bn, **kw)
    for _ in range(2): lib.gemm(a, b, out=out, block_n=bn, **kw)
    buf = torch.zeros(8 * 256 * 2, dtype=torch.int64, device="cuda")
    L.slb_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.gemm(a, b, out=out, block_n=bn, **kw); e1.record(); torch.cuda.synchronize()
    L.slb_debug_set_trace(None)
    d = buf[:8].tolist()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"{M}x{N}x{K} bn={bn}: {t*1e6:.0f} us {2.0*M*N*K/t/1e12:.0f} TF | MMA total {d[0]} wait_full {d[1]} ({100*d[1]/max(d[0],1):.0f}%) wait_tempty {d[2]} ({100*d[2]/max(d[0],1):.0f}%) | EPI total {d[3]} wait_tfull {d[4]} ({100*d[4]/max(d[3],1):.0f}%) | PROD total {d[5]} wait_empty {d[6]} ({100*d[6]/max(d[5],1):.0f}%) | clk {d[0]/t/1e9:.2f} GHz")
