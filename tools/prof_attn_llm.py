"""LLM causal GQA attention forward at the offline64 shape (B=64, L=575) / training shape (B=8, L=591)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simlingo_b200 import lib
lib.load()
B, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 575)
Hq, Hkv = 14, 2
lmax = (L + 127) // 128 * 128
qkv = torch.randn(B * L, (Hq + 2 * Hkv) * 64, device="cuda").to(torch.bfloat16)
kc = torch.zeros(B, Hkv, lmax, 64, device="cuda", dtype=torch.bfloat16); vc = torch.zeros_like(kc)
lib.rope_kv_write(qkv, kc, vc, B, L, 0)
out = torch.empty(B * L, Hq * 64, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, Hq, L, device="cuda")
for use_lse in (False, True):
    f = lambda: lib.attn_gqa(qkv, qkv.stride(0), kc, vc, B, L, 0, out=out, lse=lse if use_lse else None)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    fl = B * Hq * 4 * L * L * 64 * 0.5
    print(f"attn_gqa_fwd B={B} L={L} lse={use_lse}: {us:.1f} us {fl / us / 1e6:.1f} TFLOP/s")
