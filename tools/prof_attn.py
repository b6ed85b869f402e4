import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
lib.load()
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 8
qkv = torch.randn(tiles * 1025, 3072, device="cuda").to(torch.bfloat16)
out = lib.attn_vit(qkv, tiles, 1025)
for _ in range(3):
    lib.attn_vit(qkv, tiles, 1025, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); lib.attn_vit(qkv, tiles, 1025, out=out); b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b) * 1e-3
print("attn_vit tiles", tiles, "us", t * 1e6, "TF", tiles * 16 * 4 * 1025 * 1025 * 64 / t / 1e12)
