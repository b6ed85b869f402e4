import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
lib.load()
tiles, n, H = 1, 128, 16
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(tiles * n, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
dout = torch.randn(tiles * n, H * 64, device="cuda", generator=g).to(torch.bfloat16)
lse = torch.empty(tiles, H, n, device="cuda")
out = lib.attn_vit(qkv, tiles, n, H, lse=lse); torch.cuda.synchronize(); print("fwd ok", lse.abs().max().item(), flush=True)
delta = lib.attn_delta(out, dout, tiles, n, H); torch.cuda.synchronize(); print("delta ok", flush=True)
dq, dk, dv = lib.attn_vit_bwd(qkv, dout, lse, delta, tiles, n, H); torch.cuda.synchronize(); print("bwd ok", flush=True)
