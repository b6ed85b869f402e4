import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
L = lib.load()
tiles = 128
qkv = torch.randn(tiles * 1025, 3072, device="cuda").to(torch.bfloat16)
out = lib.attn_vit(qkv, tiles, 1025)
for _ in range(2): lib.attn_vit(qkv, tiles, 1025, out=out)
buf = torch.zeros(8 * 256 * 2, dtype=torch.int64, device="cuda")
L.slb_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
lib.attn_vit(qkv, tiles, 1025, out=out)
torch.cuda.synchronize()
L.slb_debug_set_trace(None)
b = buf.cpu().view(8, 256, 2)
ev = []
for role in range(8):
    for i in range(256):
        tag, t = int(b[role, i, 0]), int(b[role, i, 1])
        if t: ev.append((t, role, tag))
ev.sort()
t0 = ev[0][0]
names = {0: "MMA0", 1: "SMw4", 2: "KPRD", 3: "SMw5", 4: "SMw6", 5: "SMw7"}
for t, role, tag in [e for e in ev if e[0] - t0 > 11000][:110]:
    print(f"{t - t0:8d} {names[role]} {tag}")
