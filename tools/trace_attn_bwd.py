import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib
L = lib.load()
T, N, H = 16, 1025, 16
qkv = torch.randn(T * N, 3 * H * 64, device="cuda").to(torch.bfloat16)
dout = torch.randn(T * N, H * 64, device="cuda").to(torch.bfloat16)
lse = torch.empty(T, H, N, device="cuda")
out = lib.attn_vit(qkv, T, N, H, lse=lse)
delta = lib.attn_delta(out, dout, T, N, H)
for _ in range(2):
    lib.attn_vit_bwd(qkv, dout, lse, delta, T, N, H)
buf = torch.zeros(8 * 256 * 2, dtype=torch.int64, device="cuda")
L.slb_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
lib.attn_vit_bwd(qkv, dout, lse, delta, T, N, H)
torch.cuda.synchronize()
L.slb_debug_set_trace(None)
b = buf.cpu().view(8, 256, 2)
ev = []
for role in range(3):
    for i in range(256):
        tag, t = int(b[role, i, 0]), int(b[role, i, 1])
        if t:
            ev.append((t, role, tag))
ev.sort()
t0 = ev[0][0]
names = {0: "MMA ", 1: "SM_A", 2: "SM_B"}
mma_tags = {0: "wait p_ready", 1: "got p_ready", 2: "sdp(it+1) issued; wait dq_free", 3: "got dq_free -> issue dV dK dQ"}
sm_tags = {0: "wait sdp_full", 1: "got sdp_full", 2: "chunk0 math done; wait dq_full(it-1)", 3: "got dq_full(it-1)", 4: "p_ready arrived", 5: "dq_out done"}
for t, role, tag in ev:
    it = (tag - 10) // (4 if role == 0 else 6)
    k = (tag - 10) % (4 if role == 0 else 6)
    print(f"{t - t0:8d} {names[role]} it={it} {(mma_tags if role == 0 else sm_tags)[k]}")
