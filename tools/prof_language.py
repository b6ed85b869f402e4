"""Kernel-time breakdown of one language-mode batch: python tools/prof_language.py [B] [G]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from simlingo_b200 import spec as S
from torch.profiler import ProfilerActivity, profile

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
G = int(sys.argv[2]) if len(sys.argv) > 2 else 64
spec = S.INTERNVL2_1B
dev = torch.device("cuda", 0)
model = Bn.build_model(spec, dev)
eng = model._engine()
hb = Bn.host_agent_batch(spec, B, 500, None)
ids, fr, vd = hb["ids"].to(dev), hb["frames"].to(dev), hb["valid"].to(dev)
run = lambda: eng.driving_forward(fr, ids, vd, hb["placeholders"], max_new_tokens=G, eos_token_id=None, ids_cpu=hb["ids"])
for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f"B={B} G={G}: total device time {tot / 1e3:.2f} ms")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:22]:
    print(f"{e.self_device_time_total / 1e3:9.3f} ms {100 * e.self_device_time_total / tot:5.1f}% n={e.count:5d} avg={e.self_device_time_total / e.count:8.1f} us  {e.key[:100]}")
