"""Kernel-time breakdown of one agent step (G tokens): python tools/prof_agent.py [G]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from simlingo_b200 import spec as S
from torch.profiler import ProfilerActivity, profile

G = int(sys.argv[1]) if len(sys.argv) > 1 else 25
spec = S.INTERNVL2_1B
dev = torch.device("cuda", 0)
model = Bn.build_model(spec, dev)
ex = Bn.make_example(Bn.host_agent_batch(spec, 1, 99, G), dev)
for _ in range(4):
    model(ex)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model(ex)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f"G={G}: total device time {tot / 1e3:.2f} ms, tokens {len(model.sampled_tokens[0])}")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:28]:
    print(f"{e.self_device_time_total / 1e3:9.3f} ms {100 * e.self_device_time_total / tot:5.1f}% n={e.count:5d} avg={e.self_device_time_total / e.count:8.1f} us  {e.key[:100]}")
