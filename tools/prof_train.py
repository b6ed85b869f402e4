"""Kernel-time breakdown of one training step (torch.profiler / CUPTI): python tools/prof_train.py [batch] [workload]"""
import contextlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from simlingo_b200 import spec as S
from torch.profiler import ProfilerActivity, profile

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
spec = S.INTERNVL2_1B
dev = torch.device("cuda", 0)
with contextlib.redirect_stdout(sys.stderr):
    model = Bn.build_model(spec, dev)
if mode == "train":
    model.train()
    from simlingo_b200.optim import FusedAdamW
    store = model.param_store()
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], store, lr=3e-5, weight_decay=0.1, max_grad_norm=0.3)
    ex = Bn.make_train_example(Bn.host_train_batch(spec, B, 1), dev)

    def step():
        opt.zero_grad()
        out = model.training_step(ex)
        out["loss"].backward()
        opt.step()
else:
    ex = Bn.make_example(Bn.host_batch(spec, B, 1), dev)

    def step():
        Bn.offline_step(model, ex)
for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(3):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"cpu issue time/step {(t1 - t0) / 3 * 1e3:.1f} ms; wall/step {(t2 - t0) / 3 * 1e3:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f"total device time {tot / 1e3:.2f} ms")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:45]:
    print(f"{e.self_device_time_total / 1e3:9.3f} ms {100 * e.self_device_time_total / tot:5.1f}% n={e.count:5d} avg={e.self_device_time_total / e.count:8.1f} us  {e.key[:110]}")

# ---- host-side profile of one step (python overhead) ----
import cProfile, pstats, io
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
step()
pr.disable()
torch.cuda.synchronize()
sio = io.StringIO()
pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(45)
print("\n".join(l[:170] for l in sio.getvalue().splitlines()[:75]))
