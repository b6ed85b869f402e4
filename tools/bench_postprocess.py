"""Latency of the post-processing call the agent makes every tick (``ControlPID.control_pid``: one kernel, one 64-byte
read-back, host PID) and of the batched ``equal_spacing_route`` used by ``predict_step``.  GPU box only."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import postprocess

g = np.random.default_rng(0)
route = torch.from_numpy(np.cumsum(np.abs(g.normal(1, 0.1, (1, 20, 2))), 1).astype(np.float32)).cuda()
wps = torch.from_numpy(np.cumsum(np.abs(g.normal(1, 0.1, (1, 10, 2))), 1).astype(np.float32)).cuda()
speed = torch.tensor([4.0])
pid = postprocess.ControlPID()
for _ in range(50):
    pid.control_pid(route, speed, wps)
torch.cuda.synchronize()
t = []
for _ in range(500):
    t0 = time.perf_counter()
    pid.control_pid(route, speed, wps)
    t.append(time.perf_counter() - t0)
batch = torch.from_numpy(np.cumsum(np.abs(g.normal(1, 0.1, (64, 20, 2))), 1).astype(np.float32)).cuda()
for _ in range(10):
    postprocess.equal_spacing_route(batch)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    postprocess.equal_spacing_route(batch)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"control_pid_us_p50": round(float(np.median(t)) * 1e6, 1), "control_pid_us_p90": round(float(np.percentile(t, 90)) * 1e6, 1),
                  "equal_spacing_route_B64_us": round(e0.elapsed_time(e1) * 10, 2)}))
