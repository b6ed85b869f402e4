"""Why is a step slower under torchrun with N > 1?  Prints, per rank: CPU budget of the box (affinity, cgroup quota, load), the
host-side kernel launch rate through the ctypes binding (before / after the NCCL communicator exists), and CPU issue time vs
device time of one offline step.
    python tools/diag_multigpu_cpu.py                       (1 process)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/diag_multigpu_cpu.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simlingo_b200 import lib  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)


def say(*a):
    print(f"[rank {rank}/{world}]", *a, flush=True)


def read(p):
    try:
        return open(p).read().strip()
    except OSError as e:
        return f"<{e.__class__.__name__}>"


say("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "cgroup cpu.max", read("/sys/fs/cgroup/cpu.max"),
    "loadavg", read("/proc/loadavg"), "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"))
lib.load()
a = torch.zeros(8, device=dev, dtype=torch.bfloat16)
b = torch.zeros(8, device=dev, dtype=torch.bfloat16)


def launch_rate(tag, n=20000):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        lib.add_inplace(a, b)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    say(f"{tag}: {1e6 * (t1 - t0) / n:.2f} us/launch issue, {1e6 * (t2 - t0) / n:.2f} us/launch incl. drain")


def spin(tag, n=3_000_000):
    t0 = time.perf_counter()
    x = 0
    for i in range(n):
        x += i
    say(f"{tag}: python loop {1e9 * (time.perf_counter() - t0) / n:.1f} ns/iter")


spin("before nccl")
launch_rate("before nccl")
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    t = torch.ones(1, device=dev)
    dist.all_reduce(t)
    torch.cuda.synchronize()
    spin("after nccl init")
    launch_rate("after nccl init")
    say("threads in process:", len(os.listdir("/proc/self/task")))

import bench as Bn  # noqa: E402
from simlingo_b200 import spec as S  # noqa: E402

spec = S.INTERNVL2_1B
model = Bn.build_model(spec, dev)
ex = Bn.make_example(Bn.host_batch(spec, 16, 1), dev)
for _ in range(2):
    Bn.offline_step(model, ex)
torch.cuda.synchronize()
for it in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    Bn.offline_step(model, ex)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    say(f"offline step B=16: cpu issue {1e3 * (t1 - t0):.1f} ms, device {e0.elapsed_time(e1):.1f} ms")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
