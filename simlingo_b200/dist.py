"""Batch-sharded data parallelism helpers (one process per GPU, torch.distributed for the plumbing).

The hot path shards by independent samples (reference: Lightning DDP / ZeRO-2 are both batch-sharded DP,
``train_simlingo_seed1.sh:27``, ``config.py:299``): inference needs no data-path collective; training has exactly
one exchange step, the gradient all-reduce (``simlingo_b200.training``)."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of ``n_items`` independent samples for ``rank`` (first ranks get the
    remainder), e.g. 64 frames over 8 ranks -> 8 each."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Step time of the job = slowest rank (device-timed milliseconds are reduced with MAX)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_predictions(local: torch.Tensor, sizes: Sequence[int]) -> List[torch.Tensor] | None:
    """Optional final gather of per-rank outputs (e.g. route [b_r, 20, 2]) to rank 0; KBs, off the hot path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = max(sizes)  # gather needs equal shapes: pad the ragged last shard, trim after
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0)
    return [b[: sizes[r]] for r, b in enumerate(bufs)] if rank == 0 else None


class NativeComm:
    """NCCL communicator owned by ``libsimlingo_b200`` (C ABI ``slb_comm_init / slb_allreduce_bucket / slb_comm_destroy``,
    csrc/comm.cu), one per process.  ``torch.distributed`` is only the bootstrap: rank 0 draws the 128-byte NCCL unique id
    and broadcasts it over the existing process group.  All-reduces are enqueued on ``self.stream`` (high priority, so a
    finished bucket's NCCL kernel is scheduled ahead of the backward kernels queued behind it) or on a stream the caller
    names - including a stream that is being captured into a CUDA graph."""

    def __init__(self, process_group=None, device: torch.device | None = None):
        import ctypes as C
        from . import lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("torch.distributed is not initialised (it bootstraps the NCCL unique id)")
        if not torch.cuda.is_available():
            raise RuntimeError("simlingo_b200.NativeComm needs CUDA devices; there is no CPU fallback (gloo tests use backend='torch')")
        self._lib = lib.load()
        self.pg = process_group if process_group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.pg), dist.get_world_size(self.pg)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        raw = (C.c_char * 128)()
        if self.rank == 0:
            lib._check(self._lib.slb_comm_unique_id(raw), "comm_unique_id", 0)
        uid = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        on_gpu = "nccl" in str(dist.get_backend(self.pg))
        t = uid.to(self.device) if on_gpu else uid
        dist.broadcast(t, src=dist.get_global_rank(self.pg, 0), group=self.pg)
        blob = bytes(t.cpu().numpy().tobytes())
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            lib._check(self._lib.slb_comm_init(C.byref(handle), blob, self.rank, self.world), "comm_init", 0)
        self.handle = handle
        self.stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.calls = 0

    def all_reduce(self, t: torch.Tensor, average: bool = False, stream: torch.cuda.Stream | None = None) -> None:
        """In-place SUM (or mean) all-reduce of a contiguous bf16 / fp32 CUDA tensor, enqueued on ``stream`` (default: the
        communicator's own stream; ordering against the producer of ``t`` is the caller's business)."""
        import ctypes as C
        from . import lib
        if not (t.is_cuda and t.is_contiguous() and t.dtype in (torch.bfloat16, torch.float32)):
            raise RuntimeError(f"NativeComm.all_reduce: contiguous bf16 / fp32 CUDA tensor expected, got {t.dtype} on {t.device}")
        st = stream if stream is not None else self.stream
        lib._check(self._lib.slb_allreduce_bucket(self.handle, C.c_void_p(t.data_ptr()), C.c_int64(t.numel()), 0 if t.dtype == torch.bfloat16 else 1,
                                                  int(average), C.c_void_p(st.cuda_stream)), "allreduce_bucket", 0)   # NCCL's kernel, not one of ours: not counted
        self.calls += 1

    def destroy(self) -> None:
        from . import lib
        if getattr(self, "handle", None) is not None and self.handle.value:
            torch.cuda.synchronize(self.device)
            lib._check(self._lib.slb_comm_destroy(self.handle), "comm_destroy", 0)
            self.handle = None
