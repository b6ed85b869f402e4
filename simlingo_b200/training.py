"""Training runtime of the SimLingo hot path: forward with saved activations, hand-written backward, flat
gradient / parameter buffers and the bucketed data-parallel gradient exchange.

Replaces what the reference gets from torch autograd + flash-attn backward + PEFT + Lightning DDP / ZeRO-2 for
``DrivingModel.training_step`` (reference ``simlingo_training/models/driving.py:236-271``, ``train.py:160-217``):

* ``ParamStore``  - every trainable tensor (ViT, mlp1, LoRA A/B, heads, queries, wp_encoder; SURVEY 8a note 5) lives in
  ONE flat bf16 parameter buffer and ONE flat bf16 gradient buffer, laid out in the order the backward pass finishes
  them (LLM layers 23..0, mlp1, ViT layers 23..0, ViT embeddings, adaptor heads), so that completed ranges can be
  all-reduced over NCCL/NVLink while the rest of the backward is still running, and the optimizer is one fused
  kernel over the flat range (``simlingo_b200.optim.FusedAdamW``).
* ``TrainEngine`` - sequences the sm_100a kernels: tcgen05 GEMMs for fprop / dgrad / wgrad (``slb_gemm_bf16`` with the
  ``a_t`` / ``b_t`` operand forms), flash attention forward + backward, norm / activation / dropout / RoPE backward.
  LoRA runs unmerged (``y = W x + b + (alpha/r) B A dropout(x)``, PEFT semantics) because A, B and the dropout mask
  need gradients; the frozen Qwen2 base weights get no wgrad.
* three coarse ``torch.autograd.Function`` s (vision tower + projector, decoder stack, LM head + cross-entropy) splice
  the engine into the autograd graph of the drop-in modules; parameter gradients are written straight into the
  flat buffer (the Functions return ``None`` for them), only activation gradients travel through autograd.

No fallback: everything here needs the CUDA library and bf16 CUDA parameters."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn

from . import lib
from .engine import PATCH_KPAD
from .spec import LLM_PREFIX, MLP1_PREFIX, VIT_PREFIX, ModelSpec

_KEY = "_slb_train_engine"
_ALIGN = 64  # elements; keeps every parameter 128-byte aligned inside the flat buffers (TMA needs 16 B)


def _is_big(key: str) -> bool:
    """Matrices whose gradient is produced by a wgrad GEMM directly in bf16; everything else in an engine-managed
    group is accumulated in fp32 (column reductions / norm backward) and flushed once per group."""
    if ".lora_A." in key or ".lora_B." in key:
        return True
    if key.startswith(VIT_PREFIX + "encoder.layers."):
        return key.endswith(("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight"))
    if key.startswith(VIT_PREFIX + "embeddings."):
        return key.endswith("patch_embedding.weight")
    if key.startswith(MLP1_PREFIX):
        return key.endswith(("1.weight", "3.weight"))
    return False


class _Group:
    __slots__ = ("name", "keys", "small_keys", "start", "end", "small_start", "small_end", "acc_start", "managed")

    def __init__(self, name: str, managed: bool):
        self.name, self.managed = name, managed
        self.keys: List[str] = []
        self.small_keys: List[str] = []
        self.start = self.end = self.small_start = self.small_end = self.acc_start = 0


class ParamStore:
    """Flat bf16 parameter + gradient buffers over the trainable parameters of ``root``.

    ``prefix`` maps module-relative names to reference ``state_dict`` keys (``prefix + name``).  After construction
    every trainable ``Parameter.data`` is a view into ``flat_param`` and ``grad_view[key]`` is the matching view into
    ``flat_grad``."""

    def __init__(self, root: nn.Module, prefix: str, spec: ModelSpec, bucket_bytes: int = 48 << 20, allow_cpu: bool = False):
        """``allow_cpu`` exists for the host-logic tests of the layout / bucket bookkeeping (gloo, no kernels)."""
        self.spec = spec
        params: Dict[str, nn.Parameter] = {}
        for n, p in root.named_parameters():
            if p.requires_grad:
                if not ((p.is_cuda or allow_cpu) and p.dtype == torch.bfloat16):
                    raise RuntimeError(f"simlingo_b200 trains bf16 CUDA parameters only (no fallback): {n} is {p.dtype} on {p.device}")
                params[prefix + n] = p
        if not params:
            raise RuntimeError("no trainable parameters")
        self.params = params
        dev = next(iter(params.values())).device
        groups: List[_Group] = []

        def take(name: str, pred, managed=True):
            g = _Group(name, managed)
            ks = [k for k in params if pred(k) and not any(k in gg.keys for gg in groups)]
            if not ks:
                return
            g.keys = [k for k in ks if _is_big(k)] + [k for k in ks if not _is_big(k)] if managed else ks
            g.small_keys = [k for k in ks if not _is_big(k)] if managed else []
            groups.append(g)

        for i in reversed(range(spec.llm_layers)):
            take(f"llm{i}", lambda k, i=i: k.startswith(f"{LLM_PREFIX}model.layers.{i}."))
        take("mlp1", lambda k: k.startswith(MLP1_PREFIX))
        for i in reversed(range(spec.vit_layers)):
            take(f"vit{i}", lambda k, i=i: k.startswith(f"{VIT_PREFIX}encoder.layers.{i}."))
        take("vit_emb", lambda k: k.startswith(VIT_PREFIX + "embeddings."))
        take("other", lambda k: True, managed=False)  # heads, queries, wp_encoder: gradients come from torch autograd
        self.groups = groups
        self.group_index = {g.name: i for i, g in enumerate(groups)}

        off, acc_off = 0, 0
        offsets: Dict[str, Tuple[int, int]] = {}
        up = lambda n, a: (n + a - 1) // a * a
        for g in groups:
            g.start = off
            for k in g.keys:
                if k in g.small_keys:
                    continue
                offsets[k] = (off, params[k].numel())
                off += up(params[k].numel(), _ALIGN)
            g.small_start = off
            for k in g.small_keys:  # packed densely (16-byte granules): one flush kernel covers the whole run
                offsets[k] = (off, params[k].numel())
                off += up(params[k].numel(), 8)
            g.small_end = off
            off = up(off, _ALIGN)
            g.end = off
            g.acc_start = acc_off
            acc_off += g.small_end - g.small_start
        self.numel = off
        self.offsets = offsets
        self.flat_param = torch.zeros(off, device=dev, dtype=torch.bfloat16)
        self.flat_grad = torch.zeros(off, device=dev, dtype=torch.bfloat16)
        self.small_acc = torch.zeros(max(acc_off, 1), device=dev, dtype=torch.float32)
        self.grad_view: Dict[str, Tensor] = {}
        self.acc_view: Dict[str, Tensor] = {}
        with torch.no_grad():
            for k, p in params.items():
                o, n = offsets[k]
                v = self.flat_param[o:o + n].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.grad_view[k] = self.flat_grad[o:o + n].view(p.shape)
        for g in groups:
            for k in g.small_keys:
                o, n = offsets[k]
                a = g.acc_start + (o - g.small_start)
                self.acc_view[k] = self.small_acc[a:a + n]
        self.managed_keys = {k for g in groups if g.managed for k in g.keys}
        self.accumulate = False          # True once a backward has deposited gradients that must be added to
        self.generation = 0              # bumped by the optimizer: derived inference weights are stale
        self.weights_epoch = 0           # bumped when the parameters were overwritten from outside (load_state_dict): fp32 masters are stale
        # ---- data-parallel exchange -----------------------------------------------------------------
        self.pg = None
        self.world = 1
        self._buckets: List[Tuple[int, int, int]] = []  # (start, end, index of the last group inside)
        self._works: list = []
        self._next_bucket = 0
        self.bucket_bytes = bucket_bytes
        self._make_buckets()
        self.n_allreduce = 0
        self.comm = None               # simlingo_b200.dist.NativeComm once data parallelism is on with the native backend
        self.dp_backend = "torch"
        self._comm_pending = False
        self.require_sync = True       # False inside no_sync(): gradients accumulate locally, no exchange
        self._reduced = False          # a bucket of the current accumulation window has already been all-reduced
        self._touched: set = set()     # groups that received gradients since zero_grad (others are skipped by the optimizer)
        self._inflight: list = []      # captured forward records whose backward has not run yet (TrainEngine._mark_inflight)
        self._untouched_keys: set = set()  # torch-managed parameters whose .grad was None at the end of the backward
        self.capture = None            # set while a backward is being captured into CUDA graphs (see TrainEngine)
        self.layout_version = 0        # bumped when captured graphs must be thrown away (e.g. DP switched on)

    # ---- layout helpers -----------------------------------------------------------------------------
    def _make_buckets(self) -> None:
        self._buckets = []
        start, last = 0, -1
        for gi, g in enumerate(self.groups):
            last = gi
            if (g.end - start) * 2 >= self.bucket_bytes or gi == len(self.groups) - 1:
                self._buckets.append((start, g.end, gi))
                start = g.end

    def target(self, key: str) -> Tuple[Tensor, bool]:
        return self.grad_view[key], self.accumulate

    # ---- step protocol ------------------------------------------------------------------------------
    def begin_backward(self) -> None:
        if self._reduced and self.accumulate and self.world > 1:
            raise RuntimeError(
                "simlingo_b200: a second backward without zero_grad() while data parallelism is on would add local gradients to "
                "already all-reduced sums; run every micro-batch but the last one under `with store.no_sync():` (as with DDP)")
        self.wait_exchange()   # never drop an exchange that is still in flight
        self.small_acc.zero_()
        self._next_bucket = 0

    def no_sync(self):
        """Context manager for gradient accumulation under data parallelism (torch DDP's ``no_sync``): backward passes
        inside it only accumulate into the flat gradient buffer; the first backward outside it exchanges the sums."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self.require_sync = self.require_sync, False
            try:
                yield
            finally:
                self.require_sync = prev
        return ctx()

    def syncing(self) -> bool:
        return self.pg is not None and self.world > 1 and self.require_sync

    def skip_to_bucket_after(self, gi: int) -> None:
        """After replaying captured segments that already issued every bucket ending at or before group ``gi``."""
        while self._next_bucket < len(self._buckets) and self._buckets[self._next_bucket][2] <= gi:
            self._next_bucket += 1

    def flush_group(self, name: str) -> None:
        """fp32 accumulators of the group's small parameters -> flat bf16 gradients; then the group is final."""
        gi = self.group_index.get(name)
        if gi is None:
            return
        g = self.groups[gi]
        n = g.small_end - g.small_start
        if n > 0:
            lib.flush_f32(self.small_acc[g.acc_start:g.acc_start + n], self.flat_grad[g.small_start:g.small_end], self.accumulate)
        self._touched.add(gi)
        self.group_ready(gi)

    def enable_data_parallel(self, process_group=None, backend: Optional[str] = None) -> None:
        """Average gradients over the ranks of ``process_group`` (default: the world group): SUM all-reduce of the flat
        bf16 gradient range, bucketed and issued as the backward pass completes each bucket; the 1/world factor is
        folded into the fused optimizer kernel.

        ``backend``: "native" (default on CUDA) = the library's own NCCL communicator through the C ABI
        (``slb_comm_init`` / ``slb_allreduce_bucket``, csrc/comm.cu) on a dedicated high-priority stream, bootstrapped over
        ``process_group``; "torch" = ``torch.distributed.all_reduce`` (what the gloo host-logic tests on CPU use);
        "native-graph" = native, with the bucket all-reduces captured INSIDE the backward CUDA graphs on a forked stream
        instead of being issued between graph segments.  Env override: SLB_DP_BACKEND."""
        import os
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("torch.distributed is not initialised")
        backend = backend or os.environ.get("SLB_DP_BACKEND") or ("native" if self.flat_grad.is_cuda else "torch")
        if backend not in ("native", "native-graph", "torch"):
            raise ValueError(f"unknown data-parallel backend {backend!r}")
        if backend == "native-graph" and os.environ.get("SLB_ALLOW_NATIVE_GRAPH") != "1":
            raise RuntimeError("simlingo_b200: the 'native-graph' exchange (all-reduces captured inside the backward CUDA graphs) is experimental: "
                               "its 2-GPU test hung on the B200 box (profiles/r02_dp_tests_2gpu_v4.log); set SLB_ALLOW_NATIVE_GRAPH=1 to try it")
        self.pg = process_group if process_group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.pg)
        self.dp_backend = backend
        if backend != "torch" and self.world > 1 and self.comm is None:
            from .dist import NativeComm
            self.comm = NativeComm(self.pg, self.flat_grad.device)
        self.layout_version += 1

    def disable_data_parallel(self) -> None:
        self.wait_exchange()
        self.pg, self.world = None, 1
        self.layout_version += 1

    def launch_bucket(self, k: int, stream=None) -> None:
        """All-reduce of bucket ``k`` behind everything queued so far on the current stream."""
        a, b, _ = self._buckets[k]
        if self.comm is not None and self.dp_backend != "torch":
            cs = self.comm.stream
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            cs.wait_event(ev)                      # (while capturing: the communication stream joins the capture here)
            self.comm.all_reduce(self.flat_grad[a:b], stream=cs)
            self._comm_pending = True
        else:
            # async_op: NCCL runs on its own stream after an event on the current (compute) stream
            self._works.append(dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        self.n_allreduce += 1
        self._reduced = True

    def in_graph(self) -> bool:
        return self.comm is not None and self.dp_backend == "native-graph"

    def group_ready(self, gi: int) -> None:
        if self.pg is None or self.world == 1:
            return
        while self._next_bucket < len(self._buckets) and self._buckets[self._next_bucket][2] <= gi:
            k = self._next_bucket
            self._next_bucket += 1
            if self.capture is not None:
                if self.in_graph():
                    self.launch_bucket(k)   # captured: a forked branch of the backward graph, joined at the end of the capture
                    self.capture.n_inline += 1
                else:
                    self.capture.split(k)   # close the graph segment here; the all-reduce is issued between two replays
            elif self.require_sync:
                self.launch_bucket(k)

    def finish_backward(self) -> None:
        """End-of-backward hook: adopt gradients autograd produced for the torch-managed parameters, expose
        ``p.grad`` views, launch the remaining all-reduce buckets."""
        dst, src = [], []
        for k, p in self.params.items():
            v = self.grad_view[k]
            if k not in self.managed_keys:
                if p.grad is None:
                    if not self.accumulate:
                        self._untouched_keys.add(k)   # torch.optim.AdamW skips parameters without a gradient
                else:
                    self._untouched_keys.discard(k)
                    if p.grad.data_ptr() != v.data_ptr():
                        dst.append(v)
                        src.append(p.grad.to(v.dtype))
            p.grad = v
        if dst:   # one multi-tensor launch for the ~25 head / query / waypoint-encoder tensors
            if self.accumulate:
                torch._foreach_add_(dst, src)
            else:
                torch._foreach_copy_(dst, src)
        if not self.groups[-1].managed:
            self._touched.add(len(self.groups) - 1)
        self.group_ready(len(self.groups) - 1)
        self.accumulate = True

    def wait_exchange(self) -> None:
        """The current stream waits for every exchange in flight (no host blocking)."""
        for w in self._works:
            w.wait()
        self._works = []
        if self._comm_pending:
            torch.cuda.current_stream().wait_stream(self.comm.stream)
            self._comm_pending = False

    def zero_grad(self) -> None:
        """One memset of the whole flat gradient buffer (0.1 ms for 652 MB): engine-managed ranges are overwritten by the
        next backward that reaches them, but a sub-network that does not run backward in a step (text-only batch, LLM-only
        call) must not leave last step's gradients behind for the exchange, the clip norm and the optimizer."""
        self.wait_exchange()
        self.flat_grad.zero_()
        for p in self.params.values():
            p.grad = None   # autograd allocates the few torch-managed gradients afresh; finish_backward adopts them into the views
        self.accumulate = False
        self._reduced = False
        self._touched.clear()
        self._untouched_keys.clear()
        for rec in self._inflight:     # a forward whose backward never came (evaluation under grad mode) is abandoned here
            if rec:
                rec["inflight"] = False
        self._inflight.clear()

    def update_ranges(self) -> List[Tuple[int, int]]:
        """Flat ranges the optimizer must update: groups that received gradients since ``zero_grad`` minus torch-managed
        parameters whose ``.grad`` stayed None (torch.optim.AdamW, which the reference uses, skips those: no weight decay,
        no moment decay).  One range covering everything in the common case."""
        if self.pg is not None and self.world > 1:
            return [(0, self.numel)]   # as DDP: every rank holds the reduced gradient of every parameter
        out: List[Tuple[int, int]] = []

        def push(a, b):
            if b <= a:
                return
            if out and out[-1][1] == a:
                out[-1] = (out[-1][0], b)
            else:
                out.append((a, b))
        for gi, g in enumerate(self.groups):
            if gi not in self._touched:
                continue
            if g.managed or not self._untouched_keys:
                push(g.start, g.end)
            else:
                for k in g.keys:
                    if k not in self._untouched_keys:
                        o, n = self.offsets[k]
                        push(o, o + (n + _ALIGN - 1) // _ALIGN * _ALIGN)
        return out


# ====================================================================================================
class TrainEngine:
    def __init__(self, root: nn.Module, prefix: str, spec: ModelSpec):
        if not torch.cuda.is_available():
            raise RuntimeError("simlingo_b200 training needs a CUDA device (sm_100a); there is no CPU fallback")
        lib.load()
        self.spec = spec
        self.root = root
        self.store = ParamStore(root, prefix, spec)
        self.P: Dict[str, Tensor] = {prefix + k: v for k, v in root.state_dict(keep_vars=True).items()}
        self.dev = self.store.flat_param.device
        self.launches = 0
        self.base_seed = 0x5151
        self.seed_dev = torch.zeros(1, device=self.dev, dtype=torch.int64)  # step counter read by the dropout kernels
        self._frozen = None
        self._side = None
        self._deferred: list = []
        self._defer_rr = 0
        # CUDA graphs: the ~2800 launches of a step are recorded once per input shape and replayed; the first call
        # with a new shape runs eagerly (it also warms up lazily initialised state), the second one captures.
        import os
        self.graphs_enabled = os.environ.get("SLB_TRAIN_GRAPHS", "1") != "0"
        self._pool = None
        self._cap_stream = None
        self._recs: Dict[tuple, dict] = {}   # insertion-ordered: least recently used first
        self._seen: Dict[tuple, int] = {}
        self.graph_replays = 0
        # every captured shape pins its activations in the graph pool (~2.6 GB per sample at the 1B config): keep few
        self.max_graph_shapes = int(os.environ.get("SLB_TRAIN_GRAPH_SHAPES", "4"))
        # the reference's ``freeze=True`` encoder option (vlm.py:36-44) leaves only mlp1 trainable: the ViT then runs without
        # saved activations and the backward stops at the projector input
        vit_keys = [k for k in self.P if k.startswith(VIT_PREFIX)]
        n_train = sum(1 for k in vit_keys if k in self.store.params)
        if vit_keys and 0 < n_train < len(vit_keys):
            raise RuntimeError("simlingo_b200: the InternViT tower must be trainable as a whole or frozen as a whole "
                               f"({n_train} of {len(vit_keys)} tensors require grad)")
        self.vit_trainable = n_train > 0

    # ---- weights ---------------------------------------------------------------------------------------
    def w(self, key: str) -> Tensor:
        return self.P[key].detach()

    # the four LoRA-wrapped linear groups of a Qwen2 layer: (module prefix inside the layer, adapters sharing the input)
    _GROUPS = (("qkv", "self_attn.", ("q_proj", "k_proj", "v_proj")), ("o", "self_attn.", ("o_proj",)),
               ("gu", "mlp.", ("gate_proj", "up_proj")), ("d", "mlp.", ("down_proj",)))

    def packed(self):
        """Per layer and linear group, the frozen base weights of the adapters that share an input, stacked along the output
        dimension, with the LoRA up-projections appended along k:  wx = [W | (alpha/r) B_blockdiag]  ([sum out, in + r * n]).
        Forward:  y = [x | t] wx^T  with t = [A_j dropout_j(x)]_j  (slb_gemm_bf16 A2);  backward:  [dx | dt] = dy wx  (one
        dgrad).  The W columns are written once (frozen); the B columns are refreshed from the live parameters by ONE
        ``lora_pack`` launch at the start of every forward (``table``: their addresses inside the flat parameter store)."""
        if self._frozen is None:
            s, r = self.spec, self.spec.lora_r
            layers, table = [], []
            for i in range(s.llm_layers):
                p = f"{LLM_PREFIX}model.layers.{i}."
                ly = {}
                for name, sub, mods in self._GROUPS:
                    ws = [self.w(p + sub + m + ".base_layer.weight") for m in mods]
                    K = ws[0].shape[1]
                    wx = torch.zeros((sum(w.shape[0] for w in ws), K + r * len(mods)), device=self.dev, dtype=torch.bfloat16)
                    wx[:, :K] = torch.cat(ws, 0)
                    row = 0
                    for j, (m, w) in enumerate(zip(mods, ws)):
                        b = self.w(p + sub + m + ".lora_B.default.weight")
                        assert b.shape == (w.shape[0], r) and b.is_contiguous()
                        table.append((b.data_ptr(), wx[row:, K + r * j:].data_ptr(), w.shape[0], wx.stride(0)))
                        row += w.shape[0]
                    ly[name] = wx
                ly["bqkv"] = torch.cat([self.w(p + f"self_attn.{n}_proj.base_layer.bias") for n in "qkv"], 0).contiguous()
                layers.append(ly)
            self._frozen = dict(layers=layers, table=torch.tensor(table, dtype=torch.int64, device=self.dev), n=len(table))
        return self._frozen

    def still_attached(self) -> bool:
        """every trainable parameter is still a view into the flat buffer (False after ``.to()`` / ``.float()``)"""
        flat = self.store.flat_param
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        return all(lo <= p.data_ptr() < hi and p.dtype == flat.dtype for p in self.store.params.values())

    def weights_changed(self) -> None:
        """The parameters were overwritten in place (``load_state_dict``): the flat store and the captured graphs stay valid
        (same addresses); copies derived from the frozen weights are refreshed in place, inference engines and fp32 master
        weights are told to resynchronise."""
        if self._frozen is not None:
            for i, ly in enumerate(self._frozen["layers"]):
                p = f"{LLM_PREFIX}model.layers.{i}."
                for name, sub, mods in self._GROUPS:
                    ws = [self.w(p + sub + m + ".base_layer.weight") for m in mods]
                    ly[name][:, :ws[0].shape[1]] = torch.cat(ws, 0)
                torch.cat([self.w(p + f"self_attn.{n}_proj.base_layer.bias") for n in "qkv"], 0, out=ly["bqkv"])
        self.store.generation += 1
        self.store.weights_epoch += 1

    def _wgrad(self, dy: Tensor, x: Tensor, key: str, alpha: float = 1.0) -> None:
        """grad[key] (+)= alpha * dy^T x      dy [M, out], x [M, in] -> [out, in]"""
        g, acc = self.store.target(key)
        g2 = g.view(g.shape[0], -1)
        lib.gemm(dy, x, out=g2, a_t=True, b_t=True, alpha=alpha, residual=g2 if acc else None)

    def _acc(self, key: str) -> Tensor:
        return self.store.acc_view[key]

    def _par(self, *fns):
        """Runs ``fns[0]`` on the current stream and the others on side streams, concurrently, and joins.  Used for
        the LoRA chains: each of their GEMMs covers a quarter of the SMs at most, so independent chains overlap
        (under graph capture the fork / join events become graph edges).  Every tensor a lane touches is created
        before the fork or inside the lane and stays referenced until after the join."""
        if len(fns) == 1:
            return [fns[0]()]
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.dev) for _ in range(3)]
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        outs, joins = [None] * len(fns), []
        for k, fn in enumerate(fns[1:]):
            side = self._side[k]
            side.wait_event(fork)
            with torch.cuda.stream(side):
                outs[k + 1] = fn()
                ev = torch.cuda.Event()
                ev.record(side)
            joins.append(ev)
        outs[0] = fns[0]()
        for ev in joins:
            main.wait_event(ev)
        return outs


    def _defer(self, fn, *keep) -> None:
        """Runs ``fn`` on one of the side streams behind everything queued on the current stream so far and does NOT join: the
        current stream carries on, ``_join`` (end of the layer) waits for all deferred work.  Used for the LoRA parameter-gradient
        GEMMs of the decoder backward: 14 single-tile wgrads per layer (K = batch * length deep, ~20 us each) that nothing in the
        layer depends on; joined after every group they sat on the critical path for 0.3 ms per layer.  ``keep``: tensors the
        deferred kernels read - referenced until the join so that the caching allocator cannot hand their memory to later work
        of the main stream."""
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.dev) for _ in range(3)]
        main = torch.cuda.current_stream()
        side = self._side[self._defer_rr % len(self._side)]
        self._defer_rr += 1
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            fn()
            done = torch.cuda.Event()
            done.record(side)
        self._deferred.append((done, keep))

    def _join(self) -> None:
        main = torch.cuda.current_stream()
        for done, _ in self._deferred:
            main.wait_event(done)
        self._deferred = []

    # ==================================================================================================
    # CUDA-graph capture / replay
    # ==================================================================================================
    class _Capture:
        """Records kernels launched by ``fn`` on a side stream into one or more CUDA graphs (shared memory pool).
        ``split(k)`` closes the current segment: on replay, gradient bucket ``k`` is all-reduced after it."""

        def __init__(self, eng: "TrainEngine"):
            self.eng = eng
            self.segments: List[Tuple[torch.cuda.CUDAGraph, Optional[int]]] = []
            self.g: Optional[torch.cuda.CUDAGraph] = None
            self.n_inline = 0   # all-reduces captured inside the graph ("native-graph" backend)

        def _begin(self):
            self.g = torch.cuda.CUDAGraph()
            self.g.capture_begin(pool=self.eng._pool, capture_error_mode="thread_local")

        def split(self, k: Optional[int]):
            self.g.capture_end()
            self.segments.append((self.g, k))
            self._begin()

        def run(self, fn):
            eng = self.eng
            if eng._pool is None:
                eng._pool = torch.cuda.graph_pool_handle()
                eng._cap_stream = torch.cuda.Stream(device=eng.dev)
            import gc
            gc.collect()  # as torch.cuda.graph does: no CUDA object may be finalised while the stream is capturing
            torch.cuda.synchronize(eng.dev)
            cs = eng._cap_stream
            cs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cs):
                self._begin()
                try:
                    out = fn()
                    if self.n_inline:   # the forked communication branch must rejoin before the capture ends
                        cs.wait_stream(eng.store.comm.stream)
                        eng.store._comm_pending = False
                finally:
                    self.g.capture_end()
                self.segments.append((self.g, None))
            torch.cuda.current_stream().wait_stream(cs)
            self.keep = list(lib._ws_cache.values())  # scratch buffers baked into the graphs stay alive with them
            return out

    def _counted(self, fn, *args):
        l0 = lib.LAUNCHES
        out = fn(*args)
        self.launches += lib.LAUNCHES - l0
        return out

    def _replay(self, segments) -> None:
        st = self.store
        for g, k in segments:
            g.replay()
            if k is not None and st.syncing():
                st.launch_bucket(k)
        self.graph_replays += len(segments)

    def _graph_ok(self, key: tuple) -> bool:
        """eager on the first sighting of a shape, captured from the second on"""
        if not self.graphs_enabled:
            return False
        rec = self._recs.get(key)
        if rec is not None and rec["version"] != self.store.layout_version:
            self._drop(self._recs.pop(key))
            rec = None
        if rec is not None:
            self._recs[key] = self._recs.pop(key)  # mark as most recently used
            return True
        n = self._seen.get(key, 0)
        self._seen[key] = n + 1
        if n < 1:
            return False
        same_kind = [k for k in self._recs if k[0] == key[0]]
        if len(same_kind) >= self.max_graph_shapes:   # evict the least recently used shape of this kind
            self._drop(self._recs.pop(same_kind[0]))
        return True

    @staticmethod
    def _drop(rec: dict) -> None:
        """Release a captured shape now (the record references itself through saved['graph']; left to the cyclic GC it
        could be torn down in the middle of a later capture, which invalidates that capture)."""
        rec.get("saved", {}).pop("graph", None)
        rec.clear()

    def vision_forward_auto(self, pixels: Tensor):
        key = ("vis", int(pixels.shape[0]))
        if not self._graph_ok(key):
            return self._counted(self.vision_forward, pixels)
        rec = self._recs.get(key)
        if rec is not None and rec.get("inflight"):
            # this shape's static buffers still hold the activations of a forward whose backward has not run (micro-batches
            # forwarded back to back, a grad-enabled evaluation forward): do not overwrite them, launch eagerly
            return self._counted(self.vision_forward, pixels)
        if rec is None:
            rec = dict(version=self.store.layout_version, px=torch.empty_like(pixels), bwd=None)
            cap = TrainEngine._Capture(self)
            l0 = lib.LAUNCHES
            rec["out"], rec["saved"] = cap.run(lambda: self.vision_forward(rec["px"]))
            rec["fwd"], rec["fwd_launches"], rec["keep"] = cap.segments, lib.LAUNCHES - l0, [cap]
            rec["saved"]["graph"] = rec
            self._recs[key] = rec
        self.launches += rec["fwd_launches"]
        rec["px"].copy_(pixels)
        self._replay(rec["fwd"])
        self._mark_inflight(rec)
        return rec["out"].detach(), rec["saved"]  # fresh alias of the static output buffer for autograd

    def _mark_inflight(self, rec: dict) -> None:
        if torch.is_grad_enabled():
            rec["inflight"] = True
            self.store._inflight.append(rec)

    def vision_backward_auto(self, dout: Tensor, sv) -> None:
        rec = sv.get("graph")
        st = self.store
        if rec is not None:
            rec["inflight"] = False
        if rec is None or st.accumulate or rec["version"] != st.layout_version or (st.in_graph() and not st.require_sync):
            return self._counted(self.vision_backward, dout, sv)
        if rec["bwd"] is None:
            rec["dout"] = torch.empty_like(rec["out"])
            cap = TrainEngine._Capture(self)
            st.capture = cap
            nb, l0 = st._next_bucket, lib.LAUNCHES
            touched, st._touched = st._touched, set()
            try:
                cap.run(lambda: self.vision_backward(rec["dout"], sv))
            finally:
                st.capture = None
                rec["bwd_groups"], st._touched = st._touched, touched
            rec["bwd"], rec["bwd_launches"] = cap.segments, lib.LAUNCHES - l0
            rec["keep"].append(cap)
            st._next_bucket = nb
        self.launches += rec["bwd_launches"]
        rec["dout"].copy_(dout)
        self._replay(rec["bwd"])
        st._touched |= rec["bwd_groups"]
        st.skip_to_bucket_after(max(rec["bwd_groups"]))

    def llm_forward_auto(self, inputs: Tensor, mask: Optional[Tensor], dropout: bool):
        B, Lt, _ = inputs.shape
        key = ("llm", B, Lt, mask is not None, bool(dropout))
        if not self._graph_ok(key):
            return self._counted(self.llm_forward, inputs, mask, dropout)
        rec = self._recs.get(key)
        if rec is not None and rec.get("inflight"):
            return self._counted(self.llm_forward, inputs, mask, dropout)   # see vision_forward_auto
        if rec is None:
            rec = dict(version=self.store.layout_version, x=torch.empty_like(inputs), bwd=None,
                       mask=None if mask is None else torch.empty_like(mask, dtype=torch.bool))
            self.packed()
            cap = TrainEngine._Capture(self)
            l0 = lib.LAUNCHES
            rec["out"], rec["saved"] = cap.run(lambda: self.llm_forward(rec["x"], rec["mask"], dropout, static=True))
            rec["fwd"], rec["fwd_launches"], rec["keep"] = cap.segments, lib.LAUNCHES - l0, [cap]
            rec["saved"]["graph"] = rec
            self._recs[key] = rec
        self.launches += rec["fwd_launches"]
        rec["x"].copy_(inputs)
        if mask is not None:
            rec["mask"].copy_(mask)
        self._replay(rec["fwd"])
        self._mark_inflight(rec)
        return rec["out"].detach(), rec["saved"]

    def llm_backward_auto(self, dfeats: Tensor, sv) -> Tensor:
        rec = sv.get("graph")
        st = self.store
        if rec is not None:
            rec["inflight"] = False
        if rec is None or st.accumulate or rec["version"] != st.layout_version or (st.in_graph() and not st.require_sync):
            return self._counted(self.llm_backward, dfeats, sv)
        if rec["bwd"] is None:
            rec["dfeats"] = torch.empty_like(rec["out"])
            cap = TrainEngine._Capture(self)
            st.capture = cap
            nb, l0 = st._next_bucket, lib.LAUNCHES
            touched, st._touched = st._touched, set()
            try:
                rec["dx"] = cap.run(lambda: self.llm_backward(rec["dfeats"], sv))
            finally:
                st.capture = None
                rec["bwd_groups"], st._touched = st._touched, touched
            rec["bwd"], rec["bwd_launches"] = cap.segments, lib.LAUNCHES - l0
            rec["keep"].append(cap)
            st._next_bucket = nb
        self.launches += rec["bwd_launches"]
        rec["dfeats"].copy_(dfeats)
        self._replay(rec["bwd"])
        st._touched |= rec["bwd_groups"]
        st.skip_to_bucket_after(max(rec["bwd_groups"]))
        return rec["dx"].detach()

    # ==================================================================================================
    # vision tower + projector
    # ==================================================================================================
    def vision_forward(self, pixels: Tensor):
        s, w = self.spec, self.w
        T = pixels.shape[0]
        N, Dv = s.vit_tokens, s.vit_hidden
        M = T * N
        e = VIT_PREFIX + "embeddings."
        dev, bf = self.dev, torch.bfloat16
        pw = torch.zeros((Dv, PATCH_KPAD), device=dev, dtype=bf)
        pw[:, : s.patch_k] = w(e + "patch_embedding.weight").reshape(Dv, s.patch_k)
        cols = lib.im2col_patch(pixels.contiguous(), PATCH_KPAD)
        po = lib.gemm(cols, pw, bias=w(e + "patch_embedding.bias"))
        x = lib.vit_assemble(po, w(e + "class_embedding"), w(e + "position_embedding"), T)
        del po
        layers = []
        f32 = lambda n: torch.empty(n, device=dev, dtype=torch.float32)
        for i in range(s.vit_layers):
            p = f"{VIT_PREFIX}encoder.layers.{i}."
            st1, st2 = (f32(M), f32(M)), (f32(M), f32(M))
            h1 = lib.layernorm(x, w(p + "norm1.weight"), w(p + "norm1.bias"), s.vit_eps, stats=st1)
            qkv = lib.gemm(h1, w(p + "attn.qkv.weight"), bias=w(p + "attn.qkv.bias"))
            lse = torch.empty((T, s.vit_heads, N), device=dev, dtype=torch.float32)
            att = lib.attn_vit(qkv, T, N, s.vit_heads, lse=lse)
            # layer-scale + residual in the GEMM epilogue; the un-scaled branch output (needed by the layer-scale backward) leaves through
            # the epilogue's second output (aux_mode 1 = alpha * acc + bias)
            p1 = torch.empty((M, Dv), device=dev, dtype=bf)
            xm = lib.gemm(att, w(p + "attn.proj.weight"), bias=w(p + "attn.proj.bias"), scale_n=w(p + "ls1"), residual=x, aux=p1, aux_mode=1)
            h2 = lib.layernorm(xm, w(p + "norm2.weight"), w(p + "norm2.bias"), s.vit_eps, stats=st2)
            fpre = torch.empty((M, s.vit_mlp), device=dev, dtype=bf)   # GELU input, kept for backward (second epilogue output)
            fact = lib.gemm(h2, w(p + "mlp.fc1.weight"), bias=w(p + "mlp.fc1.bias"), act=lib.ACT_GELU, aux=fpre, aux_mode=1)
            p2 = torch.empty((M, Dv), device=dev, dtype=bf)
            xo = lib.gemm(fact, w(p + "mlp.fc2.weight"), bias=w(p + "mlp.fc2.bias"), scale_n=w(p + "ls2"), residual=xm, aux=p2, aux_mode=1)
            if self.vit_trainable:
                layers.append(dict(x=x, st1=st1, h1=h1, qkv=qkv, lse=lse, att=att, p1=p1, xm=xm, st2=st2, h2=h2, fpre=fpre, fact=fact, p2=p2))
            x = xo
        # projector: drop CLS + pixel shuffle + LN(4096) -> Linear -> GELU -> Linear
        stp = (f32(T * s.tokens_per_tile), f32(T * s.tokens_per_tile))
        y0 = lib.pixel_shuffle_ln(x, w(MLP1_PREFIX + "0.weight"), w(MLP1_PREFIX + "0.bias"), T, s.proj_eps, stats=stp)
        y1p = torch.empty((T * s.tokens_per_tile, s.llm_hidden), device=dev, dtype=bf)
        y1 = lib.gemm(y0, w(MLP1_PREFIX + "1.weight"), bias=w(MLP1_PREFIX + "1.bias"), act=lib.ACT_GELU, aux=y1p, aux_mode=1)
        y2 = lib.gemm(y1, w(MLP1_PREFIX + "3.weight"), bias=w(MLP1_PREFIX + "3.bias"))
        saved = dict(T=T, cols=cols if self.vit_trainable else None, layers=layers, xv=x, stp=stp, y0=y0, y1p=y1p, y1=y1)
        return y2, saved

    def vision_backward(self, dy2: Tensor, sv) -> None:
        s, w, st = self.spec, self.w, self.store
        T = sv["T"]
        dy2 = dy2.contiguous()
        # ---- projector ----
        lib.col_reduce(dy2, self._acc(MLP1_PREFIX + "3.bias"))
        self._wgrad(dy2, sv["y1"], MLP1_PREFIX + "3.weight")
        dy1p = lib.gemm(dy2, w(MLP1_PREFIX + "3.weight"), b_t=True, aux=sv["y1p"], aux_mode=2)
        lib.col_reduce(dy1p, self._acc(MLP1_PREFIX + "1.bias"))
        self._wgrad(dy1p, sv["y0"], MLP1_PREFIX + "1.weight")
        dy0 = lib.gemm(dy1p, w(MLP1_PREFIX + "1.weight"), b_t=True)
        dx = lib.pixel_shuffle_ln_bwd(dy0, sv["xv"], w(MLP1_PREFIX + "0.weight"), sv["stp"][0], sv["stp"][1],
                                      self._acc(MLP1_PREFIX + "0.weight"), self._acc(MLP1_PREFIX + "0.bias"), T)
        del dy0, dy1p
        st.flush_group("mlp1")
        if not self.vit_trainable:
            return   # frozen tower (freeze=True): nothing below the projector needs a gradient (pixels carry none)
        # ---- encoder layers ----
        for i in reversed(range(s.vit_layers)):
            p = f"{VIT_PREFIX}encoder.layers.{i}."
            a = sv["layers"][i]
            if "graph" not in sv:
                sv["layers"][i] = None  # eager: release the layer's activations as soon as they are consumed
            dp2 = lib.layerscale_bwd(dx, a["p2"], w(p + "ls2"), self._acc(p + "ls2"), self._acc(p + "mlp.fc2.bias"))
            self._wgrad(dp2, a["fact"], p + "mlp.fc2.weight")
            dfpre = lib.gemm(dp2, w(p + "mlp.fc2.weight"), b_t=True, aux=a["fpre"], aux_mode=2)  # dgrad * gelu'(pre) in the epilogue
            lib.col_reduce(dfpre, self._acc(p + "mlp.fc1.bias"))
            self._wgrad(dfpre, a["h2"], p + "mlp.fc1.weight")
            dh2 = lib.gemm(dfpre, w(p + "mlp.fc1.weight"), b_t=True, out=dp2)
            dxm = lib.layernorm_bwd(dh2, a["xm"], w(p + "norm2.weight"), a["st2"][0], a["st2"][1],
                                    self._acc(p + "norm2.weight"), self._acc(p + "norm2.bias"), add=dx)   # + the residual path's gradient
            dp1 = lib.layerscale_bwd(dxm, a["p1"], w(p + "ls1"), self._acc(p + "ls1"), self._acc(p + "attn.proj.bias"), out=dx)
            self._wgrad(dp1, a["att"], p + "attn.proj.weight")
            datt = lib.gemm(dp1, w(p + "attn.proj.weight"), b_t=True, out=dh2)
            delta = lib.attn_delta(a["att"], datt, T, s.vit_tokens, s.vit_heads)
            dqkv = lib.attn_vit_bwd(a["qkv"], datt, a["lse"], delta, T, s.vit_tokens, s.vit_heads)
            lib.col_reduce(dqkv, self._acc(p + "attn.qkv.bias"))
            self._wgrad(dqkv, a["h1"], p + "attn.qkv.weight")
            dh1 = lib.gemm(dqkv, w(p + "attn.qkv.weight"), b_t=True, out=dp1)
            dxi = lib.layernorm_bwd(dh1, a["x"], w(p + "norm1.weight"), a["st1"][0], a["st1"][1],
                                    self._acc(p + "norm1.weight"), self._acc(p + "norm1.bias"), dx=datt, add=dxm)
            dx = dxi
            st.flush_group(f"vit{i}")
        # ---- embeddings ----
        e = VIT_PREFIX + "embeddings."
        dpo = lib.vit_assemble_bwd(dx, self._acc(e + "class_embedding"), self._acc(e + "position_embedding"), T)
        lib.col_reduce(dpo, self._acc(e + "patch_embedding.bias"))
        gw = lib.gemm(dpo, sv["cols"], a_t=True, b_t=True)  # [1024, 640]
        g, acc = st.target(e + "patch_embedding.weight")
        g2 = g.view(s.vit_hidden, s.patch_k)
        if acc:
            g2.add_(gw[:, : s.patch_k])
        else:
            g2.copy_(gw[:, : s.patch_k])
        st.flush_group("vit_emb")

    # ==================================================================================================
    # Qwen2 decoder stack with un-merged LoRA
    # ==================================================================================================
    def _lora_down(self, x: Tensor, pre: str, mods, seeds, seed_t: Optional[Tensor]):
        """t = [A_j dropout_j(x)]_j  ([M, r * n]; first half of PEFT lora.Linear.forward for the n adapters that share the input x).
        Independent counter-based masks per adapter (PEFT keeps one nn.Dropout per wrapped linear), produced in one read of x;
        ``seed_t``: this forward's copy of the step counter.  Returns (masked copies, t)."""
        r, n = self.spec.lora_r, len(mods)
        if seeds[0] is not None:
            xds = lib.dropout_multi(x, self.spec.lora_dropout, seeds, seed_dev=seed_t)
        else:
            xds = [x] * n
        t = torch.empty((x.shape[0], r * n), device=self.dev, dtype=torch.bfloat16)
        self._par(*[(lambda j=j: lib.gemm(xds[j], self.w(pre + mods[j] + ".lora_A.default.weight"), out=t[:, r * j:r * (j + 1)])) for j in range(n)])
        return xds, t

    def _lora_bwd(self, dy: Tensor, cat: Tensor, K: int, pre: str, mods, rec, seeds, seed_t, out: Optional[Tensor] = None) -> Tensor:
        """cat = dy [W | s B] = [dx_base | dt] (the dgrad over the concatenated weight has already produced the gradient of
        every t_j).  Adds the adapters' parameter gradients  dB_j = s dy_j^T t_j,  dA_j = dt_j^T dropout_j(x)  (tcgen05 wgrad
        GEMMs, on side streams) and returns  dx = dx_base + sum_j mask_j o (dt_j A_j)  (``lora_dx``)."""
        r, n, sc = self.spec.lora_r, len(mods), self.spec.lora_scale
        xds, t = rec
        a_list = [self.w(pre + m + ".lora_A.default.weight") for m in mods]
        rows = [a.shape[0] for a in (self.w(pre + m + ".lora_B.default.weight") for m in mods)]

        # parameter gradients: off the critical path (a side stream, joined at the end of the layer).  All 2 n products of the group in
        # one launch (slb_lora_wgrad_grouped); as single-tile tcgen05 GEMMs they were 14 launches of ~25 us per layer = 9 ms per step
        grouped = r % 32 == 0 and all(x % 32 == 0 for x in rows) and K % 32 == 0 and 2 * n <= 16
        if grouped:
            probs, o = [], 0
            for j, m in enumerate(mods):
                gB, accB = self.store.target(pre + m + ".lora_B.default.weight")
                gA, accA = self.store.target(pre + m + ".lora_A.default.weight")
                probs.append((dy[:, o:o + rows[j]], t[:, r * j:r * (j + 1)], gB.view(gB.shape[0], -1), sc, accB))
                probs.append((cat[:, K + r * j:K + r * (j + 1)], xds[j], gA.view(gA.shape[0], -1), 1.0, accA))
                o += rows[j]
            self._defer(lambda: lib.lora_wgrad_grouped(probs, dy.shape[0]), dy, t, cat, *xds)
        else:
            o = 0
            for j, m in enumerate(mods):
                self._defer(lambda j=j, m=m, o=o: self._wgrad(dy[:, o:o + rows[j]], t[:, r * j:r * (j + 1)], pre + m + ".lora_B.default.weight", alpha=sc), dy, t)
                self._defer(lambda j=j, m=m: self._wgrad(cat[:, K + r * j:K + r * (j + 1)], xds[j], pre + m + ".lora_A.default.weight"), cat, xds[j])
                o += rows[j]
        use = seeds[0] is not None
        return lib.lora_dx(cat, K, a_list, p=self.spec.lora_dropout if use else 0.0, seeds=seeds if use else None, seed_dev=seed_t, out=out)

    def llm_forward(self, inputs: Tensor, mask: Optional[Tensor], dropout: bool, static: bool = False):
        """inputs [B, Lt, D] bf16 -> (features after the final norm [B, Lt, D], saved).  ``static``: no host sync on
        the mask contents (graph capture)."""
        s, w = self.spec, self.w
        B, Lt, D = inputs.shape
        M = B * Lt
        dev, bf = self.dev, torch.bfloat16
        Hq, Hkv, hd = s.llm_heads, s.llm_kv_heads, s.head_dim
        lmax = (Lt + 127) // 128 * 128
        kc = torch.zeros((s.llm_layers, B, Hkv, lmax, hd), device=dev, dtype=bf)
        vc = torch.zeros_like(kc)
        kv_valid = None
        if mask is not None and (static or not bool(mask.all())):
            kv_valid = torch.zeros((B, lmax), device=dev, dtype=torch.uint8)
            kv_valid[:, :Lt] = mask.to(torch.uint8)
        use_drop = dropout and s.lora_dropout > 0
        self.seed_dev.add_(1)  # on the device: a replayed graph draws fresh masks
        sd_t = self.seed_dev.clone()  # this forward's own copy: its backward regenerates the same masks even if another forward ran since

        def seeds(i, js):
            return [(self.base_seed << 8) + i * 8 + j if use_drop else None for j in js]

        x = inputs.reshape(M, D).contiguous()
        pk = self.packed()
        lib.lora_pack(pk["table"], pk["n"], s.lora_r, s.lora_scale)   # this step's B matrices -> the k-tail of the concatenated weights
        layers = []
        for i in range(s.llm_layers):
            p = f"{LLM_PREFIX}model.layers.{i}."
            pa, pm = p + "self_attn.", p + "mlp."
            ly = pk["layers"][i]
            r1 = torch.empty(M, device=dev, dtype=torch.float32)
            h1 = lib.rmsnorm(x, w(p + "input_layernorm.weight"), s.rms_eps, rstd=r1)
            lqkv = self._lora_down(h1, pa, ("q_proj", "k_proj", "v_proj"), seeds(i, (0, 1, 2)), sd_t)
            qkv = lib.gemm(h1, ly["qkv"], bias=ly["bqkv"], a2=lqkv[1])
            lib.rope_kv_write(qkv, kc[i], vc[i], B, Lt, 0, Hq, Hkv, s.rope_theta)
            lse = torch.empty((B, Hq, Lt), device=dev, dtype=torch.float32)
            att = lib.attn_gqa(qkv, s.qkv_dim, kc[i], vc[i], B, Lt, 0, Hq, Hkv, key_valid=kv_valid, lse=lse)
            lo = self._lora_down(att, pa, ("o_proj",), seeds(i, (3,)), sd_t)
            xm = lib.gemm(att, ly["o"], residual=x, a2=lo[1])
            r2 = torch.empty(M, device=dev, dtype=torch.float32)
            h2 = lib.rmsnorm(xm, w(p + "post_attention_layernorm.weight"), s.rms_eps, rstd=r2)
            lgu = self._lora_down(h2, pm, ("gate_proj", "up_proj"), seeds(i, (4, 5)), sd_t)
            gu = lib.gemm(h2, ly["gu"], a2=lgu[1])          # [M, 2I] = [gate | up]: one GEMM for both projections
            act = lib.silu_mul_cat(gu)
            ld = self._lora_down(act, pm, ("down_proj",), seeds(i, (6,)), sd_t)
            xo = lib.gemm(act, ly["d"], residual=xm, a2=ld[1])
            layers.append(dict(x=x, r1=r1, qkv=qkv, lse=lse, att=att, xm=xm, r2=r2, gu=gu, lora=(lqkv, lo, lgu, ld),
                               seeds=(seeds(i, (0, 1, 2)), seeds(i, (3,)), seeds(i, (4, 5)), seeds(i, (6,)))))
            x = xo
        rf = torch.empty(M, device=dev, dtype=torch.float32)
        feats = lib.rmsnorm(x, w(LLM_PREFIX + "model.norm.weight"), s.rms_eps, rstd=rf)
        saved = dict(B=B, Lt=Lt, layers=layers, kc=kc, vc=vc, kv_valid=kv_valid, xf=x, rf=rf, seed_t=sd_t)
        return feats.view(B, Lt, D), saved

    def llm_backward(self, dfeats: Tensor, sv) -> Tensor:
        s, w, st = self.spec, self.w, self.store
        B, Lt = sv["B"], sv["Lt"]
        D, I = s.llm_hidden, s.llm_mlp
        M = B * Lt
        Hq, Hkv = s.llm_heads, s.llm_kv_heads
        kvv, sd_t = sv["kv_valid"], sv["seed_t"]
        pk = self.packed()
        dx = lib.rmsnorm_bwd(dfeats.reshape(M, D).contiguous(), sv["xf"], w(LLM_PREFIX + "model.norm.weight"), sv["rf"])
        for i in reversed(range(s.llm_layers)):
            p = f"{LLM_PREFIX}model.layers.{i}."
            pa, pm = p + "self_attn.", p + "mlp."
            a, ly = sv["layers"][i], pk["layers"][i]
            if "graph" not in sv:
                sv["layers"][i] = None
            lqkv, lo, lgu, ld = a["lora"]
            sq, so, sg, sdn = a["seeds"]
            # ---- MLP ----
            cat = lib.gemm(dx, ly["d"], b_t=True)                       # [dact_base | dt_down]
            dact = self._lora_bwd(dx, cat, I, pm, ("down_proj",), ld, sdn, sd_t)
            dgu = lib.silu_mul_cat_bwd(a["gu"], dact)                   # [dgate | dup]
            cat = lib.gemm(dgu, ly["gu"], b_t=True)                     # [dh2_base | dt_gate dt_up]
            dh2 = self._lora_bwd(dgu, cat, D, pm, ("gate_proj", "up_proj"), lgu, sg, sd_t)
            dxm = lib.rmsnorm_bwd(dh2, a["xm"], w(p + "post_attention_layernorm.weight"), a["r2"], add=dx)   # + the residual path's gradient
            # ---- attention ----
            cat = lib.gemm(dxm, ly["o"], b_t=True)                      # [datt_base | dt_o]
            datt = self._lora_bwd(dxm, cat, Hq * s.head_dim, pa, ("o_proj",), lo, so, sd_t)
            delta = lib.attn_delta(a["att"], datt, B, Lt, Hq)
            dq, dk, dv = lib.attn_gqa_bwd(a["qkv"], s.qkv_dim, sv["kc"][i], sv["vc"][i], datt, a["lse"], delta, B, Lt, Hq, Hkv, key_valid=kvv)
            dqkv = lib.rope_bwd(dq, dk, dv, B, Lt, Hq, Hkv, s.rope_theta)
            del dq, dk, dv
            cat = lib.gemm(dqkv, ly["qkv"], b_t=True)                   # [dh1_base | dt_q dt_k dt_v]
            dh1 = self._lora_bwd(dqkv, cat, D, pa, ("q_proj", "k_proj", "v_proj"), lqkv, sq, sd_t)
            dxi = lib.rmsnorm_bwd(dh1, a["x"], w(p + "input_layernorm.weight"), a["r1"], add=dxm)
            dx = dxi
            self._join()   # the layer's LoRA wgrads (side streams) are complete before the group is declared final
            st.flush_group(f"llm{i}")
        return dx.view(B, Lt, D)


# ====================================================================================================
# engine lookup and autograd splice
# ====================================================================================================
def engine_for(module: nn.Module, prefix: str, spec: ModelSpec) -> TrainEngine:
    eng: Optional[TrainEngine] = module.__dict__.get(_KEY)
    if eng is None:
        eng = TrainEngine(module, prefix, spec)
        module.__dict__[_KEY] = eng
    return eng


def attach(module: nn.Module, eng: TrainEngine) -> None:
    module.__dict__[_KEY] = eng


def ensure_store(root: nn.Module, spec: ModelSpec) -> ParamStore:
    """Builds (once) the flat parameter / gradient store over the whole model and shares its engine with the
    sub-modules that can also be entered on their own (``extract_feature``, ``Qwen2ForCausalLM``)."""
    cached = root.__dict__.get(_KEY)
    if cached is not None:
        return cached.store
    eng = engine_for(root, "", spec)
    for m in root.modules():
        if hasattr(m, "spec") and m is not root and type(m).__name__ in ("InternVLChatModel", "Qwen2ForCausalLM"):
            attach(m, eng)
    return eng.store


def _anchor(eng: TrainEngine, prefix: str) -> Tensor:
    """Any trainable parameter of the sub-network: ties the Function into the autograd graph (its own gradient is
    written to the flat buffer like all the others)."""
    for k, p in eng.store.params.items():
        if k.startswith(prefix):
            return p
    raise RuntimeError(f"no trainable parameter under {prefix}")


def _queue_finish(store: ParamStore) -> None:
    if not getattr(store, "_finish_queued", False):
        store.begin_backward()   # may refuse (second reduction of the same sums): nothing is queued in that case
        store._finish_queued = True

        def cb():
            store._finish_queued = False
            store.finish_backward()
        torch.autograd.Variable._execution_engine.queue_callback(cb)


class _VisionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pixels: Tensor, anchor: Tensor, eng: TrainEngine):
        out, saved = eng.vision_forward_auto(pixels)
        ctx.eng, ctx.saved = eng, saved
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        _queue_finish(ctx.eng.store)
        ctx.eng.vision_backward_auto(dout.contiguous(), ctx.saved)
        ctx.saved = None
        return None, None, None


class _LLMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs: Tensor, anchor: Tensor, eng: TrainEngine, mask: Optional[Tensor], dropout: bool):
        feats, saved = eng.llm_forward_auto(inputs.contiguous(), mask, dropout)
        ctx.eng, ctx.saved = eng, saved
        return feats

    @staticmethod
    def backward(ctx, dfeats: Tensor):
        _queue_finish(ctx.eng.store)
        dx = ctx.eng.llm_backward_auto(dfeats.to(torch.bfloat16).contiguous(), ctx.saved)
        ctx.saved = None
        return dx, None, None, None, None


class _HeadCEFn(torch.autograd.Function):
    """loss[r] = CE(lm_head(x[r]), label[r]) with the frozen LM head: fp32 logits from the tcgen05 GEMM, fused
    softmax / loss / dlogits kernel, and d x = dlogits @ W computed right away (the upstream gradient is a per-row
    scale)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, labels: Tensor):
        xb = x.to(torch.bfloat16).contiguous()
        logits = lib.gemm(xb, weight.detach(), out_fp32=True)
        need = x.requires_grad
        loss, dl = lib.ce_fwd_bwd(logits, labels, 1.0, want_grad=need)
        if need:
            wpad = weight.detach()
            V = wpad.shape[0]
            ctx.dx_unit = lib.gemm(dl[:, :V], wpad, b_t=True)
        return loss

    @staticmethod
    def backward(ctx, dloss: Tensor):
        return (ctx.dx_unit.float() * dloss.float()[:, None]).to(ctx.dx_unit.dtype), None, None


# ---- entry points used by simlingo_b200.runtime / the drop-in modules ------------------------------
def extract_feature(chat_model: nn.Module, pixel_values: Tensor) -> Tensor:
    spec = chat_model.spec
    prefix = VIT_PREFIX[: -len("vision_model.")]
    eng = engine_for(chat_model, prefix, spec)
    px = pixel_values.to(torch.bfloat16).contiguous()
    out = _VisionFn.apply(px, _anchor(eng, MLP1_PREFIX), eng)
    return out.view(px.shape[0], spec.tokens_per_tile, spec.llm_hidden)


def llm_forward(causal_lm: nn.Module, inputs_embeds: Tensor, attention_mask: Optional[Tensor], want_logits: bool):
    spec = causal_lm.spec
    eng = engine_for(causal_lm, LLM_PREFIX, spec)
    mask = None if attention_mask is None else attention_mask.to(torch.bool)
    feats = _LLMFn.apply(inputs_embeds.to(torch.bfloat16), _anchor(eng, LLM_PREFIX), eng, mask, bool(causal_lm.training))
    logits = None
    if want_logits:
        B, L, D = feats.shape
        logits = linear_frozen(feats.reshape(B * L, D), causal_lm.lm_head.weight).view(B, L, spec.vocab)
    return feats, logits


class _FrozenLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor):
        ctx.weight, ctx.shape = weight.detach(), x.shape
        flat = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        return lib.gemm(flat, ctx.weight, out_fp32=True).view(*x.shape[:-1], weight.shape[0]).to(x.dtype)

    @staticmethod
    def backward(ctx, dy: Tensor):
        V = ctx.weight.shape[0]
        vp = (V + 7) // 8 * 8
        d2 = torch.zeros((dy.numel() // V, vp), device=dy.device, dtype=torch.bfloat16)
        d2[:, :V] = dy.reshape(-1, V)
        return lib.gemm(d2[:, :V], ctx.weight, b_t=True).view(ctx.shape), None


def linear_frozen(x: Tensor, weight: Tensor) -> Tensor:
    """x @ weight^T for a frozen weight (LM head) with gradient to x only."""
    return _FrozenLinearFn.apply(x, weight)


def lm_head_ce(features_rows: Tensor, weight: Tensor, labels: Tensor) -> Tensor:
    """Per-row cross-entropy of ``lm_head(features_rows)`` against ``labels`` (fp32 [R]); ``adaptors.py:268-274``
    evaluated only where a label exists."""
    return _HeadCEFn.apply(features_rows, weight, labels.contiguous())
