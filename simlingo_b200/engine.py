"""Inference runtime of the SimLingo hot path on top of the C-ABI kernels.

Sequences the sm_100a kernels for: InternViT-300M (UPSTREAM ``InternVisionModel``), pixel-shuffle +
``mlp1`` (``InternVLChatModel.extract_feature``), placeholder substitution
(reference ``internvl2_model.py:17-144``), Qwen2-0.5B with LoRA folded into scratch weights
(``llm.py:106-118``; state_dict untouched), KV-cached greedy decoding with the reference's
``greedy_sample`` semantics (``llm.py:178-250``), the 30-query append pass and the driving heads
(``driving.py:104-187``, ``adaptors.py:163-180``).

The engine reads weights from a mapping keyed by the reference ``state_dict`` names holding bf16 CUDA
tensors (normally the live ``nn.Parameter`` storage of ``simlingo_training.models.driving.DrivingModel``).
All math runs in ``libsimlingo_b200.so``; torch is used for allocation, views and tiny index glue only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib
from .spec import LLM_PREFIX, MLP1_PREFIX, VIT_PREFIX, ModelSpec

Tensor = torch.Tensor
PATCH_KPAD = 640  # 588 = 3*14*14 padded to a multiple of 64 (TMA rows must be 16-byte aligned)


def _interleave_gate_up(wg: Tensor, wu: Tensor) -> Tensor:
    """[g0..127, u0..127, g128..255, u128..255, ...] so that one 256-wide GEMM tile holds matching
    gate / up columns for the fused SwiGLU epilogue."""
    I, K = wg.shape
    assert I % 128 == 0
    return torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], dim=1).reshape(2 * I, K).contiguous()


class Engine:
    def __init__(self, sd: Dict[str, Tensor], spec: ModelSpec, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("simlingo_b200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        lib.load()
        self.spec = spec
        self.sd = sd
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self._packed_version = None
        self.launches = 0
        self.generation_fn = None  # set by runtime when a training ParamStore updates the weights behind torch's back
        # CUDA graphs for the launch-bound paths (ViT on 2 tiles, batch-1 prefill, decode steps, query append): first call
        # with a shape runs eagerly, the second captures, later ones replay.  SLB_INFER_GRAPHS=0 disables.
        import os
        self.graphs_enabled = os.environ.get("SLB_INFER_GRAPHS", "1") != "0"
        # greedy decode as ONE persistent kernel (csrc/decode.cu: token loop, grid barriers and EOS test on the device) instead of the
        # per-token chain of ~100 launches.  Opt-in (SLB_DECODE_MEGA=1): correct (tests/test_decode_gpu.py) but measured 1.14 ms per
        # token at batch 1 against 0.93 ms for the PDL-chained kernels and level with them at batch 32 (profiles/r02_trace_decode_v7.log)
        self.decode_mega = os.environ.get("SLB_DECODE_MEGA", "0") == "1"
        self._graphs: Dict[tuple, dict] = {}
        self._seen: Dict[tuple, int] = {}
        self._pool = None
        self._cap_stream = None
        self.graph_replays = 0
        self.refresh()

    # ------------------------------------------------------------------------------------------
    # derived (scratch) weights
    # ------------------------------------------------------------------------------------------
    def _w(self, name: str) -> Tensor:
        t = self.sd[name]
        if t.dtype != torch.bfloat16 or not t.is_cuda:
            raise RuntimeError(f"engine weight {name} must be a bf16 CUDA tensor, got {t.dtype} on {t.device}")
        return t.detach()

    def _version(self) -> int:
        gen = self.generation_fn() if self.generation_fn is not None else 0
        return sum(int(getattr(v, "_version", 0)) for v in self.sd.values()) + (gen << 32)

    def refresh(self, force: bool = True) -> None:
        """(Re)build the scratch weights derived from the state_dict: zero-padded patch-embed matrix, fused
        QKV / gate-up matrices and LoRA folded in fp32 (W + (alpha/r) B A) then rounded once to bf16.
        Parts whose keys are absent (an engine owned by a sub-module) are skipped."""
        ver = self._version()
        if not force and ver == self._packed_version:
            return
        self._graphs.clear()  # captured graphs hold pointers into the scratch weights rebuilt below
        self._seen.clear()
        s, w = self.spec, self._w
        e = VIT_PREFIX + "embeddings."
        self.has_vit = (e + "patch_embedding.weight") in self.sd
        if self.has_vit:
            pw = torch.zeros((s.vit_hidden, PATCH_KPAD), device=self.dev, dtype=torch.bfloat16)
            pw[:, : s.patch_k] = w(e + "patch_embedding.weight").reshape(s.vit_hidden, s.patch_k)
            self.patch_w = pw
        sc = s.lora_scale

        def merged(prefix: str) -> Tensor:
            base = w(prefix + "base_layer.weight").float()
            a, b = self.sd.get(prefix + "lora_A.default.weight"), self.sd.get(prefix + "lora_B.default.weight")
            if a is not None:
                base = base + sc * (b.detach().float() @ a.detach().float())
            return base

        self.llm_layers = []
        self.has_llm = (LLM_PREFIX + "model.norm.weight") in self.sd
        self.has_lora = self.has_llm and (f"{LLM_PREFIX}model.layers.0.self_attn.q_proj.lora_A.default.weight") in self.sd
        bf = torch.bfloat16

        def lora(prefix: str):
            """(A [r, in], sc * B [out, r]) in bf16 (sc = alpha / r = 2: exact)"""
            a, b = self.sd[prefix + "lora_A.default.weight"].detach(), self.sd[prefix + "lora_B.default.weight"].detach()
            return a.to(bf), (sc * b.float()).to(bf)

        for i in range(s.llm_layers if self.has_llm else 0):
            p = f"{LLM_PREFIX}model.layers.{i}."
            # (1) LoRA folded into the weights, rounded once: the decode steps (one new position per sequence, weight-streaming
            #     kernels, launch-bound) - token ids and the K/V of generated positions
            qkv = torch.cat([merged(p + f"self_attn.{n}_proj.") for n in "qkv"], 0).to(bf).contiguous()
            bqkv = torch.cat([w(p + f"self_attn.{n}_proj.base_layer.bias") for n in "qkv"], 0).contiguous()
            o = merged(p + "self_attn.o_proj.").to(bf).contiguous()
            gu = _interleave_gate_up(merged(p + "mlp.gate_proj.").to(bf), merged(p + "mlp.up_proj.").to(bf))
            d = merged(p + "mlp.down_proj.").to(bf).contiguous()
            ly = dict(qkv=qkv, bqkv=bqkv, o=o, gu=gu, d=d, ln1=w(p + "input_layernorm.weight"), ln2=w(p + "post_attention_layernorm.weight"))
            if self.has_lora:
                # (2) un-merged, exact: y = [x | t] [W | sc B]^T with t = A x riding in the base GEMM's k loop (slb_gemm_bf16 A2).
                #     Folding rounds W + sc B A to bf16, which alone costs ~2 % of route / waypoint accuracy at 24 layers
                #     (profiles/r02_diag_rounding_e2e.log); prefill, the teacher-forced pass and the query pass use this set.
                r = s.lora_r
                aq, bq = lora(p + "self_attn.q_proj.")
                ak, bk = lora(p + "self_attn.k_proj.")
                av, bv = lora(p + "self_attn.v_proj.")
                qd, kd = bq.shape[0], bk.shape[0]
                b2 = torch.zeros((qd + 2 * kd, 3 * r), device=self.dev, dtype=bf)
                b2[:qd, :r], b2[qd:qd + kd, r:2 * r], b2[qd + kd:, 2 * r:] = bq, bk, bv
                ly["qkv_x"] = torch.cat([torch.cat([w(p + f"self_attn.{n}_proj.base_layer.weight") for n in "qkv"], 0), b2], 1).contiguous()
                ly["a_qkv"] = torch.cat([aq, ak, av], 0).contiguous()
                ao, bo = lora(p + "self_attn.o_proj.")
                ly["o_x"], ly["a_o"] = torch.cat([w(p + "self_attn.o_proj.base_layer.weight"), bo], 1).contiguous(), ao.contiguous()
                ag, bg = lora(p + "mlp.gate_proj.")
                au, bu = lora(p + "mlp.up_proj.")
                z = torch.zeros_like(bg)
                ly["gu_x"] = _interleave_gate_up(torch.cat([w(p + "mlp.gate_proj.base_layer.weight"), bg, z], 1),
                                                 torch.cat([w(p + "mlp.up_proj.base_layer.weight"), z, bu], 1))
                ly["a_gu"] = torch.cat([ag, au], 0).contiguous()
                ad, bd = lora(p + "mlp.down_proj.")
                ly["d_x"], ly["a_d"] = torch.cat([w(p + "mlp.down_proj.base_layer.weight"), bd], 1).contiguous(), ad.contiguous()
            self.llm_layers.append(ly)
        a = "adaptors.driving."
        if (a + "route_head.0.weight") in self.sd:
            self.heads_w = lib.HeadsWeights(*[w(a + k).data_ptr() for k in (
                "route_head.0.weight", "route_head.0.bias", "route_head.2.weight", "route_head.2.bias", "route_head.4.weight",
                "speed_wps_head.0.weight", "speed_wps_head.0.bias", "speed_wps_head.2.weight")])
            self.queries = torch.cat([w(a + "query_embeds_wps"), w(a + "query_embeds_speed")], 1)[0].contiguous()  # [30, D]
        if "wp_encoder.mlp.0.weight" in self.sd:
            self.wp_w = lib.WpWeights(*[w("wp_encoder.mlp." + k).data_ptr() for k in (
                "0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias")])
        self._packed_version = ver


    # ------------------------------------------------------------------------------------------
    # CUDA-graph helpers
    # ------------------------------------------------------------------------------------------
    def _capture(self, fn):
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
            self._cap_stream = torch.cuda.Stream(device=self.dev)
        import gc
        gc.collect()  # as torch.cuda.graph does: no CUDA object may be finalised while the stream is capturing
        torch.cuda.synchronize(self.dev)
        cs = self._cap_stream
        cs.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        l0 = lib.LAUNCHES
        with torch.cuda.stream(cs):
            g.capture_begin(pool=self._pool, capture_error_mode="thread_local")
            try:
                out = fn()
            finally:
                g.capture_end()
        torch.cuda.current_stream().wait_stream(cs)
        return g, out, lib.LAUNCHES - l0

    def _use_graph(self, key: tuple) -> bool:
        if not self.graphs_enabled:
            return False
        if key in self._graphs:
            return True
        n = self._seen.get(key, 0)
        self._seen[key] = n + 1
        if n >= 1 and len(self._graphs) >= 12:   # bound the memory held by static buffers: drop the oldest shape
            self._graphs.pop(next(iter(self._graphs)))
        return n >= 1

    def _replay(self, g, n_launches: int) -> None:
        g.replay()
        self.graph_replays += 1
        self.launches += n_launches

    def extract_feature_auto(self, pixels: Tensor) -> Tensor:
        """``extract_feature`` behind a CUDA graph for small tile counts (the agent's 2 tiles are launch-bound)."""
        T = int(pixels.shape[0])
        key = ("vit", T)
        if T > 8 or not self._use_graph(key):
            return self.extract_feature(pixels)
        rec = self._graphs.get(key)
        if rec is None:
            rec = dict(px=torch.empty_like(pixels))
            rec["g"], rec["out"], rec["n"] = self._capture(lambda: self.extract_feature(rec["px"]))
            self._graphs[key] = rec
        rec["px"].copy_(pixels)
        self._replay(rec["g"], rec["n"])
        return rec["out"]

    # ------------------------------------------------------------------------------------------
    # InternViT + projector
    # ------------------------------------------------------------------------------------------
    def vit(self, pixels: Tensor, collect: Optional[list] = None) -> Tensor:
        """pixels [T,3,448,448] bf16 -> hidden [T*1025, 1024] (after the last encoder layer), fp32: as in the decoder the
        residual stream stays in fp32 (LayerNorm reads it, the proj / fc2 epilogues add the layer-scaled branch to it)."""
        s, w = self.spec, self._w
        T = pixels.shape[0]
        e = VIT_PREFIX + "embeddings."
        # patch embedding as an implicit GEMM: the operand tile is gathered from the pixels inside the kernel, class token and
        # position embedding are added in the epilogue (one launch instead of im2col + GEMM + assemble, no [T*1024, 640] matrix in HBM)
        x = lib.patch_embed(pixels.contiguous(), self.patch_w, w(e + "patch_embedding.bias"), w(e + "class_embedding").reshape(-1),
                            w(e + "position_embedding").reshape(-1, s.vit_hidden))
        N = s.vit_tokens
        h = torch.empty((T * N, s.vit_hidden), device=self.dev, dtype=torch.bfloat16)
        qkv = torch.empty((T * N, 3 * s.vit_hidden), device=self.dev, dtype=torch.bfloat16)
        att = torch.empty_like(h)
        f = torch.empty((T * N, s.vit_mlp), device=self.dev, dtype=torch.bfloat16)
        for i in range(s.vit_layers):
            p = f"{VIT_PREFIX}encoder.layers.{i}."
            lib.layernorm(x, w(p + "norm1.weight"), w(p + "norm1.bias"), s.vit_eps, out=h)
            lib.gemm(h, w(p + "attn.qkv.weight"), out=qkv, bias=w(p + "attn.qkv.bias"))
            lib.attn_vit(qkv, T, N, s.vit_heads, out=att)
            lib.gemm(att, w(p + "attn.proj.weight"), out=x, bias=w(p + "attn.proj.bias"), scale_n=w(p + "ls1"), residual=x, out_fp32=True)
            lib.layernorm(x, w(p + "norm2.weight"), w(p + "norm2.bias"), s.vit_eps, out=h)
            lib.gemm(h, w(p + "mlp.fc1.weight"), out=f, bias=w(p + "mlp.fc1.bias"), act=lib.ACT_GELU)
            lib.gemm(f, w(p + "mlp.fc2.weight"), out=x, bias=w(p + "mlp.fc2.bias"), scale_n=w(p + "ls2"), residual=x, out_fp32=True)
            if collect is not None:
                collect.append(x.clone())
        self.launches += 1 + 7 * s.vit_layers
        return x

    def extract_feature(self, pixels: Tensor) -> Tensor:
        """``InternVLChatModel.extract_feature``: [T,3,448,448] -> [T*256, 896]."""
        s, w = self.spec, self._w
        T = pixels.shape[0]
        x = self.vit(pixels)
        y = lib.pixel_shuffle_ln(x, w(MLP1_PREFIX + "0.weight"), w(MLP1_PREFIX + "0.bias"), T, s.proj_eps)
        y = lib.gemm(y, w(MLP1_PREFIX + "1.weight"), bias=w(MLP1_PREFIX + "1.bias"), act=lib.ACT_GELU)
        self.launches += 3
        return lib.gemm(y, w(MLP1_PREFIX + "3.weight"), bias=w(MLP1_PREFIX + "3.bias"))

    # ------------------------------------------------------------------------------------------
    # prompt embeddings with <IMG_CONTEXT> / <TARGET_POINT> substitution
    # ------------------------------------------------------------------------------------------
    def wp_rows(self, ids_cpu: Tensor, placeholder_values: Sequence[dict]) -> Tuple[Optional[Tensor], Optional[Tensor], int]:
        """Host-side bookkeeping of ``replace_placeholder_tokens`` step 2a (internvl2_model.py:54-91): first
        occurrence of each added special id per row and the coordinates to encode there.  Only
        ``<TARGET_POINT>`` ever carries values, so one run per row is supported."""
        s = self.spec
        B = ids_cpu.shape[0]
        if placeholder_values is None or len(placeholder_values) == 0:
            return None, None, 0
        starts = np.full((B,), -1, dtype=np.int32)
        coords: List[np.ndarray] = []
        n_per = 0
        ids_np = ids_cpu.numpy()
        for b in range(B):
            special = sorted(set(ids_np[b][ids_np[b] >= s.first_added_id].tolist()))
            special = [sid for sid in special if int(np.nonzero(ids_np[b] == sid)[0][0]) != 0]  # reference: first_occurrences.nonzero() drops index 0
            if not special:
                continue
            if len(special) > 1:
                raise NotImplementedError("more than one placeholder id in a prompt (the released configuration only uses <TARGET_POINT>)")
            sid = special[0]
            pos = int(np.nonzero(ids_np[b] == sid)[0][0])
            c = np.asarray(placeholder_values[b][sid], dtype=np.float32).reshape(-1, 2)   # KeyError if absent, as internvl2_model.py:80
            if pos + c.shape[0] > ids_np.shape[1]:
                raise RuntimeError(f"placeholder run of token {sid} in row {b} does not fit: {pos} + {c.shape[0]} > {ids_np.shape[1]}")
            if n_per and c.shape[0] != n_per:
                raise NotImplementedError("placeholder runs of different length in one batch")
            n_per = c.shape[0]
            starts[b] = pos
            coords.append(c)
        if not coords:
            return None, None, 0
        full = np.zeros((B, n_per, 2), dtype=np.float32)
        k = 0
        for b in range(B):
            if starts[b] >= 0:
                full[b] = coords[k]
                k += 1
        return torch.from_numpy(full), torch.from_numpy(starts), n_per

    def embed_prompt(self, ids: Tensor, pixels: Optional[Tensor], placeholder_values: Optional[Sequence[dict]],
                     ids_cpu: Optional[Tensor] = None) -> Tensor:
        """ids [B,L] int64 (cuda), pixels [B,1,NP,3,448,448] bf16 -> language_inputs [B,L,896] with image and
        waypoint rows substituted (adaptors.py:256 + internvl2_model.py:54-131)."""
        s = self.spec
        B, L = ids.shape
        vit = None
        n_img = 0
        if pixels is not None and pixels.numel() > 0 and L != 1:
            BS, T, NP = pixels.shape[:3]
            assert T == 1, "Only one frame is supported for now"
            vit = self.extract_feature_auto(pixels.reshape(BS * NP, *pixels.shape[3:]).contiguous())
            n_img = NP * s.tokens_per_tile
        wp = wp_start = None
        wp_len = 0
        if placeholder_values is not None and len(placeholder_values) > 0:
            if ids_cpu is None:
                ids_cpu = ids.cpu()
            coords, starts, wp_len = self.wp_rows(ids_cpu, placeholder_values)
            if coords is not None:
                c = coords.to(self.dev, non_blocking=True).to(torch.bfloat16).float()  # wp_encoder dtype is bf16 in the agent
                wp = lib.wp_encoder(c.reshape(-1, 2).contiguous(), self.wp_w).view(B, wp_len, s.llm_hidden)
                wp_start = starts.to(self.dev, non_blocking=True)
                self.launches += 1
        self.launches += 1
        return lib.embed_assemble(ids.contiguous(), self._w(LLM_PREFIX + "model.embed_tokens.weight"), vit, wp, wp_start, wp_len,
                                  s.img_context_id, n_img)

    # ------------------------------------------------------------------------------------------
    # Qwen2 with KV cache
    # ------------------------------------------------------------------------------------------
    def new_cache(self, batch: int, lmax: int) -> Tuple[Tensor, Tensor]:
        s = self.spec
        lmax = (lmax + 127) // 128 * 128
        shape = (s.llm_layers, batch, s.llm_kv_heads, lmax, s.head_dim)
        return (torch.zeros(shape, device=self.dev, dtype=torch.bfloat16), torch.zeros(shape, device=self.dev, dtype=torch.bfloat16))

    def llm_chunk(self, x: Tensor, batch: int, lq: int, past: int, cache: Tuple[Tensor, Tensor],
                  key_valid: Optional[Tensor] = None, collect: Optional[list] = None, past_dev: Optional[Tensor] = None) -> Tensor:
        """One pass of ``lq`` new positions per sequence through all decoder layers.
        x [B*lq, 896] -> residual stream before the final norm, fp32.  The residual stream is kept in fp32 (branch outputs
        are added to it in the o_proj / down_proj epilogues without an intermediate rounding): rounding it to bf16 after each
        of the 48 sub-layers is by far the largest term of the bf16-vs-fp32 feature error (2.0 % of 2.5 % at 24 layers,
        profiles/r02_diag_rounding.log), for +4 bytes per element and sub-layer of HBM traffic."""
        s = self.spec
        kc, vc = cache
        M = batch * lq
        x = x.float() if x.dtype != torch.float32 else x
        h = torch.empty((M, s.llm_hidden), device=self.dev, dtype=torch.bfloat16)
        qkv = torch.empty((M, s.qkv_dim), device=self.dev, dtype=torch.bfloat16)
        att = torch.empty((M, s.llm_heads * s.head_dim), device=self.dev, dtype=torch.bfloat16)
        act = torch.empty((M, s.llm_mlp), device=self.dev, dtype=torch.bfloat16)
        exact = self.has_lora and lq > 1   # prefill / teacher-forced / query pass: un-merged LoRA (see refresh)
        fuse_norm = M <= 4 and not exact  # decode: RMSNorm folded into the weight-streaming GEMVs (two launches fewer per layer)
        for i, ly in enumerate(self.llm_layers):
            if exact:
                lib.rmsnorm(x, ly["ln1"], s.rms_eps, out=h)
                t = lib.gemm(h, ly["a_qkv"])
                lib.gemm(h, ly["qkv_x"], out=qkv, bias=ly["bqkv"], a2=t)
            elif fuse_norm:
                lib.gemm(x, ly["qkv"], out=qkv, bias=ly["bqkv"], rms_weight=ly["ln1"], rms_eps=s.rms_eps)
            else:
                lib.rmsnorm(x, ly["ln1"], s.rms_eps, out=h)
                lib.gemm(h, ly["qkv"], out=qkv, bias=ly["bqkv"])
            if lq == 1:   # decode step: RoPE + KV write + attention in one launch
                lib.attn_decode_rope(qkv, kc[i], vc[i], batch, past, s.llm_heads, s.llm_kv_heads, s.rope_theta, key_valid=key_valid, out=att,
                                     past_dev=past_dev)
            else:
                lib.rope_kv_write(qkv, kc[i], vc[i], batch, lq, past, s.llm_heads, s.llm_kv_heads, s.rope_theta, past_dev=past_dev)
                lib.attn_gqa(qkv, s.qkv_dim, kc[i], vc[i], batch, lq, past, s.llm_heads, s.llm_kv_heads, key_valid=key_valid, out=att,
                             past_dev=past_dev)
            if exact:
                t = lib.gemm(att, ly["a_o"])
                lib.gemm(att, ly["o_x"], out=x, residual=x, out_fp32=True, a2=t)
                lib.rmsnorm(x, ly["ln2"], s.rms_eps, out=h)
                t = lib.gemm(h, ly["a_gu"])
                lib.gemm(h, ly["gu_x"], out=act, swiglu=True, a2=t)
                t = lib.gemm(act, ly["a_d"])
                lib.gemm(act, ly["d_x"], out=x, residual=x, out_fp32=True, a2=t)
            else:
                lib.gemm(att, ly["o"], out=x, residual=x, out_fp32=True)
                if fuse_norm:
                    lib.gemm(x, ly["gu"], out=act, swiglu=True, rms_weight=ly["ln2"], rms_eps=s.rms_eps)
                else:
                    lib.rmsnorm(x, ly["ln2"], s.rms_eps, out=h)
                    lib.gemm(h, ly["gu"], out=act, swiglu=True)
                lib.gemm(act, ly["d"], out=x, residual=x, out_fp32=True)
            if collect is not None:
                collect.append(x.clone())
        self.launches += (12 if exact else 8) * s.llm_layers - (0 if not fuse_norm else 2 * s.llm_layers) - (s.llm_layers if lq == 1 else 0)
        return x

    def final_norm(self, x: Tensor) -> Tensor:
        self.launches += 1
        return lib.rmsnorm(x, self._w(LLM_PREFIX + "model.norm.weight"), self.spec.rms_eps)

    def logits(self, feats: Tensor) -> Tensor:
        """fp32 logits of the given feature rows (lm_head is not LoRA-wrapped)."""
        self.launches += 1
        return lib.gemm(feats, self._w(LLM_PREFIX + "lm_head.weight"), out_fp32=True)

    def logits_from_residual(self, x: Tensor) -> Tensor:
        """final RMSNorm + LM head on <= 4 residual-stream rows in one weight-streaming launch (decode)."""
        self.launches += 1
        return lib.gemm(x, self._w(LLM_PREFIX + "lm_head.weight"), out_fp32=True, rms_weight=self._w(LLM_PREFIX + "model.norm.weight"),
                        rms_eps=self.spec.rms_eps)

    def heads(self, feats30: Tensor, batch: int, ld_batch: int) -> Tuple[Tensor, Tensor]:
        self.launches += 2
        return lib.driving_heads(feats30, ld_batch, self.heads_w, batch)


    # ------------------------------------------------------------------------------------------
    # graphed generation: prefill graph, one decode-step graph replayed per token (position counter on the device),
    # query-append graph.  Sequences must all stop at the same step (batch 1, or EOS suppressed).
    # ------------------------------------------------------------------------------------------
    def _generate_graphed(self, lang_embeds: Tensor, max_new_tokens: int, eos_token_id: Optional[int]):
        s = self.spec
        B, L, D = lang_embeds.shape
        nq = s.n_queries
        key = ("gen", B, L, max_new_tokens, eos_token_id)
        mega = self.decode_mega and B <= 32 and max_new_tokens > 1
        rec = self._graphs.get(key)
        if rec is not None and rec["mega"] != mega:
            rec = None
        if rec is None:
            dev = self.dev
            lmax = L + max_new_tokens + nq
            ngb = torch.zeros(B + 1, device=dev, dtype=torch.int64)   # per-sequence counts + the decode kernel's status word
            rec = dict(x=torch.empty_like(lang_embeds), cache=self.new_cache(B, lmax), pos=torch.zeros(1, device=dev, dtype=torch.int32),
                       nxt=torch.zeros(B, device=dev, dtype=torch.int64), step=torch.zeros(1, device=dev, dtype=torch.int64),
                       sampled=torch.zeros((B, max_new_tokens), device=dev, dtype=torch.int64), done=torch.zeros(B, device=dev, dtype=torch.bool),
                       n_gen=ngb[:B], n_gen_buf=ngb, status=ngb[B:], mega=mega)
            emb_w = self._w(LLM_PREFIX + "model.embed_tokens.weight")
            lm_cap = rec["cache"][0].shape[3]

            def sample(last, residual=False):
                lg = self.logits_from_residual(last) if residual else self.logits(last)
                # argmax + sampled[:, pos - L] / n_gen / done / step in the same launch (was ~8 element-wise torch kernels per token)
                lib.argmax_sample(lg, rec["nxt"], rec["sampled"], rec["pos"], L, rec["done"], rec["n_gen"], rec["step"], eos_token_id)
                self.launches += 1

            def prefill():
                rec["pos"].fill_(L)
                rec["step"].zero_(); rec["done"].zero_(); rec["n_gen_buf"].zero_()
                rec["sampled"].fill_(eos_token_id if eos_token_id is not None else 0)
                x = self.llm_chunk(rec["x"].reshape(B * L, D).clone(), B, L, 0, rec["cache"], None)
                sample(self.final_norm(x.view(B, L, D)[:, -1].contiguous()))

            def decode():
                e = lib.gather_rows(emb_w, rec["nxt"])
                x = self.llm_chunk(e, B, 1, lm_cap - 1, rec["cache"], None, past_dev=rec["pos"])
                rec["pos"].add_(1)
                if B <= 4:
                    sample(x, residual=True)
                else:
                    sample(self.final_norm(x))

            if mega:
                rec["table"] = lib.decode_layer_table(self.llm_layers, dev)
                rec["ws"] = lib.decode_workspace(B, s.llm_hidden, s.llm_mlp, s.llm_heads, s.llm_kv_heads, dev)

                def decode():   # every remaining token in one persistent launch: token loop, position counter and EOS test on the device
                    lib.decode_loop(rec["table"], s.llm_layers, B, s.llm_hidden, s.llm_mlp, s.llm_heads, s.llm_kv_heads, emb_w,
                                    self._w(LLM_PREFIX + "model.norm.weight"), self._w(LLM_PREFIX + "lm_head.weight"), rec["cache"][0],
                                    rec["cache"][1], rec["pos"], rec["nxt"], rec["sampled"], rec["step"], rec["done"], rec["n_gen"],
                                    rec["status"], rec["ws"], max_new_tokens - 1, eos_token_id, s.rope_theta, s.rms_eps)

            def queries():
                e = lib.gather_rows(emb_w, rec["nxt"])
                chunk = torch.cat([e.view(B, 1, D), self.queries.unsqueeze(0).expand(B, nq, D)], 1).reshape(B * (nq + 1), D).contiguous()
                xx = self.llm_chunk(chunk, B, nq + 1, lm_cap - nq - 1, rec["cache"], None, past_dev=rec["pos"])
                f = self.final_norm(xx.view(B, nq + 1, D)[:, 1:].reshape(B * nq, D).contiguous())
                return self.heads(f, B, nq * D)

            l0 = self.launches
            rec["g_pre"], _, rec["n_pre"] = self._capture(prefill)
            rec["g_dec"], _, rec["n_dec"] = self._capture(decode)
            rec["g_q"], rec["out"], rec["n_q"] = self._capture(queries)
            self.launches = l0
            self._graphs[key] = rec
        rec["x"].copy_(lang_embeds)
        self._replay(rec["g_pre"], rec["n_pre"])
        if rec["mega"]:
            self._replay(rec["g_dec"], rec["n_dec"])
        else:
            for i in range(1, max_new_tokens):
                if eos_token_id is not None and bool(rec["done"].all()):  # host sync, as llm.py:245
                    break
                self._replay(rec["g_dec"], rec["n_dec"])
        self._replay(rec["g_q"], rec["n_q"])
        route, speed = rec["out"]
        n_gen_cpu = rec["n_gen_buf"].tolist()
        if n_gen_cpu[B] != 0:
            raise RuntimeError("simlingo_b200: the persistent decode kernel reported a grid-barrier timeout (slb_decode_loop status != 0)")
        toks = [rec["sampled"][b, : max(n_gen_cpu[b], 1)].clone() for b in range(B)]
        return speed.clone(), route.clone(), toks

    # ------------------------------------------------------------------------------------------
    # DrivingModel.forward (inference): greedy decode + 30-query pass + heads
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, lang_embeds: Tensor, valid: Optional[Tensor], max_new_tokens: int, eos_token_id: Optional[int],
                 margins: Optional[list] = None):
        """Batched, KV-cached equivalent of the reference's per-item loop
        (``driving.py:133-176`` + ``llm.py:178-250``).  lang_embeds [B,L,896]; valid [B,L] bool or None.

        Every sequence is decoded until its own EOS (EOS-prefilled ``sampled_tokens``, the EOS embedding is
        still appended, ``llm.py:232-248``); then the 30 driving queries are appended after that sequence's
        last generated token and one more chunk is run *without* key mask (``driving.py:154-156``).
        Returns (speed_wps [B,10,2] fp32, route [B,20,2] fp32, list of int64 token tensors)."""
        s = self.spec
        B, L, D = lang_embeds.shape
        nq = s.n_queries
        has_pad = valid is not None and not bool(valid.all())
        if (not has_pad and margins is None and (B == 1 or eos_token_id is None)
                and self._use_graph(("gen", B, L, max_new_tokens, eos_token_id))):
            return self._generate_graphed(lang_embeds, max_new_tokens, eos_token_id)
        lmax = L + max_new_tokens + nq
        cache = self.new_cache(B, lmax)
        kv_valid = None
        if has_pad:
            kv_valid = torch.ones((B, cache[0].shape[3]), device=self.dev, dtype=torch.uint8)
            kv_valid[:, :L] = valid.to(torch.uint8)
        emb_w = self._w(LLM_PREFIX + "model.embed_tokens.weight")
        x = self.llm_chunk(lang_embeds.reshape(B * L, D).clone(), B, L, 0, cache, kv_valid)
        last = self.final_norm(x.view(B, L, D)[:, -1].contiguous())
        sampled = torch.full((B, max_new_tokens), eos_token_id if eos_token_id is not None else 0, device=self.dev, dtype=torch.int64)
        done = torch.zeros((B,), device=self.dev, dtype=torch.bool)
        n_gen = torch.zeros((B,), device=self.dev, dtype=torch.int64)
        gen_embeds = []
        steps = 0
        for i in range(max_new_tokens):
            lg = self.logits(last)
            mg = torch.empty((B,), device=self.dev, dtype=torch.float32) if margins is not None else None
            nxt = lib.argmax(lg, out_margin=mg)
            self.launches += 1
            if margins is not None:
                margins.append(mg)
            sampled[:, i] = torch.where(done, sampled[:, i], nxt)
            n_gen += (~done).long()
            steps = i + 1
            e = lib.gather_rows(emb_w, nxt)
            gen_embeds.append(e)
            if eos_token_id is not None:
                done = done | (nxt == eos_token_id)
                if bool(done.all()):  # host sync, as llm.py:245
                    break
            if i + 1 < max_new_tokens:
                x = self.llm_chunk(e.clone(), B, 1, L + i, cache, kv_valid)
                last = self.final_norm(x)
        n_gen_cpu = n_gen.tolist()
        toks = [sampled[b, : max(n_gen_cpu[b], 1)] for b in range(B)]
        # 30-query append: positions L+G_b .. L+G_b+29, no key mask (driving.py:154-156)
        route = torch.empty((B, s.n_route, 2), device=self.dev, dtype=torch.float32)
        speed = torch.empty((B, s.n_speed, 2), device=self.dev, dtype=torch.float32)
        if has_pad:
            # the reference re-runs prompt + generated + queries from scratch without a mask, so pads are
            # attended by every position: recompute exactly that for padded rows
            for b in range(B):
                G = n_gen_cpu[b]
                seq = torch.cat([lang_embeds[b], torch.cat([g[b:b + 1] for g in gen_embeds[:G]], 0), self.queries], 0)
                c1 = self.new_cache(1, seq.shape[0])
                xx = self.llm_chunk(seq.clone(), 1, seq.shape[0], 0, c1, None)
                f = self.final_norm(xx[-nq:].contiguous())
                r, sp = self.heads(f, 1, nq * D)
                route[b], speed[b] = r[0], sp[0]
        elif all(g == steps for g in n_gen_cpu):
            # every sequence stopped on its last sampled token, whose K/V are not cached yet (decoding stops right
            # after sampling it): run [last token | 30 queries] as one 31-row chunk at positions L+G-1 .. L+G+29
            G = steps
            chunk = torch.cat([gen_embeds[G - 1].view(B, 1, D), self.queries.unsqueeze(0).expand(B, nq, D)], 1)
            xx = self.llm_chunk(chunk.reshape(B * (nq + 1), D).contiguous(), B, nq + 1, L + G - 1, cache, None)
            f = self.final_norm(xx.view(B, nq + 1, D)[:, 1:].reshape(B * nq, D).contiguous())
            route, speed = self.heads(f, B, nq * D)
        else:
            # sequences ended at different lengths: cache the final sampled token for everybody, then append the
            # queries group-wise at each group's own offset (rows beyond a sequence's end are overwritten / never
            # visible under the causal mask)
            self.llm_chunk(gen_embeds[steps - 1].clone(), B, 1, L + steps - 1, cache, None)
            for G in sorted(set(n_gen_cpu)):
                idx = [b for b in range(B) if n_gen_cpu[b] == G]
                ii = torch.tensor(idx, device=self.dev)
                sub = (cache[0][:, ii].contiguous(), cache[1][:, ii].contiguous())
                q = self.queries.unsqueeze(0).expand(len(idx), nq, D).reshape(len(idx) * nq, D).clone()
                xx = self.llm_chunk(q, len(idx), nq, L + G, sub, None)
                r, sp = self.heads(self.final_norm(xx), len(idx), nq * D)
                route[ii], speed[ii] = r, sp
        return speed, route, toks

    @torch.no_grad()
    def driving_forward(self, camera_images: Tensor, phrase_ids: Tensor, phrase_valid: Tensor, placeholder_values,
                        max_new_tokens: int = 100, eos_token_id: Optional[int] = None, margins: Optional[list] = None,
                        ids_cpu: Optional[Tensor] = None):
        emb = self.embed_prompt(phrase_ids, camera_images, placeholder_values, ids_cpu)
        return self.generate(emb, phrase_valid, max_new_tokens, eos_token_id, margins)

    # ------------------------------------------------------------------------------------------
    # teacher-forced single pass (DrivingModel.forward_model; BASELINE config 3 "offline batched forward")
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_model(self, inputs: Tensor, inputs_mask: Optional[Tensor], want_logits_rows: Optional[Tensor] = None):
        """inputs [B, L+30, 896] (already permuted "valid first"), mask [B, L+30] -> (features [B,L+30,896]
        after the final norm, fp32 logits for the requested flat row indices or None)."""
        s = self.spec
        B, Lt, D = inputs.shape
        cache = self.new_cache(B, Lt)
        kv_valid = None
        if inputs_mask is not None and not bool(inputs_mask.all()):
            kv_valid = torch.zeros((B, cache[0].shape[3]), device=self.dev, dtype=torch.uint8)
            kv_valid[:, :Lt] = inputs_mask.to(torch.uint8)
        x = self.llm_chunk(inputs.reshape(B * Lt, D).clone(), B, Lt, 0, cache, kv_valid)
        feats = self.final_norm(x)
        lg = None
        if want_logits_rows is not None:
            lg = self.logits(feats[want_logits_rows].contiguous())
        return feats.view(B, Lt, D), lg
