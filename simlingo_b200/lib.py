"""ctypes binding of ``libsimlingo_b200.so`` (C ABI declared in ``include/simlingo_b200.h``).

This is the only way the Python side reaches the GPU math: tensors are passed as raw device
pointers + sizes, the launch stream is ``torch.cuda.current_stream()``.  There is no fallback: if
the shared library is missing or a call returns non-zero a ``RuntimeError`` is raised."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libsimlingo_b200.so"
_lib: Optional[C.CDLL] = None

ACT_NONE, ACT_GELU, ACT_SILU, ACT_RELU = 0, 1, 2, 3


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("bias", C.c_void_p), ("scale_n", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("alpha", C.c_float), ("act", C.c_int32), ("swiglu", C.c_int32), ("out_fp32", C.c_int32),
        ("a_t", C.c_int32), ("b_t", C.c_int32), ("block_n", C.c_int32),
        ("rms_weight", C.c_void_p), ("rms_eps", C.c_float),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64), ("aux_mode", C.c_int32),
        ("A2", C.c_void_p), ("lda2", C.c_int64), ("K2", C.c_int32), ("a_fp32", C.c_int32),
    ]


class DecodeArgs(C.Structure):
    """slb_decode_args (include/simlingo_b200.h)"""
    _fields_ = [
        ("layers", C.c_void_p),
        ("n_layers", C.c_int32), ("batch", C.c_int32), ("hidden", C.c_int32), ("mlp", C.c_int32), ("vocab", C.c_int32),
        ("hq", C.c_int32), ("hkv", C.c_int32), ("lmax", C.c_int32), ("n_steps", C.c_int32), ("max_new", C.c_int32),
        ("emb_rows", C.c_int64), ("eos", C.c_int64), ("ld_sampled", C.c_int64),
        ("emb", C.c_void_p), ("norm_w", C.c_void_p), ("lm_head", C.c_void_p),
        ("kcache", C.c_void_p), ("vcache", C.c_void_p),
        ("pos", C.c_void_p), ("nxt", C.c_void_p), ("sampled", C.c_void_p), ("step", C.c_void_p), ("done", C.c_void_p), ("n_gen", C.c_void_p),
        ("rope_theta", C.c_float), ("rms_eps", C.c_float),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("status", C.c_void_p),
    ]


class WgradProblem(C.Structure):
    """slb_wgrad_problem (include/simlingo_b200.h)"""
    _fields_ = [("P", C.c_void_p), ("Q", C.c_void_p), ("out", C.c_void_p), ("ldp", C.c_int64), ("ldq", C.c_int64), ("ldo", C.c_int64),
                ("mo", C.c_int32), ("no", C.c_int32), ("alpha", C.c_float), ("accumulate", C.c_int32)]


class HeadsWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("r0w", "r0b", "r2w", "r2b", "r4w", "s0w", "s0b", "s2w")]


class WpWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "w2", "b2", "w4", "b4")]


def load() -> C.CDLL:
    """dlopen the library (building is ``python -m simlingo_b200.build`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (run `python simlingo_b200/build.py`). "
            "There is no CPU fallback.")
    # the torch-bundled NCCL / cudart must be resolvable before our library is opened
    try:
        import torch  # noqa: F401  (loads libcudart / libnccl into the process)
        nccl_dir = Path(torch.__file__).resolve().parent.parent / "nvidia" / "nccl" / "lib"
        so = nccl_dir / "libnccl.so.2"
        if so.exists():
            C.CDLL(str(so), mode=C.RTLD_GLOBAL)
    except OSError:
        pass
    _lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    _lib.slb_last_error.restype = C.c_char_p
    return _lib


LAUNCHES = 0  # kernels launched through this binding (one per call unless stated; read as deltas by the engines)


def _check(rc: int, what: str, n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n
    if rc != 0:
        msg = load().slb_last_error()
        raise RuntimeError(f"simlingo_b200: {what} failed ({rc}): {msg.decode() if msg else ''}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _bf16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.bfloat16, (t.device, t.dtype)


# ------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, bias=None, scale_n=None,
         residual=None, alpha: float = 1.0, act: int = ACT_NONE, swiglu: bool = False, out_fp32: bool = False,
         a_t: bool = False, b_t: bool = False, block_n: int = 0, rms_weight=None, rms_eps: float = 0.0, aux=None,
         aux_mode: int = 0, a2=None) -> torch.Tensor:
    """out[M,N] = epilogue(alpha * A @ B^T).  A: [M,K] (or [K,M] if a_t), B: [N,K] (or [K,N] if b_t);
    2-D, unit inner stride, arbitrary (multiple-of-8) row stride.  A may be fp32 when ``rms_weight`` is given (M <= 4)."""
    a_fp32 = a.dtype == torch.float32
    if a_fp32:
        assert rms_weight is not None and a.is_cuda and a.shape[0] <= 4, "fp32 A rows are only taken by the fused RMSNorm prologue (M <= 4)"
    _bf16(None if a_fp32 else a, b, bias, scale_n)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_t else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_t else (b.shape[0], b.shape[1])
    K2 = 0
    if a2 is not None:   # second A source appended along k: b holds K + K2 columns
        _bf16(a2)
        assert not a_t and not b_t and a2.dim() == 2 and a2.shape[0] == M and a2.stride(1) == 1
        K2 = a2.shape[1]
    assert K + K2 == Kb, (a.shape, b.shape, a_t, b_t, K2)
    n_out = N // 2 if swiglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape == (M, n_out)
    assert out.dtype == (torch.float32 if out_fp32 else torch.bfloat16)
    if residual is not None:
        assert residual.dtype == out.dtype and residual.stride(1) == 1 and residual.shape == out.shape
    g = GemmArgs(M, N, K, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0),
                 None if bias is None else bias.data_ptr(), None if scale_n is None else scale_n.data_ptr(),
                 None if residual is None else residual.data_ptr(), 0 if residual is None else residual.stride(0),
                 alpha, act, int(swiglu), int(out_fp32), int(a_t), int(b_t), block_n,
                 None if rms_weight is None else rms_weight.data_ptr(), rms_eps,
                 None if aux is None else aux.data_ptr(), 0 if aux is None else aux.stride(0), aux_mode,
                 None if a2 is None else a2.data_ptr(), 0 if a2 is None else a2.stride(0), K2, int(a_fp32))
    if aux is not None:
        assert aux.dtype == torch.bfloat16 and aux.shape == (M, n_out) and aux.stride(1) == 1
    _check(load().slb_gemm_bf16(C.byref(g), _stream()), "gemm_bf16")
    return out


def layernorm(x, w, b, eps, out=None, stats=None):
    rows, cols = x.shape[0], x.shape[1]
    if x.dtype == torch.float32:   # fp32 residual stream (InternViT inference path)
        _bf16(w, b)
        assert x.is_cuda and x.is_contiguous() and stats is None
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if out is None else out
        _check(load().slb_layernorm_fwd_f32(_p(x), _p(w), _p(b), _p(out), rows, cols, C.c_float(eps), _stream()), "layernorm_fwd_f32")
        return out
    _bf16(x, w, b)
    out = torch.empty_like(x) if out is None else out
    mean, rstd = (stats if stats is not None else (None, None))
    _check(load().slb_layernorm_fwd(_p(x), _p(w), _p(b), _p(out), rows, cols, C.c_float(eps), _p(mean), _p(rstd), _stream()),
           "layernorm_fwd")
    return out


def rmsnorm(x, w, eps, out=None, rstd=None):
    """x bf16 or fp32 [rows, cols] (fp32: the residual stream of the Qwen2 inference path) -> bf16"""
    _bf16(w)
    if x.dtype == torch.float32:
        assert x.is_cuda and x.is_contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if out is None else out
        _check(load().slb_rmsnorm_fwd_f32(_p(x), _p(w), _p(out), x.shape[0], x.shape[1], C.c_float(eps), _p(rstd), _stream()), "rmsnorm_fwd_f32")
        return out
    _bf16(x)
    out = torch.empty_like(x) if out is None else out
    _check(load().slb_rmsnorm_fwd(_p(x), _p(w), _p(out), x.shape[0], x.shape[1], C.c_float(eps), _p(rstd), _stream()),
           "rmsnorm_fwd")
    return out


def im2col_patch(pixels, kpad=640, out=None):
    _bf16(pixels)
    tiles = pixels.shape[0]
    assert pixels.shape[1:] == (3, 448, 448) and pixels.is_contiguous()
    out = torch.empty((tiles * 1024, kpad), device=pixels.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_im2col_patch(_p(pixels), _p(out), tiles, kpad, _stream()), "im2col_patch")
    return out


def patch_embed(pixels, weight_padded, bias, cls, pos, fp32=True, out=None):
    """implicit-GEMM patch embedding: pixels bf16 [T,3,448,448], weight_padded bf16 [1024, >= 640] -> token stream [T*1025, 1024]"""
    _bf16(pixels, weight_padded, bias, cls, pos)
    tiles = pixels.shape[0]
    assert pixels.shape[1:] == (3, 448, 448) and pixels.is_contiguous() and weight_padded.shape[0] == 1024 and weight_padded.stride(1) == 1
    out = torch.empty((tiles * 1025, 1024), device=pixels.device, dtype=torch.float32 if fp32 else torch.bfloat16) if out is None else out
    assert out.is_contiguous() and out.dtype == (torch.float32 if fp32 else torch.bfloat16)
    _check(load().slb_patch_embed(_p(pixels), _p(weight_padded), C.c_int64(weight_padded.stride(0)), _p(bias), _p(cls.contiguous()), _p(pos.contiguous()),
                                  _p(out), tiles, int(fp32), _stream()), "patch_embed")
    return out


def vit_assemble(patch_out, cls, pos, tiles, out=None, fp32=False):
    _bf16(patch_out, cls, pos)
    if fp32:
        out = torch.empty((tiles * 1025, 1024), device=patch_out.device, dtype=torch.float32) if out is None else out
        _check(load().slb_vit_assemble_f32(_p(patch_out), _p(cls), _p(pos), _p(out), tiles, _stream()), "vit_assemble_f32")
        return out
    out = torch.empty((tiles * 1025, 1024), device=patch_out.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_vit_assemble(_p(patch_out), _p(cls), _p(pos), _p(out), tiles, _stream()), "vit_assemble")
    return out


def pixel_shuffle_ln(x, w, b, tiles, eps, out=None, stats=None):
    out = torch.empty((tiles * 256, 4096), device=x.device, dtype=torch.bfloat16) if out is None else out
    if x.dtype == torch.float32:
        _bf16(w, b)
        assert x.is_cuda and x.is_contiguous() and stats is None
        _check(load().slb_pixel_shuffle_ln_f32(_p(x), _p(w), _p(b), _p(out), tiles, C.c_float(eps), _stream()), "pixel_shuffle_ln_f32")
        return out
    _bf16(x, w, b)
    mean, rstd = (stats if stats is not None else (None, None))
    _check(load().slb_pixel_shuffle_ln(_p(x), _p(w), _p(b), _p(out), tiles, C.c_float(eps), _p(mean), _p(rstd), _stream()),
           "pixel_shuffle_ln")
    return out


def attn_vit(qkv, tiles, n_tokens, heads=16, out=None, lse=None):
    _bf16(qkv)
    assert qkv.is_contiguous() and qkv.shape == (tiles * n_tokens, 3 * heads * 64)
    out = torch.empty((tiles * n_tokens, heads * 64), device=qkv.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_attn_vit_fwd(_p(qkv), _p(out), _p(lse), tiles, n_tokens, heads, _stream()), "attn_vit_fwd")
    return out


def attn_gqa(q, ldq, kcache, vcache, batch, lq, past, hq=14, hkv=2, key_valid=None, out=None, lse=None, past_dev=None):
    """q: view into the fused qkv buffer ([B*Lq, ldq] rows, first hq*64 columns are q).
    past_dev: int32 CUDA tensor [1] holding the chunk position (overrides ``past`` on the device)."""
    _bf16(q, kcache, vcache)
    lmax = kcache.shape[2]
    assert kcache.shape == (batch, hkv, lmax, 64) and kcache.is_contiguous() and vcache.is_contiguous()
    out = torch.empty((batch * lq, hq * 64), device=q.device, dtype=torch.bfloat16) if out is None else out
    kv_ld = 0
    ws, ws_bytes = None, 0
    if key_valid is not None:
        assert key_valid.dtype == torch.uint8 and key_valid.dim() == 2 and key_valid.stride(1) == 1
        kv_ld = key_valid.stride(0)
        if lq >= 128 and past == 0 and past_dev is None:   # masked prefill: validity words per key block
            fn = load().slb_attn_gqa_fwd_workspace
            fn.restype = C.c_size_t
            ws_bytes = int(fn(batch, lq))
            ws = _workspace(ws_bytes, q.device)
    _check(load().slb_attn_gqa_fwd(_p(q), C.c_int64(ldq), _p(kcache), _p(vcache), _p(key_valid), kv_ld, _p(out), _p(lse),
                                   batch, lq, past, _p(past_dev), lmax, hq, hkv, _p(ws), C.c_size_t(ws_bytes), _stream()), "attn_gqa_fwd",
           2 if ws is not None else 1)
    return out


def attn_decode_rope(qkv, kcache, vcache, batch, past, hq=14, hkv=2, theta=1.0e6, key_valid=None, out=None, past_dev=None):
    """one decode step in one launch: RoPE on q / the new key, KV-cache write at ``past`` and attention over positions 0..past"""
    _bf16(qkv, kcache, vcache)
    lmax = kcache.shape[2]
    assert qkv.is_contiguous() and qkv.shape == (batch, (hq + 2 * hkv) * 64)
    assert kcache.shape == (batch, hkv, lmax, 64) and kcache.is_contiguous() and vcache.is_contiguous()
    out = torch.empty((batch, hq * 64), device=qkv.device, dtype=torch.bfloat16) if out is None else out
    kv_ld = 0
    if key_valid is not None:
        assert key_valid.dtype == torch.uint8 and key_valid.dim() == 2 and key_valid.stride(1) == 1
        kv_ld = key_valid.stride(0)
    _check(load().slb_attn_decode_rope(_p(qkv), C.c_int64(qkv.stride(0)), _p(kcache), _p(vcache), _p(key_valid), kv_ld, _p(out), batch, past,
                                       _p(past_dev), lmax, hq, hkv, C.c_float(theta), _stream()), "attn_decode_rope")
    return out


def rope_kv_write(qkv, kcache, vcache, batch, lq, past, hq=14, hkv=2, theta=1.0e6, past_dev=None):
    _bf16(qkv, kcache, vcache)
    assert qkv.is_contiguous() and qkv.shape == (batch * lq, (hq + 2 * hkv) * 64)
    lmax = kcache.shape[2]
    _check(load().slb_rope_kv_write(_p(qkv), _p(kcache), _p(vcache), batch, lq, past, _p(past_dev), lmax, hq, hkv, C.c_float(theta), _stream()),
           "rope_kv_write")


def embed_assemble(ids, table, vit, wp, wp_start, wp_len, img_id, n_img, out=None):
    """ids int64 [B,L]; table bf16 [V,H]; vit bf16 [B*n_img,H] or None; wp bf16 [B,wp_len,H] or None;
    wp_start int32 [B] (-1 = none)."""
    assert ids.dtype == torch.int64 and ids.is_cuda and ids.is_contiguous()
    B, L = ids.shape
    V, H = table.shape
    out = torch.empty((B, L, H), device=ids.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_embed_assemble(_p(ids), _p(table), _p(vit), _p(wp), _p(wp_start), wp_len, _p(out), B, L, H, V, img_id, n_img,
                                     _stream()), "embed_assemble")
    return out


def gather_rows(src, idx, out=None):
    _bf16(src)
    assert idx.dtype == torch.int64 and idx.is_cuda
    n = idx.numel()
    out = torch.empty((n, src.shape[1]), device=src.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_gather_rows(_p(src), _p(idx), _p(out), n, src.shape[1], C.c_int64(src.shape[0]), _stream()), "gather_rows")
    return out


def scatter_rows(dst, idx, src):
    """dst[idx[i]] = src[i] in place; dst must be a contiguous 2-D bf16 view."""
    _bf16(dst, src)
    assert dst.is_contiguous() and src.is_contiguous() and idx.dtype == torch.int64 and idx.is_cuda
    _check(load().slb_scatter_rows(_p(dst), _p(idx.contiguous()), _p(src), idx.numel(), dst.shape[1], C.c_int64(dst.shape[0]), _stream()),
           "scatter_rows")
    return dst


def silu_mul(g, u, out=None):
    _bf16(g, u)
    out = torch.empty_like(g) if out is None else out
    _check(load().slb_silu_mul(_p(g), _p(u), _p(out), C.c_int64(g.numel()), _stream()), "silu_mul")
    return out


def add(a, b, out=None):
    _bf16(a, b)
    out = torch.empty_like(a) if out is None else out
    _check(load().slb_add_bf16(_p(a), _p(b), _p(out), C.c_int64(a.numel()), _stream()), "add_bf16")
    return out


def argmax(logits, out_idx=None, out_margin=None):
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    rows, cols = logits.shape
    out_idx = torch.empty((rows,), device=logits.device, dtype=torch.int64) if out_idx is None else out_idx
    _check(load().slb_argmax_f32(_p(logits), C.c_int64(logits.stride(0)), rows, cols, _p(out_idx), _p(out_margin), _stream()),
           "argmax_f32")
    return out_idx


def decode_workspace(batch: int, hidden: int, mlp: int, hq: int, hkv: int, device) -> torch.Tensor:
    """scratch of the persistent decode kernel (residual stream, q|k|v, attention partials, arg-max partials, barrier counters)"""
    f = load().slb_decode_workspace_bytes
    f.restype = C.c_size_t
    n = int(f(batch, hidden, mlp, hq, hkv))
    if n <= 0:
        raise RuntimeError("simlingo_b200: decode_workspace_bytes rejected the shape")
    return torch.zeros((n + 256,), device=device, dtype=torch.uint8)


def decode_layer_table(layers, device) -> torch.Tensor:
    """device array of slb_decode_layer (7 pointers per layer) from the engine's folded per-layer weights"""
    rows = []
    for ly in layers:
        ts = [ly[k] for k in ("qkv", "bqkv", "o", "gu", "d", "ln1", "ln2")]
        for t in ts:
            assert t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()
        rows.append([t.data_ptr() for t in ts])
    return torch.tensor(rows, dtype=torch.int64).to(device)


def decode_loop(table, n_layers, batch, hidden, mlp, hq, hkv, emb, norm_w, lm_head, kcache, vcache, pos, nxt, sampled, step, done, n_gen,
                status, workspace, n_steps, eos, theta, eps):
    """greedy decode of up to ``n_steps`` tokens in ONE persistent launch (slb_decode_loop; llm.py:217-248 per token).  All state
    tensors live on the device and are updated in place; stops early once every ``done`` flag is set (``eos`` not None)."""
    _bf16(emb, norm_w, lm_head, kcache, vcache)
    assert table.dtype == torch.int64 and table.is_cuda and table.shape == (n_layers, 7) and table.is_contiguous()
    assert kcache.dim() == 5 and kcache.shape[:3] == (n_layers, batch, hkv) and kcache.shape[4] == 64 and kcache.is_contiguous()
    assert vcache.shape == kcache.shape and vcache.is_contiguous()
    assert emb.is_contiguous() and lm_head.is_contiguous() and emb.shape[1] == hidden and lm_head.shape[1] == hidden
    assert pos.dtype == torch.int32 and step.dtype == torch.int64 and nxt.dtype == torch.int64 and nxt.numel() == batch
    assert sampled.dtype == torch.int64 and sampled.dim() == 2 and sampled.shape[0] == batch and sampled.stride(1) == 1
    assert done.dtype in (torch.bool, torch.uint8) and done.numel() == batch and n_gen.dtype == torch.int64 and n_gen.numel() == batch
    assert status.dtype in (torch.int32, torch.int64) and status.is_cuda and workspace.dtype == torch.uint8
    ws_ptr = (workspace.data_ptr() + 255) // 256 * 256
    a = DecodeArgs()
    a.layers = table.data_ptr()
    a.n_layers, a.batch, a.hidden, a.mlp, a.vocab, a.hq, a.hkv = n_layers, batch, hidden, mlp, lm_head.shape[0], hq, hkv
    a.lmax, a.n_steps, a.max_new = kcache.shape[3], n_steps, sampled.shape[1]
    a.emb_rows, a.eos, a.ld_sampled = emb.shape[0], (-1 if eos is None else int(eos)), sampled.stride(0)
    a.emb, a.norm_w, a.lm_head = emb.data_ptr(), norm_w.data_ptr(), lm_head.data_ptr()
    a.kcache, a.vcache = kcache.data_ptr(), vcache.data_ptr()
    a.pos, a.nxt, a.sampled, a.step, a.done, a.n_gen = (pos.data_ptr(), nxt.data_ptr(), sampled.data_ptr(), step.data_ptr(),
                                                          done.data_ptr(), n_gen.data_ptr())
    a.rope_theta, a.rms_eps = theta, eps
    a.workspace, a.workspace_bytes = ws_ptr, workspace.numel() - (ws_ptr - workspace.data_ptr())
    a.status = status.data_ptr()
    _check(load().slb_decode_loop(C.byref(a), _stream()), "decode_loop")


def argmax_sample(logits, nxt, sampled, pos, base, done, n_gen, step, eos):
    """argmax of every logits row + the bookkeeping of ``greedy_sample`` (llm.py:232-248) in one launch; all state on the device"""
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    rows, cols = logits.shape
    assert nxt.dtype == torch.int64 and nxt.numel() == rows and sampled.dtype == torch.int64 and sampled.shape[0] == rows and sampled.stride(1) == 1
    assert pos.dtype == torch.int32 and step.dtype == torch.int64 and n_gen.dtype == torch.int64 and done.dtype in (torch.bool, torch.uint8)
    _check(load().slb_argmax_sample(_p(logits), C.c_int64(logits.stride(0)), rows, cols, _p(nxt), _p(sampled), C.c_int64(sampled.stride(0)),
                                    sampled.shape[1], _p(pos), int(base), _p(done), _p(n_gen), _p(step),
                                    C.c_int64(-1 if eos is None else int(eos)), _stream()), "argmax_sample")
    return nxt


def driving_heads(feats, ld_batch, hw: HeadsWeights, batch, route=None, speed=None, ws=None):
    dev = feats.device
    route = torch.empty((batch, 20, 2), device=dev, dtype=torch.float32) if route is None else route
    speed = torch.empty((batch, 10, 2), device=dev, dtype=torch.float32) if speed is None else speed
    ws = torch.empty((batch, 30, 2), device=dev, dtype=torch.float32) if ws is None else ws
    _check(load().slb_driving_heads(_p(feats), C.c_int64(ld_batch), C.byref(hw), _p(route), _p(speed), _p(ws), batch, _stream()),
           "driving_heads")
    return route, speed


def wp_encoder(coords, ww: WpWeights, out=None):
    assert coords.dtype == torch.float32 and coords.is_cuda and coords.is_contiguous()
    n = coords.shape[0]
    out = torch.empty((n, 896), device=coords.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_wp_encoder(_p(coords), C.byref(ww), _p(out), n, _stream()), "wp_encoder")
    return out


# ------------------------------------------------------------------------------------------------
# training-only ops
# ------------------------------------------------------------------------------------------------
def _f32(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), (t.device, t.dtype)


def layernorm_bwd(dy, x, w, mean, rstd, dw_acc, db_acc, dx=None, add=None):
    """add: gradient arriving over the residual connection (bf16, same shape), summed into dx before its one rounding"""
    _bf16(dy, x, w, add); _f32(mean, rstd, dw_acc, db_acc)
    dx = torch.empty_like(x) if dx is None else dx
    assert add is None or (add.is_contiguous() and add.shape == x.shape)
    _check(load().slb_layernorm_bwd(_p(dy), _p(x), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw_acc), _p(db_acc), x.shape[0], x.shape[1], _p(add), _stream()),
           "layernorm_bwd")
    return dx


def rmsnorm_bwd(dy, x, w, rstd, dx=None, add=None):
    _bf16(dy, x, w, add); _f32(rstd)
    dx = torch.empty_like(x) if dx is None else dx
    assert add is None or (add.is_contiguous() and add.shape == x.shape)
    _check(load().slb_rmsnorm_bwd(_p(dy), _p(x), _p(w), _p(rstd), _p(dx), None, x.shape[0], x.shape[1], _p(add), _stream()), "rmsnorm_bwd")
    return dx


def pixel_shuffle_ln_bwd(dy, x, w, mean, rstd, dw_acc, db_acc, tiles, dx=None):
    _bf16(dy, x, w); _f32(mean, rstd, dw_acc, db_acc)
    dx = torch.empty_like(x) if dx is None else dx
    _check(load().slb_pixel_shuffle_ln_bwd(_p(dy), _p(x), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw_acc), _p(db_acc), tiles, _stream()),
           "pixel_shuffle_ln_bwd")
    return dx


def gelu_fwd(x, out=None):
    _bf16(x)
    out = torch.empty_like(x) if out is None else out
    _check(load().slb_gelu_fwd(_p(x), _p(out), C.c_int64(x.numel()), _stream()), "gelu_fwd")
    return out


def gelu_bwd(pre, dout, out=None):
    _bf16(pre, dout)
    out = torch.empty_like(pre) if out is None else out
    _check(load().slb_gelu_bwd(_p(pre), _p(dout), _p(out), C.c_int64(pre.numel()), _stream()), "gelu_bwd")
    return out


def silu_mul_bwd(g, u, dout, dg=None, du=None):
    _bf16(g, u, dout)
    dg = torch.empty_like(g) if dg is None else dg
    du = torch.empty_like(u) if du is None else du
    _check(load().slb_silu_mul_bwd(_p(g), _p(u), _p(dout), _p(dg), _p(du), C.c_int64(g.numel()), _stream()), "silu_mul_bwd")
    return dg, du


def dropout(x, p, seed, out=None, seed_dev=None):
    """seed_dev: optional int64 CUDA tensor [1] (step counter read on the device, see slb_dropout)."""
    _bf16(x)
    out = torch.empty_like(x) if out is None else out
    _check(load().slb_dropout(_p(x), _p(out), C.c_int64(x.numel()), C.c_float(p), C.c_uint64(seed), _p(seed_dev), _stream()), "dropout")
    return out


def dropout_add(x, y, p, seed, seed_dev=None):
    """y += dropout(x) (same mask as ``dropout(.., p, seed, seed_dev)``)."""
    _bf16(x, y)
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
    _check(load().slb_dropout_add(_p(x), _p(y), C.c_int64(x.numel()), C.c_float(p), C.c_uint64(seed), _p(seed_dev), _stream()), "dropout_add")
    return y


def dropout_multi(x, p, seeds, seed_dev=None):
    """-> list of len(seeds) independently masked copies of x (copy j == dropout(x, p, seeds[j], seed_dev)), one read of x"""
    _bf16(x)
    n = len(seeds)
    assert 1 <= n <= 4 and x.is_contiguous()
    outs = [torch.empty_like(x) for _ in range(n)]
    ys = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    sd = (C.c_uint64 * n)(*[int(s) for s in seeds])
    _check(load().slb_dropout_multi(_p(x), ys, sd, n, C.c_int64(x.numel()), C.c_float(p), _p(seed_dev), _stream()), "dropout_multi")
    return outs


def lora_pack(table, n_entries, rank, scale):
    """table: int64 CUDA tensor [n_entries, 4] = (src ptr, dst ptr, rows, dst row stride); dst[:, :rank] = scale * src"""
    assert table.dtype == torch.int64 and table.is_cuda and table.is_contiguous() and table.shape == (n_entries, 4)
    _check(load().slb_lora_pack(_p(table), n_entries, rank, C.c_float(scale), _stream()), "lora_pack")


def lora_dx(cat, K, a_list, p=0.0, seeds=None, seed_dev=None, out=None):
    """cat bf16 [M, K + r*n] (row stride free) = [base dgrad | dt_0 .. dt_{n-1}] -> out [M, K] = base + sum_j mask_j o (dt_j @ A_j) / (1-p)"""
    _bf16(cat, *a_list)
    n = len(a_list)
    r = a_list[0].shape[0]
    M = cat.shape[0]
    assert cat.stride(1) == 1 and cat.shape[1] >= K + r * n and all(a.shape == (r, K) and a.is_contiguous() for a in a_list)
    out = torch.empty((M, K), device=cat.device, dtype=torch.bfloat16) if out is None else out
    assert out.stride(1) == 1 and out.shape == (M, K)
    ap = (C.c_void_p * n)(*[a.data_ptr() for a in a_list])
    sd = None if seeds is None else (C.c_uint64 * n)(*[int(s) for s in seeds])
    _check(load().slb_lora_dx(_p(cat), C.c_int64(cat.stride(0)), _p(out), C.c_int64(out.stride(0)), ap, sd, n, M, K, r, C.c_float(p),
                              _p(seed_dev), _stream()), "lora_dx")
    return out


def lora_wgrad_grouped(problems, rows: int) -> None:
    """One launch for the parameter gradients of a group of adapters: for every (P [rows, mo], Q [rows, no], out [mo, no], alpha,
    accumulate) in ``problems`` (at most 16):  out (+)= alpha * P^T Q.  P / Q may be column slices of wider matrices."""
    n = len(problems)
    arr = (WgradProblem * n)()
    for k, (P, Q, out, alpha, acc) in enumerate(problems):
        _bf16(P, Q, out)
        assert P.dim() == 2 and Q.dim() == 2 and out.dim() == 2 and P.stride(1) == 1 and Q.stride(1) == 1 and out.stride(1) == 1
        assert P.shape[0] == rows and Q.shape[0] == rows and out.shape == (P.shape[1], Q.shape[1]), (P.shape, Q.shape, out.shape, rows)
        a = arr[k]
        a.P, a.Q, a.out = P.data_ptr(), Q.data_ptr(), out.data_ptr()
        a.ldp, a.ldq, a.ldo = P.stride(0), Q.stride(0), out.stride(0)
        a.mo, a.no, a.alpha, a.accumulate = P.shape[1], Q.shape[1], float(alpha), int(bool(acc))
    _check(load().slb_lora_wgrad_grouped(arr, n, rows, _stream()), "lora_wgrad_grouped")


def silu_mul_cat(gu, out=None):
    """gu bf16 [rows, 2I] = [gate | up] -> silu(gate) * up  [rows, I]"""
    _bf16(gu)
    rows, I2 = gu.shape
    assert gu.is_contiguous() and I2 % 16 == 0
    out = torch.empty((rows, I2 // 2), device=gu.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_silu_mul_cat(_p(gu), _p(out), rows, I2 // 2, _stream()), "silu_mul_cat")
    return out


def silu_mul_cat_bwd(gu, dout, out=None):
    """-> dgu bf16 [rows, 2I] = [dgate | dup]"""
    _bf16(gu, dout)
    rows, I2 = gu.shape
    assert gu.is_contiguous() and dout.is_contiguous() and dout.shape == (rows, I2 // 2)
    out = torch.empty_like(gu) if out is None else out
    _check(load().slb_silu_mul_cat_bwd(_p(gu), _p(dout), _p(out), rows, I2 // 2, _stream()), "silu_mul_cat_bwd")
    return out


def flush_f32(x, y, accumulate=False):
    """y (bf16) = [y +] x (fp32)"""
    _f32(x); _bf16(y)
    assert y.is_contiguous() and x.numel() == y.numel()
    _check(load().slb_flush_f32_to_bf16(_p(x), _p(y), C.c_int64(x.numel()), int(accumulate), _stream()), "flush_f32")
    return y


def add_inplace(a, b):
    _bf16(a, b)
    _check(load().slb_add_inplace_bf16(_p(a), _p(b), C.c_int64(a.numel()), _stream()), "add_inplace")
    return a


def scale_cols(x, s, out=None):
    _bf16(x, s)
    out = torch.empty_like(x) if out is None else out
    _check(load().slb_scale_cols(_p(x), _p(s), _p(out), x.shape[0], x.shape[1], _stream()), "scale_cols")
    return out


def scale_cols_add(x, s, res, out=None):
    _bf16(x, s, res)
    out = torch.empty_like(x) if out is None else out
    _check(load().slb_scale_cols_add(_p(x), _p(s), _p(res), _p(out), x.shape[0], x.shape[1], _stream()), "scale_cols_add")
    return out


def col_reduce(a, acc, b=None, alpha=1.0):
    """acc[c] += alpha * sum_r a[r,c] * (b[r,c] if b is given else 1); acc fp32 [cols]."""
    _bf16(a, b); _f32(acc)
    assert a.dim() == 2 and a.stride(1) == 1 and (b is None or (b.stride(1) == 1 and b.shape == a.shape))
    _check(load().slb_col_reduce(_p(a), C.c_int64(a.stride(0)), _p(b), C.c_int64(0 if b is None else b.stride(0)), _p(acc), a.shape[0], a.shape[1],
                                 C.c_float(alpha), _stream()), "col_reduce")
    return acc


def layerscale_bwd(dx, branch, ls, dls_acc, dbias_acc, out=None):
    """-> dbranch = dx * ls (bf16); dls_acc += colsum(dx * branch); dbias_acc += colsum(dbranch)"""
    _bf16(dx, branch, ls); _f32(dls_acc, dbias_acc)
    assert dx.is_contiguous() and branch.is_contiguous() and dx.shape == branch.shape
    out = torch.empty_like(dx) if out is None else out
    _check(load().slb_layerscale_bwd(_p(dx), _p(branch), _p(ls), _p(out), _p(dls_acc), _p(dbias_acc), dx.shape[0], dx.shape[1], _stream()),
           "layerscale_bwd")
    return out


def vit_assemble_bwd(dx, dcls_acc, dpos_acc, tiles, dpatch=None):
    _bf16(dx); _f32(dcls_acc, dpos_acc)
    dpatch = torch.empty((tiles * 1024, 1024), device=dx.device, dtype=torch.bfloat16) if dpatch is None else dpatch
    _check(load().slb_vit_assemble_bwd(_p(dx), _p(dpatch), _p(dcls_acc), _p(dpos_acc), tiles, _stream()), "vit_assemble_bwd")
    return dpatch


def rope_bwd(dq, dk, dv, batch, lq, hq, hkv, theta, out=None):
    _f32(dq, dk, dv)
    out = torch.empty((batch * lq, (hq + 2 * hkv) * 64), device=dq.device, dtype=torch.bfloat16) if out is None else out
    _check(load().slb_rope_bwd(_p(dq), _p(dk), _p(dv), _p(out), batch, lq, hq, hkv, C.c_float(theta), _stream()), "rope_bwd")
    return out


def attn_delta(o, dout, batch, lq, heads, out=None):
    _bf16(o, dout)
    out = torch.empty((batch, heads, lq), device=o.device, dtype=torch.float32) if out is None else out
    _check(load().slb_attn_delta(_p(o), _p(dout), _p(out), batch, lq, heads, _stream()), "attn_delta")
    return out


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per device (callers of the C ABI own all memory; kernels on one stream serialise)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream if torch.cuda.is_current_stream_capturing() else 0)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, device=device, dtype=torch.uint8)
        _ws_cache[key] = buf
    return buf


def attn_bwd_workspace(batch, lq, hq) -> int:
    fn = load().slb_attn_bwd_workspace
    fn.restype = C.c_size_t
    return int(fn(batch, lq, hq))


def attn_vit_bwd(qkv, dout, lse, delta, tiles, n_tokens, heads=16, out=None, workspace=None):
    """-> dqkv bf16 [T*n, 3*H*64] (dQ | dK | dV packed like qkv)"""
    _bf16(qkv, dout); _f32(lse, delta)
    out = torch.empty_like(qkv) if out is None else out
    nb = attn_bwd_workspace(tiles, n_tokens, heads)
    ws = _workspace(nb, qkv.device) if workspace is None else workspace
    _check(load().slb_attn_vit_bwd(_p(qkv), _p(dout), _p(lse), _p(delta), _p(out), _p(ws), C.c_size_t(ws.numel() * ws.element_size()), tiles, n_tokens,
                                   heads, _stream()), "attn_vit_bwd", 2)
    return out


def attn_gqa_bwd(q, ldq, kcache, vcache, dout, lse, delta, batch, lq, hq=14, hkv=2, key_valid=None, workspace=None):
    """-> (dq fp32 [B*L, Hq*64] (post-RoPE space), dk fp32 [B,Hkv,L,64], dv fp32 [B,Hkv,L,64])"""
    _bf16(q, kcache, vcache, dout); _f32(lse, delta)
    dev = q.device
    lmax = kcache.shape[2]
    dq = torch.empty((batch * lq, hq * 64), device=dev, dtype=torch.float32)
    dk = torch.empty((batch, hkv, lq, 64), device=dev, dtype=torch.float32)
    dv = torch.empty_like(dk)
    kv_ld = 0 if key_valid is None else key_valid.stride(0)
    nb = attn_bwd_workspace(batch, lq, hq)
    ws = _workspace(nb, dev) if workspace is None else workspace
    _check(load().slb_attn_gqa_bwd(_p(q), C.c_int64(ldq), _p(kcache), _p(vcache), _p(key_valid), kv_ld, _p(dout), _p(lse), _p(delta), _p(dq),
                                   _p(dk), _p(dv), _p(ws), C.c_size_t(ws.numel() * ws.element_size()), batch, lq, lmax, hq, hkv, _stream()), "attn_gqa_bwd", 2)
    return dq, dk, dv


def ce_fwd_bwd(logits, labels, grad_scale=1.0, want_grad=True):
    """logits fp32 [R, V]; labels int64 [R] (<0 = ignore) -> (loss fp32 [R], dlogits bf16 [R, Vpad] or None)"""
    _f32(logits)
    R, V = logits.shape
    loss = torch.empty((R,), device=logits.device, dtype=torch.float32)
    dl, ldd = None, 0
    if want_grad:
        ldd = (V + 63) // 64 * 64
        dl = torch.empty((R, ldd), device=logits.device, dtype=torch.bfloat16)
    _check(load().slb_ce_fwd_bwd(_p(logits), C.c_int64(logits.stride(0)), _p(labels.contiguous()), _p(loss), _p(dl), C.c_int64(ldd),
                                 C.c_float(grad_scale), R, V, _stream()), "ce_fwd_bwd")
    return loss, dl


def grad_sqnorm(g, out):
    _bf16(g); _f32(out)
    _check(load().slb_grad_sqnorm(_p(g), C.c_int64(g.numel()), _p(out), _stream()), "grad_sqnorm")
    return out


def adamw_fused(master, m, v, grad, param, lr, beta1, beta2, eps, wd, step, sqnorm=None, max_norm=0.0, prescale=1.0):
    _f32(master, m, v, sqnorm); _bf16(grad, param)
    _check(load().slb_adamw_fused(_p(master), _p(m), _p(v), _p(grad), _p(param), C.c_int64(master.numel()), C.c_float(lr), C.c_float(beta1),
                                  C.c_float(beta2), C.c_float(eps), C.c_float(wd), step, _p(sqnorm), C.c_float(max_norm), C.c_float(prescale),
                                  _stream()), "adamw_fused")
