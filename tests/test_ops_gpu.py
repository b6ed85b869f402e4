"""Per-kernel numerics on a B200: every C-ABI op against a plain PyTorch fp32 reference of the same op
(computed on the GPU from the same bf16 inputs).  Tolerances are stated per test; bf16 outputs are
compared with max-abs error relative to the reference's max magnitude."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def lib():
    from simlingo_b200 import lib as L
    L.load()
    return L


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


GEMM_CASES = [
    # M, N, K, block_n
    (128, 128, 64, 128), (256, 256, 128, 256), (300, 200, 72, 0), (2050, 3072, 1024, 0), (2050, 1024, 4096, 128),
    (2050, 1024, 1024, 256), (545, 1152, 896, 0), (545, 896, 4864, 0), (2048, 1024, 640, 0), (545, 32, 896, 128),
    (545, 896, 32, 0), (7, 896, 896, 0), (1, 1152, 896, 0),
    (4728, 32, 896, 0), (4728, 64, 896, 0), (2050, 32, 4864, 32), (1500, 48, 896, 64),   # narrow tiles (LoRA down-projections)
]


@pytest.mark.parametrize("M,N,K,bn", GEMM_CASES)
def test_gemm_plain(lib, M, N, K, bn):
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2)
    out = lib.gemm(a, b, block_n=bn)
    ref = a.float() @ b.float().t()
    assert relerr(out, ref) < 1e-2  # bf16 output rounding (2^-9) + fp32 accumulation order


@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 64, 2256), (512, 512, 256, 2256), (2050, 3072, 1024, 2256), (2050, 1024, 4096, 2256),
                                      (300, 200, 72, 2256), (4728, 896, 4864, 2224), (4728, 1152, 896, 2192), (131, 896, 896, 2224),
                                      (16400, 4096, 1024, 2256), (7, 1152, 896, 2192)])
def test_gemm_two_cta(lib, M, N, K, bn):
    """cta_group::2 kernel (cluster of two CTAs per 256 x BN tile)."""
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2)
    bias, res = rnd(N, seed=3), rnd(M, N, seed=5)
    out = lib.gemm(a, b, block_n=bn)
    ref = a.float() @ b.float().t()
    assert relerr(out, ref) < 1e-2
    out = lib.gemm(a, b, bias=bias, residual=res, block_n=bn)
    assert relerr(out, ref + bias.float() + res.float()) < 1e-2


def test_gemm_two_cta_swiglu(lib):
    M, K, I = 1200, 896, 4864
    a = rnd(M, K, seed=1)
    wg, wu = rnd(I, K, seed=2, scale=0.05), rnd(I, K, seed=3, scale=0.05)
    w = torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], dim=1).reshape(2 * I, K).contiguous()
    out = lib.gemm(a, w, swiglu=True, block_n=2256)
    ref = F.silu(a.float() @ wg.float().t()) * (a.float() @ wu.float().t())
    assert relerr(out, ref) < 1e-2


def test_gemm_epilogues(lib):
    M, N, K = 2050, 1024, 1024
    a, b = rnd(M, K, seed=1, scale=0.5), rnd(N, K, seed=2, scale=0.05)
    bias, ls, res = rnd(N, seed=3), rnd(N, seed=4, scale=0.2), rnd(M, N, seed=5)
    acc = a.float() @ b.float().t()
    out = lib.gemm(a, b, bias=bias)
    assert relerr(out, acc + bias.float()) < 1e-2
    out = lib.gemm(a, b, bias=bias, act=lib.ACT_GELU)
    assert relerr(out, F.gelu(acc + bias.float())) < 1e-2
    out = lib.gemm(a, b, bias=bias, scale_n=ls, residual=res)
    assert relerr(out, res.float() + ls.float() * (acc + bias.float())) < 1e-2
    out = lib.gemm(a, b, residual=res, alpha=2.0)
    assert relerr(out, res.float() + 2.0 * acc) < 1e-2
    out = lib.gemm(a, b, act=lib.ACT_SILU)
    assert relerr(out, F.silu(acc)) < 1e-2
    out = lib.gemm(a, b, out_fp32=True)
    assert relerr(out, acc) < 1e-5
    acc32 = torch.randn(M, N, device="cuda")
    out = lib.gemm(a, b, out=acc32.clone(), residual=acc32, out_fp32=True)
    assert relerr(out, acc + acc32) < 1e-5


@pytest.mark.parametrize("M,N,K", [(2050, 4096, 1024), (300, 896, 4096)])
def test_gemm_aux_epilogues(lib, M, N, K):
    """Training epilogues: fc1 forward also stores the GELU input; fc2 dgrad multiplies by gelu'(pre)."""
    a, b, bias = rnd(M, K, seed=1, scale=0.5), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3)
    pre_ref = a.float() @ b.float().t() + bias.float()
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = lib.gemm(a, b, bias=bias, act=lib.ACT_GELU, aux=pre, aux_mode=1)
    assert relerr(pre, pre_ref) < 1e-2 and relerr(out, F.gelu(pre_ref)) < 1e-2
    dy, w2 = rnd(M, 256, seed=4), rnd(256, N, seed=5, scale=0.05)     # d act = dy @ W2  (W2 [out=256, in=N], b_t form)
    got = lib.gemm(dy, w2, b_t=True, aux=pre, aux_mode=2)
    pr = pre.float().requires_grad_()
    F.gelu(pr).backward(dy.float() @ w2.float())
    assert relerr(got, pr.grad) < 2e-2


@pytest.mark.parametrize("M,N,K,K2,kw", [
    (36800 // 8, 1152, 896, 96, dict(bias=True)),                 # 1-CTA tile, 2 tail k-blocks (second one half out of bounds)
    (4728, 896, 896, 32, dict(residual=True, out_fp32=True)),     # 2-CTA <224>, fp32 residual stream
    (4728, 9728, 896, 64, dict(swiglu=True)),                     # 2-CTA <256> SwiGLU
    (4600, 896, 4864, 32, dict(residual=True, out_fp32=True)),
    (31, 1152, 896, 96, dict(bias=True)),                         # mma.sync weight streaming (query pass)
    (20, 9728, 896, 64, dict(swiglu=True)),
    (3, 896, 4864, 32, dict(residual=True, out_fp32=True)),       # GEMV
    (1, 1152, 896, 96, dict(bias=True)),
])
def test_gemm_second_a_source(lib, M, N, K, K2, kw):
    """y = [x | t] [W | B]^T: the un-merged LoRA branch rides in the base GEMM's k loop (slb_gemm_args.A2) - every kernel family"""
    x, t, w = rnd(M, K, seed=1), rnd(M, K2, seed=2), rnd(N, K + K2, scale=0.05, seed=3)
    ref = torch.cat([x, t], 1).float() @ w.float().t()
    args = {}
    if kw.get("bias"):
        args["bias"] = rnd(N, seed=4)
        ref = ref + args["bias"].float()
    if kw.get("swiglu"):
        g = ref.view(M, N // 256, 2, 128)
        ref = (F.silu(g[:, :, 0]) * g[:, :, 1]).reshape(M, N // 2)
        args["swiglu"] = True
    if kw.get("residual"):
        res = torch.randn(M, N, device="cuda")
        ref = ref + res
        out = res.clone()
        got = lib.gemm(x, w, out=out, residual=out, out_fp32=True, a2=t)
    else:
        got = lib.gemm(x, w, a2=t, **args)
    assert relerr(got, ref) < (1e-2 if kw.get("swiglu") else 5e-3)


@pytest.mark.parametrize("M,N,K,bn", [(4200, 1024, 1024, 0), (8200, 896, 4864, 0), (4100, 1024, 512, 2192), (4097, 256, 128, 0)])
def test_gemm_tma_residual_stream_epilogue(lib, M, N, K, bn):
    """2-CTA kernel, fp32 residual stream moved by TMA (32 x 32 boxes per epilogue warp): bias * layer-scale + residual in place,
    out-of-place, alpha, ragged last row tile (rows beyond M are zero-filled on load and clipped on store)"""
    a, b = rnd(M, K, seed=1, scale=0.5), rnd(N, K, seed=2, scale=0.05)
    bias, ls = rnd(N, seed=3), rnd(N, seed=4, scale=0.2)
    acc = a.float() @ b.float().t()
    x = torch.randn(M + 3, N, device="cuda")           # guard rows behind the tensor must stay untouched
    guard = x[M:].clone()
    ref = x[:M] + ls.float() * (acc + bias.float())
    lib.gemm(a, b, out=x[:M], bias=bias, scale_n=ls, residual=x[:M], out_fp32=True, block_n=bn)
    assert relerr(x[:M], ref) < 2e-3 and torch.equal(x[M:], guard)
    res = torch.randn(M, N, device="cuda")
    out = lib.gemm(a, b, residual=res, out_fp32=True, alpha=0.5, block_n=bn)
    assert relerr(out, res + 0.5 * acc) < 2e-3
    out = lib.gemm(a, b, bias=bias, residual=res, out_fp32=True, block_n=bn)
    assert relerr(out, res + acc + bias.float()) < 2e-3


@pytest.mark.parametrize("tiles,fp32", [(2, True), (3, False)])
def test_patch_embed_implicit_gemm(lib, tiles, fp32):
    """implicit-GEMM patch embedding (operand gathered from the pixels inside the kernel, class token + position embedding in the
    epilogue) against conv2d in fp32 and against the explicit im2col + GEMM + assemble path"""
    px = rnd(tiles, 3, 448, 448, seed=1)
    wconv = rnd(1024, 3, 14, 14, seed=2, scale=0.05)
    bias, cls, pos = rnd(1024, seed=3), rnd(1024, seed=4), rnd(1025, 1024, seed=5)
    wp = torch.zeros(1024, 640, device="cuda", dtype=torch.bfloat16)
    wp[:, :588] = wconv.reshape(1024, 588)
    got = lib.patch_embed(px, wp, bias, cls, pos, fp32=fp32)
    conv = F.conv2d(px.float(), wconv.float(), bias.float(), stride=14)                   # [T, 1024, 32, 32]
    ref = torch.cat([cls.float().expand(tiles, 1, 1024), conv.flatten(2).transpose(1, 2)], 1) + pos.float()
    assert got.shape == (tiles * 1025, 1024) and got.dtype == (torch.float32 if fp32 else torch.bfloat16)
    assert relerr(got.view(tiles, 1025, 1024), ref) < (1e-3 if fp32 else 1e-2)
    old = lib.vit_assemble(lib.gemm(lib.im2col_patch(px, 640), wp, bias=bias), cls, pos, tiles, fp32=fp32)
    assert relerr(got, old) < 1e-2


@pytest.mark.parametrize("M,N,K", [(4728, 992, 1152), (16400, 1024, 4096), (4100, 304, 136)])
def test_gemm_cluster_kernel_mn_major_dgrad(lib, M, N, K):
    """dX = dY W (b_t: W is [K, N] in memory) on the cta_group::2 kernel"""
    dy, w = rnd(M, K, seed=1, scale=0.5), rnd(K, N + 8, seed=2, scale=0.05)[:, :N]
    ref = dy.float() @ w.float()
    assert relerr(lib.gemm(dy, w, b_t=True, block_n=2256), ref) < 5e-3
    assert relerr(lib.gemm(dy, w, b_t=True), ref) < 5e-3            # whatever the tile selection picks


@pytest.mark.parametrize("M,N,K", [(3072, 1024, 16400), (2048, 4096, 4728), (2104, 640, 1000)])
def test_gemm_cluster_kernel_mn_major_wgrad(lib, M, N, K):
    """dW = dY^T X (a_t + b_t: both operands [K, *] in memory), accumulating into the output"""
    dy, x = rnd(K, M, seed=1, scale=0.5), rnd(K, N, seed=2, scale=0.5)
    ref = dy.float().t() @ x.float()
    assert relerr(lib.gemm(dy, x, a_t=True, b_t=True, block_n=2256), ref) < 5e-3
    g = rnd(M, N, seed=3)
    got = lib.gemm(dy, x, out=g.clone(), a_t=True, b_t=True, residual=g, alpha=0.5)
    assert relerr(got, g.float() + 0.5 * ref) < 5e-3


def test_gemm_strided_views(lib):
    """A as a column slice of a wider buffer (q part of the fused qkv), out as a slice."""
    M, K, N = 545, 896, 896
    big = rnd(M, 1152, seed=7)
    a = big[:, :K]
    b = rnd(N, K, seed=8, scale=0.05)
    outbig = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    lib.gemm(a, b, out=outbig[:, N:])
    assert relerr(outbig[:, N:], a.float() @ b.float().t()) < 1e-2
    assert outbig[:, :N].abs().max().item() == 0


def test_gemm_swiglu(lib):
    M, K, I = 545, 896, 4864
    a = rnd(M, K, seed=1)
    wg, wu = rnd(I, K, seed=2, scale=0.05), rnd(I, K, seed=3, scale=0.05)
    # interleave in groups of 128 rows: [g0..127, u0..127, g128..255, ...]
    w = torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], dim=1).reshape(2 * I, K).contiguous()
    out = lib.gemm(a, w, swiglu=True)
    ref = F.silu(a.float() @ wg.float().t()) * (a.float() @ wu.float().t())
    assert out.shape == (M, I)
    assert relerr(out, ref) < 1e-2


@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 16, 17, 31, 32])
def test_gemv_small_m(lib, M):
    """M <= 4 takes the weight-streaming GEMV, 4 < M <= 32 the mma.sync skinny GEMM (decode paths); same epilogue
    contract as the tcgen05 GEMM."""
    K, N, I = 896, 1152, 4864
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05)
    bias, ls, res = rnd(N, seed=3), rnd(N, seed=4, scale=0.2), rnd(M, N, seed=5)
    acc = a.float() @ b.float().t()
    assert relerr(lib.gemm(a, b), acc) < 1e-2
    assert relerr(lib.gemm(a, b, bias=bias, act=lib.ACT_GELU), F.gelu(acc + bias.float())) < 1e-2
    assert relerr(lib.gemm(a, b, bias=bias, scale_n=ls, residual=res, alpha=0.5), res.float() + ls.float() * (0.5 * acc + bias.float())) < 1e-2
    assert relerr(lib.gemm(a, b, out_fp32=True), acc) < 1e-5
    # forced tensor-core path agrees
    assert relerr(lib.gemm(a, b, block_n=128), acc) < 1e-2
    # long K (down_proj), in-place residual, strided A (a slice of a wider buffer)
    wide = rnd(M, I + 64, seed=6)
    a2, w2 = wide[:, :I], rnd(K, I, seed=7, scale=0.05)
    x = rnd(M, K, seed=8)
    ref = x.float() + a2.float() @ w2.float().t()
    assert relerr(lib.gemm(a2, w2, out=x, residual=x), ref) < 1e-2
    # fused SwiGLU over the interleaved gate|up layout
    wg, wu = rnd(I, K, seed=9, scale=0.05), rnd(I, K, seed=10, scale=0.05)
    w = torch.stack([wg.view(I // 128, 128, K), wu.view(I // 128, 128, K)], dim=1).reshape(2 * I, K).contiguous()
    out = lib.gemm(a, w, swiglu=True)
    assert out.shape == (M, I)
    assert relerr(out, F.silu(a.float() @ wg.float().t()) * (a.float() @ wu.float().t())) < 1e-2
    # odd N (LM head), fp32 logits
    V = 151655 if M == 1 else 4099
    wv = rnd(V, K, seed=11, scale=0.05)
    assert relerr(lib.gemm(a, wv, out_fp32=True), a.float() @ wv.float().t()) < 1e-5


def test_gemm_lm_head_odd_vocab(lib):
    """N = 151655 (odd): fp32 logits with an unaligned row stride."""
    M, K, N = 3, 896, 151655
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05)
    out = lib.gemm(a, b, out_fp32=True)
    ref = a.float() @ b.float().t()
    assert relerr(out, ref) < 1e-5
    assert (out.argmax(-1) == ref.argmax(-1)).all()


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (2050, 1024, 4096), (545, 896, 1152), (591, 4864, 896)])
def test_gemm_dgrad_form(lib, M, N, K):
    """b_t: B given as [K, N] row-major (dX = dY @ W with W stored [out,in])."""
    a, bt = rnd(M, K, seed=1), rnd(K, N, seed=2, scale=0.05)
    out = lib.gemm(a, bt, b_t=True)
    assert relerr(out, a.float() @ bt.float()) < 1e-2


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1024, 4096, 2050), (896, 32, 545), (32, 896, 4728)])
def test_gemm_wgrad_form(lib, M, N, K):
    """a_t + b_t: dW[M=out, N=in] = dY^T X with dY stored [K, M], X stored [K, N] (K = tokens)."""
    at, bt = rnd(K, M, seed=1), rnd(K, N, seed=2)
    out = lib.gemm(at, bt, a_t=True, b_t=True, out_fp32=True)
    assert relerr(out, at.float().t() @ bt.float()) < 1e-4


def test_layernorm_rmsnorm(lib):
    for rows, cols in [(2050, 1024), (512, 4096), (37, 1024)]:
        x, w, b = rnd(rows, cols, seed=1), rnd(cols, seed=2), rnd(cols, seed=3)
        y = lib.layernorm(x, w, b, 1e-6)
        ref = F.layer_norm(x.float(), (cols,), w.float(), b.float(), 1e-6)
        assert relerr(y, ref) < 1e-2
    x, w = rnd(575, 896, seed=1), rnd(896, seed=2)
    y = lib.rmsnorm(x, w, 1e-6)
    ref = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
    assert relerr(y, ref) < 1e-2


def test_patch_embed_glue(lib):
    T = 2
    px = rnd(T, 3, 448, 448, seed=1)
    w = rnd(1024, 3, 14, 14, seed=2, scale=0.02)
    bias, cls, pos = rnd(1024, seed=3, scale=0.02), rnd(1, 1, 1024, seed=4), rnd(1, 1025, 1024, seed=5)
    cols = lib.im2col_patch(px, 640)
    ref_cols = F.unfold(px.float(), kernel_size=14, stride=14).transpose(1, 2).reshape(T * 1024, 588)
    assert torch.equal(cols[:, :588].float(), ref_cols)
    assert cols[:, 588:].abs().max().item() == 0
    wpad = torch.zeros(1024, 640, device="cuda", dtype=torch.bfloat16)
    wpad[:, :588] = w.reshape(1024, 588)
    po = lib.gemm(cols, wpad, bias=bias)
    x = lib.vit_assemble(po, cls, pos, T)
    conv = F.conv2d(px.float(), w.float(), bias.float(), stride=14).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.float().expand(T, 1, -1), conv], 1) + pos.float()
    assert relerr(x.view(T, 1025, 1024), ref) < 1e-2


def test_pixel_shuffle_ln(lib):
    from oracle.model import pixel_shuffle_closed_form
    T = 3
    x = rnd(T * 1025, 1024, seed=1)
    w, b = rnd(4096, seed=2), rnd(4096, seed=3)
    y = lib.pixel_shuffle_ln(x, w, b, T, 1e-5)
    xs = pixel_shuffle_closed_form(x.float().view(T, 1025, 1024)[:, 1:], 32)
    ref = F.layer_norm(xs, (4096,), w.float(), b.float(), 1e-5).reshape(T * 256, 4096)
    assert relerr(y, ref) < 1e-2


@pytest.mark.parametrize("tiles,n", [(1, 128), (2, 1025), (3, 300)])
def test_attn_vit(lib, tiles, n):
    H = 16
    qkv = rnd(tiles * n, 3 * H * 64, seed=1)
    out = lib.attn_vit(qkv, tiles, n, H)
    q, k, v = qkv.float().view(tiles, n, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(tiles * n, H * 64)
    assert relerr(out, ref) < 2e-2


def _gqa_ref(q, k, v, past, valid):
    # q [B,Hq,Lq,64], k/v [B,Hkv,Lk,64]
    B, Hq, Lq, _ = q.shape
    Lk = k.shape[2]
    k = k.repeat_interleave(Hq // k.shape[1], 1)
    v = v.repeat_interleave(Hq // v.shape[1], 1)
    s = (q @ k.transpose(-1, -2)) * 0.125
    i = torch.arange(Lq, device=q.device)[:, None] + past
    j = torch.arange(Lk, device=q.device)[None]
    m = j <= i
    if valid is not None:
        m = m[None, None] & valid[:, None, None, :Lk].bool()
    s = s.masked_fill(~m, float("-inf"))
    return torch.softmax(s, -1).nan_to_num(0.0) @ v


@pytest.mark.parametrize("B,lq,past,use_valid", [(1, 545, 0, False), (2, 200, 0, True), (2, 30, 549, False),
                                                 (3, 1, 577, True), (1, 130, 64, False),
                                                 (6, 1, 600, True), (5, 31, 560, False)])  # batch >= 4: GQA-grouped decode kernel
def test_attn_gqa_and_rope(lib, B, lq, past, use_valid):
    Hq, Hkv, lmax = 14, 2, 704
    qkv = rnd(B * lq, (Hq + 2 * Hkv) * 64, seed=1)
    kc = torch.zeros(B, Hkv, lmax, 64, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    if past > 0:
        kc[:, :, :past] = rnd(B, Hkv, past, 64, seed=2)
        vc[:, :, :past] = rnd(B, Hkv, past, 64, seed=3)
    valid = None
    if use_valid:
        valid = torch.ones(B, lmax, device="cuda", dtype=torch.uint8)
        valid[1, :5] = 0
    ref_qkv = qkv.float().view(B, lq, Hq + 2 * Hkv, 64)
    pos = torch.arange(past, past + lq, device="cuda").float()
    inv = 1.0 / (1.0e6 ** (torch.arange(0, 64, 2, device="cuda").float() / 64))
    fr = pos[:, None] * inv
    cos, sin = torch.cat([fr, fr], -1).cos()[None, :, None], torch.cat([fr, fr], -1).sin()[None, :, None]

    def rot(x):
        return torch.cat([-x[..., 32:], x[..., :32]], -1)

    qr = ref_qkv[:, :, :Hq] * cos + rot(ref_qkv[:, :, :Hq]) * sin
    kr = ref_qkv[:, :, Hq:Hq + Hkv] * cos + rot(ref_qkv[:, :, Hq:Hq + Hkv]) * sin
    vr = ref_qkv[:, :, Hq + Hkv:]
    kfull = kc.float().clone()
    vfull = vc.float().clone()
    kfull[:, :, past:past + lq] = kr.transpose(1, 2)
    vfull[:, :, past:past + lq] = vr.transpose(1, 2)

    lib.rope_kv_write(qkv, kc, vc, B, lq, past)
    assert relerr(qkv.view(B, lq, Hq + 2 * Hkv, 64)[:, :, :Hq], qr) < 1e-2
    assert relerr(kc[:, :, :past + lq], kfull[:, :, :past + lq]) < 1e-2
    assert torch.equal(vc[:, :, past:past + lq].float(), vr.transpose(1, 2))

    out = lib.attn_gqa(qkv, qkv.stride(0), kc, vc, B, lq, past, key_valid=valid)
    qq = qkv.float().view(B, lq, Hq + 2 * Hkv, 64)[:, :, :Hq].transpose(1, 2)
    ref = _gqa_ref(qq, kc.float()[:, :, :past + lq], vc.float()[:, :, :past + lq], past, valid)
    ref = ref.transpose(1, 2).reshape(B * lq, Hq * 64)
    assert relerr(out, ref) < 2e-2


@pytest.mark.parametrize("B,L,lmax,pad", [(1, 128, 128, None), (3, 575, 640, None), (5, 591, 640, "right"), (4, 545, 768, "left"), (40, 575, 640, None)])
def test_attn_gqa_prefill_kernel(lib, B, L, lmax, pad):
    """The persistent head-pair prefill kernel (attention_gqa.cu; Lq >= 128, past = 0): causal block skipping + diagonal mask,
    7 heads per kv head in pairs (one single-head item per group), left / right key padding through validity words, a cache
    longer than the sequence, enough items for several per CTA (B = 40: 1600 items), output and log-sum-exp."""
    Hq, Hkv = 14, 2
    q = rnd(B * L, (Hq + 2 * Hkv) * 64, seed=11)
    kc = torch.zeros(B, Hkv, lmax, 64, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    kc[:, :, :L] = rnd(B, Hkv, L, 64, seed=12)
    vc[:, :, :L] = rnd(B, Hkv, L, 64, seed=13)
    if lmax > L:   # stale rows beyond the sequence (a reused cache): must never be attended
        kc[:, :, L:] = 7.0
        vc[:, :, L:] = -9.0
    valid = None
    if pad is not None:
        valid = torch.ones(B, lmax, device="cuda", dtype=torch.uint8)
        if pad == "left":
            valid[1, :9] = 0
            valid[2, :200] = 0      # more than one whole key block masked
        else:
            valid[1, L - 40:] = 0
            valid[3, L - 300:] = 0
    lse = torch.empty(B, Hq, L, device="cuda")
    out = lib.attn_gqa(q, q.stride(0), kc, vc, B, L, 0, key_valid=valid, lse=lse)
    qq = q.float().view(B, L, Hq + 2 * Hkv, 64)[:, :, :Hq].transpose(1, 2)
    kk, vv = kc.float()[:, :, :L].repeat_interleave(Hq // Hkv, 1), vc.float()[:, :, :L].repeat_interleave(Hq // Hkv, 1)
    s = (qq @ kk.transpose(-1, -2)) * 0.125
    m = torch.ones(L, L, device="cuda", dtype=torch.bool).tril()[None, None]
    if valid is not None:
        m = m & valid[:, None, None, :L].bool()
    s = s.masked_fill(~m, float("-inf"))
    ref = (torch.softmax(s, -1).nan_to_num(0.0) @ vv).transpose(1, 2).reshape(B * L, Hq * 64)
    lse_ref = torch.logsumexp(s, -1)
    assert torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 2e-2
    live = torch.isfinite(lse_ref)
    assert torch.equal(torch.isfinite(lse), live)                     # fully masked (padding) rows: out = 0, lse = -inf
    assert (lse[live] - lse_ref[live]).abs().max().item() < 2e-2
    dead_rows = (~live).transpose(1, 2).reshape(B * L, Hq)
    if dead_rows.any():
        assert out.view(B * L, Hq, 64)[dead_rows].abs().max().item() == 0.0


def test_embed_assemble_gather_argmax(lib):
    B, L, H, V, n_img = 2, 545, 896, 4096, 512
    img_id = V - 7
    table = rnd(V, H, seed=1)
    ids = torch.randint(0, V - 20, (B, L), device="cuda")
    ids[:, 4:4 + n_img] = img_id
    ids[:, 540:542] = V + 7  # out-of-table id -> clamp
    vit = rnd(B * n_img, H, seed=2)
    wp = rnd(B, 2, H, seed=3)
    wp_start = torch.tensor([540, -1], device="cuda", dtype=torch.int32)
    out = lib.embed_assemble(ids, table, vit, wp, wp_start, 2, img_id, n_img)
    ref = table[ids.clamp(0, V - 1)].clone()
    ref[:, 4:4 + n_img] = vit.view(B, n_img, H)
    ref[0, 540:542] = wp[0]
    assert torch.equal(out, ref)
    idx = torch.tensor([5, 0, V + 3, 17], device="cuda")
    assert torch.equal(lib.gather_rows(table, idx), table[idx.clamp(0, V - 1)])
    lg = torch.randn(5, 151655, device="cuda")
    marg = torch.empty(5, device="cuda")
    am = lib.argmax(lg, out_margin=marg)
    assert torch.equal(am, lg.argmax(-1))
    t2 = lg.topk(2, -1).values
    assert torch.allclose(marg, t2[:, 0] - t2[:, 1])


def test_heads_and_wp_encoder(lib):
    B = 3
    feats = rnd(B, 30, 896, seed=1)
    W = {k: rnd(*s, seed=i + 10, scale=0.05) for i, (k, s) in enumerate({
        "r0w": (512, 896), "r0b": (512,), "r2w": (256, 512), "r2b": (256,), "r4w": (2, 256),
        "s0w": (256, 896), "s0b": (256,), "s2w": (2, 256)}.items())}
    hw = lib.HeadsWeights(*[W[k].data_ptr() for k in ("r0w", "r0b", "r2w", "r2b", "r4w", "s0w", "s0b", "s2w")])
    route, speed = lib.driving_heads(feats, 30 * 896, hw, B)
    f = feats.float()
    r = F.silu(F.linear(f[:, :20], W["r0w"].float(), W["r0b"].float()))
    r = F.silu(F.linear(r, W["r2w"].float(), W["r2b"].float()))
    r = F.linear(r, W["r4w"].float()).cumsum(1)
    s = F.linear(F.silu(F.linear(f[:, 20:], W["s0w"].float(), W["s0b"].float())), W["s2w"].float()).cumsum(1)
    assert relerr(route, r) < 1e-4 and relerr(speed, s) < 1e-4
    P = {k: rnd(*s_, seed=i + 30, scale=0.05) for i, (k, s_) in enumerate({
        "w0": (256, 2), "b0": (256,), "w2": (512, 256), "b2": (512,), "w4": (896, 512), "b4": (896,)}.items())}
    ww = lib.WpWeights(*[P[k].data_ptr() for k in ("w0", "b0", "w2", "b2", "w4", "b4")])
    coords = torch.randn(4, 2, device="cuda") * 10
    out = lib.wp_encoder(coords, ww)
    h = F.relu(F.linear(coords, P["w0"].float(), P["b0"].float()))
    h = F.relu(F.linear(h, P["w2"].float(), P["b2"].float()))
    ref = F.linear(h, P["w4"].float(), P["b4"].float())
    assert relerr(out, ref) < 1e-2


def test_silu_mul_add(lib):
    g, u = rnd(545, 4864, seed=1), rnd(545, 4864, seed=2)
    assert relerr(lib.silu_mul(g, u), F.silu(g.float()) * u.float()) < 1e-2
    assert relerr(lib.add(g, u), g.float() + u.float()) < 1e-2
