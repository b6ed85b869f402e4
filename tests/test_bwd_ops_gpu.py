"""Backward / optimizer kernels on a B200 against torch autograd (fp32, computed on the GPU from the same bf16
inputs).  Tolerance: max-abs error relative to the reference's max magnitude < 2e-2 (bf16 outputs), 1e-3 (fp32)."""
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


def test_layernorm_bwd(lib):
    for rows, cols in [(2050, 1024), (300, 4096)]:
        x, w, b, dy = rnd(rows, cols, seed=1), rnd(cols, seed=2), rnd(cols, seed=3), rnd(rows, cols, seed=4)
        mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
        lib.layernorm(x, w, b, 1e-6, stats=(mean, rstd))
        dw, db = torch.zeros(cols, device="cuda"), torch.zeros(cols, device="cuda")
        dx = lib.layernorm_bwd(dy, x, w, mean, rstd, dw, db)
        xr, wr, br = x.float().requires_grad_(), w.float().requires_grad_(), b.float().requires_grad_()
        F.layer_norm(xr, (cols,), wr, br, 1e-6).backward(dy.float())
        assert relerr(dx, xr.grad) < 2e-2 and relerr(dw, wr.grad) < 1e-3 and relerr(db, br.grad) < 1e-3
        res = rnd(rows, cols, seed=9)   # gradient of the residual path, summed into dx before the single rounding
        dx2 = lib.layernorm_bwd(dy, x, w, mean, rstd, torch.zeros_like(dw), torch.zeros_like(db), add=res)
        assert relerr(dx2, xr.grad + res.float()) < 2e-2


def test_rmsnorm_bwd(lib):
    rows, cols = 621, 896
    x, w, dy = rnd(rows, cols, seed=1), rnd(cols, seed=2), rnd(rows, cols, seed=4)
    rstd = torch.empty(rows, device="cuda")
    lib.rmsnorm(x, w, 1e-6, rstd=rstd)
    dx = lib.rmsnorm_bwd(dy, x, w, rstd)
    xr = x.float().requires_grad_()
    (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()).backward(dy.float())
    assert relerr(dx, xr.grad) < 2e-2
    res = rnd(rows, cols, seed=9)
    assert relerr(lib.rmsnorm_bwd(dy, x, w, rstd, add=res), xr.grad + res.float()) < 2e-2


def test_pixel_shuffle_ln_bwd(lib):
    from oracle.model import pixel_shuffle_closed_form
    T = 2
    x, w, b, dy = rnd(T * 1025, 1024, seed=1), rnd(4096, seed=2), rnd(4096, seed=3), rnd(T * 256, 4096, seed=4)
    mean = torch.empty(T * 256, device="cuda"); rstd = torch.empty(T * 256, device="cuda")
    lib.pixel_shuffle_ln(x, w, b, T, 1e-5, stats=(mean, rstd))
    dw, db = torch.zeros(4096, device="cuda"), torch.zeros(4096, device="cuda")
    dx = lib.pixel_shuffle_ln_bwd(dy, x, w, mean, rstd, dw, db, T)
    xr, wr, br = x.float().requires_grad_(), w.float().requires_grad_(), b.float().requires_grad_()
    xs = pixel_shuffle_closed_form(xr.view(T, 1025, 1024)[:, 1:], 32)
    F.layer_norm(xs, (4096,), wr, br, 1e-5).reshape(T * 256, 4096).backward(dy.float())
    assert relerr(dx, xr.grad) < 2e-2 and relerr(dw, wr.grad) < 1e-3 and relerr(db, br.grad) < 1e-3


def test_activation_and_dropout(lib):
    pre, dout = rnd(300, 4096, seed=1), rnd(300, 4096, seed=2)
    assert relerr(lib.gelu_fwd(pre), F.gelu(pre.float())) < 1e-2
    pr = pre.float().requires_grad_()
    F.gelu(pr).backward(dout.float())
    assert relerr(lib.gelu_bwd(pre, dout), pr.grad) < 2e-2
    g, u = rnd(300, 4864, seed=3), rnd(300, 4864, seed=4)
    d = rnd(300, 4864, seed=5)
    gr, ur = g.float().requires_grad_(), u.float().requires_grad_()
    (F.silu(gr) * ur).backward(d.float())
    dg, du = lib.silu_mul_bwd(g, u, d)
    assert relerr(dg, gr.grad) < 2e-2 and relerr(du, ur.grad) < 2e-2
    x = rnd(1000, 896, seed=6)
    y = lib.dropout(x, 0.1, 1234)
    keep = y.float() != 0
    assert abs(keep.float().mean().item() - 0.9) < 0.01
    assert relerr(y[keep], x.float()[keep] / 0.9) < 1e-2
    assert torch.equal(lib.dropout(x, 0.1, 1234), y) and not torch.equal(lib.dropout(x, 0.1, 99), y)
    assert torch.equal(lib.dropout(x, 0.0, 5), x)


def test_col_reduce_and_scale(lib):
    a, b = rnd(2050, 1024, seed=1), rnd(2050, 1024, seed=2)
    acc = torch.zeros(1024, device="cuda")
    lib.col_reduce(a, acc)
    assert relerr(acc, a.float().sum(0)) < 1e-3
    acc.zero_()
    lib.col_reduce(a, acc, b, alpha=2.0)
    assert relerr(acc, 2.0 * (a.float() * b.float()).sum(0)) < 1e-3
    s = rnd(1024, seed=3)
    assert relerr(lib.scale_cols(a, s), a.float() * s.float()) < 1e-2
    c = a.clone()
    lib.add_inplace(c, b)
    assert relerr(c, a.float() + b.float()) < 1e-2


def test_layerscale_bwd(lib):
    dx, br, ls = rnd(2050, 1024, seed=1), rnd(2050, 1024, seed=2), rnd(1024, seed=3, scale=0.2)
    dls, db = torch.zeros(1024, device="cuda"), torch.zeros(1024, device="cuda")
    out = lib.layerscale_bwd(dx, br, ls, dls, db)
    assert relerr(out, dx.float() * ls.float()) < 1e-2
    assert relerr(dls, (dx.float() * br.float()).sum(0)) < 1e-3
    assert relerr(db, out.float().sum(0)) < 1e-3


def test_vit_assemble_bwd(lib):
    T = 3
    dx = rnd(T * 1025, 1024, seed=1)
    dcls, dpos = torch.zeros(1024, device="cuda"), torch.zeros(1025, 1024, device="cuda")
    dpatch = lib.vit_assemble_bwd(dx, dcls, dpos, T)
    d3 = dx.float().view(T, 1025, 1024)
    assert torch.equal(dpatch.view(T, 1024, 1024), dx.view(T, 1025, 1024)[:, 1:])
    assert relerr(dcls, d3[:, 0].sum(0)) < 1e-4 and relerr(dpos, d3.sum(0)) < 1e-4


@pytest.mark.parametrize("tiles,n", [(1, 128), (2, 1025), (2, 300)])
def test_attn_vit_bwd(lib, tiles, n):
    H = 16
    qkv, dout = rnd(tiles * n, 3 * H * 64, seed=1), rnd(tiles * n, H * 64, seed=2)
    lse = torch.empty(tiles, H, n, device="cuda")
    out = lib.attn_vit(qkv, tiles, n, H, lse=lse)
    delta = lib.attn_delta(out, dout, tiles, n, H)
    dqkv = lib.attn_vit_bwd(qkv, dout, lse, delta, tiles, n, H)
    ref_in = qkv.float().requires_grad_()
    q, k, v = ref_in.view(tiles, n, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(tiles * n, H * 64)
    ref.backward(dout.float())
    assert relerr(out, ref) < 2e-2
    assert relerr(dqkv, ref_in.grad) < 3e-2


@pytest.mark.parametrize("B,L,use_valid", [(2, 200, False), (2, 621, True)])
def test_attn_gqa_bwd(lib, B, L, use_valid):
    Hq, Hkv, lmax = 14, 2, 640
    qkv, dout = rnd(B * L, (Hq + 2 * Hkv) * 64, seed=1), rnd(B * L, Hq * 64, seed=2)
    kc = torch.zeros(B, Hkv, lmax, 64, device="cuda", dtype=torch.bfloat16); vc = torch.zeros_like(kc)
    valid = None
    if use_valid:
        valid = torch.ones(B, lmax, device="cuda", dtype=torch.uint8)
        valid[1, L - 40:] = 0  # right padding as in the permuted training stream
    x = qkv.float().requires_grad_()
    x4 = x.view(B, L, Hq + 2 * Hkv, 64)
    pos = torch.arange(L, device="cuda").float()
    inv = 1.0 / (1.0e6 ** (torch.arange(0, 64, 2, device="cuda").float() / 64))
    fr = pos[:, None] * inv
    cos, sin = torch.cat([fr, fr], -1).cos()[None, :, None], torch.cat([fr, fr], -1).sin()[None, :, None]
    rot = lambda t: torch.cat([-t[..., 32:], t[..., :32]], -1)
    qr = (x4[:, :, :Hq] * cos + rot(x4[:, :, :Hq]) * sin).transpose(1, 2)
    kr = (x4[:, :, Hq:Hq + Hkv] * cos + rot(x4[:, :, Hq:Hq + Hkv]) * sin).transpose(1, 2)
    vr = x4[:, :, Hq + Hkv:].transpose(1, 2)
    s = (qr @ kr.repeat_interleave(Hq // Hkv, 1).transpose(-1, -2)) * 0.125
    m = torch.ones(L, L, device="cuda", dtype=torch.bool).tril()[None, None]
    if valid is not None:
        m = m & valid[:, None, None, :L].bool()
    s = s.masked_fill(~m, float("-inf"))
    ref = (torch.softmax(s, -1) @ vr.repeat_interleave(Hq // Hkv, 1)).transpose(1, 2).reshape(B * L, Hq * 64)
    rowmask = torch.ones(B * L, 1, device="cuda")
    if valid is not None:
        rowmask = valid[:, :L].reshape(B * L, 1).float()   # padded query rows carry no gradient in the real loss
    ref.backward(dout.float() * rowmask)

    work = qkv.clone()
    lib.rope_kv_write(work, kc, vc, B, L, 0)
    lse = torch.empty(B, Hq, L, device="cuda")
    out = lib.attn_gqa(work, work.stride(0), kc, vc, B, L, 0, key_valid=valid, lse=lse)
    d_in = (dout.float() * rowmask).to(torch.bfloat16)
    delta = lib.attn_delta(out, d_in, B, L, Hq)
    dq, dk, dv = lib.attn_gqa_bwd(work, work.stride(0), kc, vc, d_in, lse, delta, B, L, key_valid=valid)
    dqkv = lib.rope_bwd(dq, dk, dv, B, L, Hq, Hkv, 1.0e6)
    assert relerr(dqkv, x.grad) < 3e-2


def test_ce_and_adamw(lib):
    from oracle.model import adamw_step
    R, V = 5, 151655
    lg = torch.randn(R, V, device="cuda") * 3
    labels = torch.tensor([5, 151654, -1, 77, 1000], device="cuda")
    loss, dl = lib.ce_fwd_bwd(lg, labels, grad_scale=0.25)
    lr_ = lg.clone().requires_grad_()
    ref = F.cross_entropy(lr_, labels.clamp_min(0), reduction="none") * (labels >= 0)
    (ref.sum() * 0.25).backward()
    assert relerr(loss, ref) < 1e-4
    assert relerr(dl[:, :V], lr_.grad) < 2e-2 and dl[:, V:].abs().max().item() == 0
    n = 4096 * 3
    p0 = torch.randn(n, device="cuda")
    g = (torch.randn(n, device="cuda") * 0.5).to(torch.bfloat16)
    master, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    sq = torch.zeros(1, device="cuda")
    q, qm, qv = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        sq.zero_()
        lib.grad_sqnorm(g, sq)
        assert relerr(sq, g.float().pow(2).sum()[None]) < 1e-4
        lib.adamw_fused(master, m, v, g, pb, 3e-3, 0.9, 0.999, 1e-8, 0.1, step, sqnorm=sq, max_norm=0.3)
        coef = min(1.0, 0.3 / (g.float().norm().item() + 1e-6))
        q, qm, qv = adamw_step(q, g.float() * coef, qm, qv, step, 3e-3)
        assert relerr(master, q) < 1e-5 and relerr(pb, q) < 1e-2


@pytest.mark.parametrize("B,L,H", [(1, 5, 3), (2, 1025, 16), (3, 7, 14)])
def test_attn_delta_rowwise_dot(lib, B, L, H):
    """delta[b, h, i] = <dO, O> per (row, head); totals that are not a multiple of the 4 units a warp covers"""
    g = torch.Generator(device="cuda").manual_seed(B * 100 + L)
    o = torch.randn(B * L, H * 64, device="cuda", generator=g).bfloat16()
    do = torch.randn(B * L, H * 64, device="cuda", generator=g).bfloat16()
    got = lib.attn_delta(o, do, B, L, H)
    want = (o.float() * do.float()).view(B, L, H, 64).sum(-1).permute(0, 2, 1)
    assert got.shape == (B, H, L)
    assert (got - want).abs().max().item() <= 1e-4 * max(1.0, want.abs().max().item())


# ---- un-merged LoRA training path: side kernels of the K-concatenated formulation (csrc/lora.cu) ----
def test_dropout_multi_matches_dropout(lib):
    x = rnd(621, 896, seed=1)
    sd = torch.tensor([5], device="cuda", dtype=torch.int64)
    seeds = [0x515100, 0x515101, 0x515102]
    outs = lib.dropout_multi(x, 0.1, seeds, seed_dev=sd)
    for s, o in zip(seeds, outs):
        assert torch.equal(o, lib.dropout(x, 0.1, s, seed_dev=sd))
    assert not torch.equal(outs[0], outs[1])   # independent masks per adapter
    keep = (outs[0] != 0).float().mean().item()
    assert abs(keep - 0.9) < 0.01, keep


def test_lora_pack(lib):
    r = 32
    srcs = [rnd(896, r, seed=1), rnd(128, r, seed=2), rnd(4864, r, seed=3)]
    K = 896
    wx = torch.zeros(896 + 128, K + 2 * r, device="cuda", dtype=torch.bfloat16)
    wd = torch.zeros(4864, 64 + r, device="cuda", dtype=torch.bfloat16)
    dsts = [wx[:, K:], wx[896:, K + r:], wd[:, 64:]]
    table = torch.tensor([(s.data_ptr(), d.data_ptr(), s.shape[0], d.stride(0)) for s, d in zip(srcs, dsts)], dtype=torch.int64, device="cuda")
    lib.lora_pack(table, 3, r, 2.0)
    assert torch.equal(wx[:896, K:K + r], (2.0 * srcs[0].float()).to(torch.bfloat16))
    assert torch.equal(wx[896:, K + r:], (2.0 * srcs[1].float()).to(torch.bfloat16))
    assert torch.equal(wd[:, 64:], (2.0 * srcs[2].float()).to(torch.bfloat16))
    assert float(wx[:, :K].abs().max()) == 0.0 and float(wx[:896, K + r:].abs().max()) == 0.0 and float(wd[:, :64].abs().max()) == 0.0


@pytest.mark.parametrize("rows", [4728, 64, 37, 1000])
def test_lora_wgrad_grouped(lib, rows):
    """one launch for the adapters' parameter gradients of a group: out_j (+)= alpha_j P_j^T Q_j with P / Q column slices of wider
    matrices (dy, t, [dx | dt]), a row count that is no multiple of the 64-row chunk, accumulation into an existing gradient"""
    r = 32
    dy = rnd(rows, 896 + 128 + 128, seed=1)            # q | k | v output gradients
    t = rnd(rows, 3 * r, seed=2, scale=0.3)
    cat = rnd(rows, 896 + 3 * r + 8, seed=3)[:, :896 + 3 * r]
    xs = [rnd(rows, 896, seed=4 + j) for j in range(3)]
    wide = rnd(rows, 4864, seed=9)                     # a 152-tile problem (down_proj lora_A / gate lora_B shapes)
    outs, probs, refs = [], [], []
    o = 0
    for j, w in enumerate((896, 128, 128)):
        P, Q = dy[:, o:o + w], t[:, r * j:r * (j + 1)]
        out = rnd(w, r, seed=20 + j)
        acc = j == 1
        refs.append(2.0 * P.float().t() @ Q.float() + (out.float() if acc else 0.0))
        probs.append((P, Q, out, 2.0, acc)); outs.append(out)
        P, Q = cat[:, 896 + r * j:896 + r * (j + 1)], xs[j]
        out = rnd(r, 896, seed=30 + j)
        refs.append(P.float().t() @ Q.float() + (out.float() if j == 2 else 0.0))
        probs.append((P, Q, out, 1.0, j == 2)); outs.append(out)
        o += w
    out = torch.empty(r, 4864, device="cuda", dtype=torch.bfloat16)
    refs.append(t[:, :r].float().t() @ wide.float())
    probs.append((t[:, :r], wide, out, 1.0, False)); outs.append(out)
    out = torch.empty(4864, r, device="cuda", dtype=torch.bfloat16)
    refs.append(0.5 * wide.float().t() @ t[:, r:2 * r].float())
    probs.append((wide, t[:, r:2 * r], out, 0.5, False)); outs.append(out)
    lib.lora_wgrad_grouped(probs, rows)
    torch.cuda.synchronize()
    for k, (got, ref) in enumerate(zip(outs, refs)):
        assert relerr(got, ref) < 1e-2, (k, relerr(got, ref))


@pytest.mark.parametrize("M,K,n,p", [(621, 896, 3, 0.1), (300, 4864, 1, 0.1), (130, 896, 2, 0.0), (64, 256, 4, 0.25)])
def test_lora_dx(lib, M, K, n, p):
    """out = base + sum_j mask_j o (dt_j A_j) / (1 - p): masks taken from slb_dropout with the same seeds"""
    r = 32
    cat = rnd(M, K + r * n + 8, seed=1)[:, :K + r * n]          # row stride != width
    A = [rnd(r, K, seed=10 + j, scale=0.2) for j in range(n)]
    seeds = [0xABC00 + j for j in range(n)]
    sd = torch.tensor([3], device="cuda", dtype=torch.int64)
    got = lib.lora_dx(cat, K, A, p=p, seeds=seeds if p > 0 else None, seed_dev=sd)
    ref = cat[:, :K].float()
    ones = torch.ones(M, K, device="cuda", dtype=torch.bfloat16)
    for j in range(n):
        m = lib.dropout(ones, p, seeds[j], seed_dev=sd).float() if p > 0 else ones.float()   # keep / (1 - p)
        ref = ref + m * (cat[:, K + r * j:K + r * (j + 1)].float() @ A[j].float())
    assert relerr(got, ref) < 1e-2


def test_silu_mul_cat(lib):
    M, I = 333, 4864
    gu, dout = rnd(M, 2 * I, seed=1), rnd(M, I, seed=2)
    g, u = gu[:, :I].float().requires_grad_(), gu[:, I:].float().requires_grad_()
    ref = F.silu(g) * u
    ref.backward(dout.float())
    assert relerr(lib.silu_mul_cat(gu), ref) < 1e-2
    dgu = lib.silu_mul_cat_bwd(gu, dout)
    assert relerr(dgu[:, :I], g.grad) < 1e-2 and relerr(dgu[:, I:], u.grad) < 1e-2
