"""Per-kernel parity on the GPU: every C-ABI entry point against a plain fp32 PyTorch restatement of the same op
(teacher-forced: identical inputs, so only accumulation order / operand rounding differ)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from deepsvg_b200 import ops
    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


# ------------------------------------------------------------------------------------------------ GEMMs
@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("M,N,K", [(512, 256, 256), (300, 768, 256), (496, 7, 256), (992, 2827, 256), (1024, 256, 512),
                                   (64, 256, 64), (16650, 2827, 256), (16650, 7, 256)])   # last two: wide / unaligned rows
def test_linear_matches_fp32(planes, M, N, K):
    ops = _ops()
    X, W, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
    xa, wa = ops.act_from_float(X, planes), ops.act_from_float(W, planes)
    out = torch.empty(M, N, device=DEV)
    ops.linear(xa, wa, M, N, K, bias=b, out_f32=out)
    ref = (xa.float().double() @ wa.float().double().t()).float() + b
    assert _rel(out, ref) < 2e-5
    if planes == 2:  # bf16x3 must track the un-rounded fp32 product
        full = (X.double() @ W.double().t()).float() + b
        assert _rel(out, full) < 3e-5


def test_linear_full_epilogue():
    ops = _ops()
    M, N, K, rpg = 620, 512, 256, 31
    X, W, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
    res, rv = _rand(M, N, seed=4), _rand(M // rpg, N, seed=5)
    msk = (_rand(M, N, seed=6) > 0).float()
    sc = torch.tensor([0.75], device=DEV)
    xa, wa, ma = ops.act_from_float(X, 1), ops.act_from_float(W, 1), ops.act_from_float(msk, 1)
    out = torch.empty(M, N, device=DEV)
    oa = ops.Act(M, N, 2, DEV)
    ops.linear(xa, wa, M, N, K, bias=b, scale_cols=100, scale=0.5, relu=True, rowvec=rv, rows_per_group=rpg, mask=ma,
               mask_scale=1.25, residual=res, out_f32=out, out_act=oa, acc_scale=sc)
    ref = (xa.float() @ wa.float().t()) * 0.75 + b
    ref[:, :100] *= 0.5
    ref = torch.relu(ref) + rv.repeat_interleave(rpg, 0)
    ref = ref * msk * 1.25 + res
    assert _rel(out, ref) < 2e-5
    assert _rel(oa.float(), ref) < 2e-4


@pytest.mark.parametrize("mode", ["dgrad", "qkv", "ffn1", "resid", "mask", "head_dgrad", "bias_f32", "res_f32"])
@pytest.mark.parametrize("M,N,planes", [(1000, 256, 1), (1000, 768, 1), (1000, 128, 1), (16650, 256, 1), (16650, 768, 1),
                                        (33000, 512, 1), (1000, 256, 2), (16650, 768, 2), (33000, 512, 2)])
def test_linear_lean_epilogues(mode, M, N, planes):
    """The compile-time specialised epilogues (one per GEMM role of the model) against the same fp32 restatement.
    M = 1000: the 128-wide, two-CTAs-per-SM kernels; M > 16384: the 256-wide persistent kernels (streamed bulk stores,
    16 epilogue warps for the fp32 / mask modes), with a ragged last row tile and several tiles per CTA.
    planes = 2: the parity-mode (bf16x3) kernels <128, 2, mode> with the same feature sets and hi + lo act outputs."""
    ops = _ops()
    K, rpg = 256, 25
    X, W = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5)
    b, res, rv = _rand(N, seed=3), _rand(M, N, seed=4), _rand(M // rpg, N, seed=5)
    msk = torch.relu(_rand(M, N, seed=6))
    xa, wa, ma = ops.act_from_float(X, planes), ops.act_from_float(W, planes), ops.act_from_float(msk, planes)
    acc = (xa.float().double() @ wa.float().double().t()).float()
    of, oa = torch.zeros(M, N, device=DEV), ops.Act(M, N, planes, DEV)
    sc = torch.tensor([0.37], device=DEV)
    if mode == "dgrad":
        ops.linear(xa, wa, M, N, K, out_act=oa)
        ref, got = acc, oa.float()
    elif mode == "qkv":
        ops.linear(xa, wa, M, N, K, bias=b, scale_cols=N // 4 * 2, scale=0.25, out_act=oa)
        ref = acc + b
        ref[:, :N // 4 * 2] *= 0.25
        got = oa.float()
    elif mode == "ffn1":
        ops.linear(xa, wa, M, N, K, bias=b, relu=True, drop=(0.2, 5, 77), out_act=oa)
        ka = ops.Act(M, N, 1, DEV)
        ops.cast_act(torch.ones(M, N, device=DEV), M, N, out=ka, drop=(0.2, 5, 77))
        ref, got = torch.relu(acc + b) * ka.float(), oa.float()
    elif mode == "resid":
        ops.linear(xa, wa, M, N, K, bias=b, drop=(0.2, 6, 77), rowvec=rv, rows_per_group=rpg, residual=res, out_f32=of)
        ka = ops.Act(M, N, 1, DEV)
        ops.cast_act(torch.ones(M, N, device=DEV), M, N, out=ka, drop=(0.2, 6, 77))
        got = of
        ref = (acc + b) * (ka.float() != 0) / (1 - 13107 / 65536.0) + rv.repeat_interleave(rpg, 0) + res
    elif mode == "mask":
        ops.linear(xa, wa, M, N, K, mask=ma, mask_scale=1.25, out_act=oa)
        ref, got = acc * (ma.float() != 0) * 1.25, oa.float()
    elif mode == "bias_f32":     # linear_global / VAE heads: bias only, fp32 out (lean mode 4 without residual)
        ops.linear(xa, wa, M, N, K, bias=b, out_f32=of)
        ref, got = acc + b, of
    elif mode == "res_f32":      # dgrad accumulated into an fp32 gradient (lean mode 4 without bias)
        of.copy_(res)
        ops.linear(xa, wa, M, N, K, residual=of, out_f32=of)
        ref, got = acc + res, of
    else:
        ops.linear(xa, wa, M, N, K, acc_scale=sc, residual=res, out_f32=of)
        ref, got = acc * 0.37 + res, of
    tol = 2e-5 if got is of else (6e-3 if planes == 1 else 4e-5)   # bf16 output rounding (one plane) / hi + lo planes
    if mode == "ffn1":
        ref = torch.relu(acc + b) * (ka.float() != 0) / (1 - 13107 / 65536.0)
    assert _rel(got, ref) < tol, mode


def test_linear_dropout_statistics_and_determinism():
    ops = _ops()
    M, N, K = 1024, 256, 64
    X = torch.ones(M, K, device=DEV)
    W = torch.ones(N, K, device=DEV) / K
    xa, wa = ops.act_from_float(X, 1), ops.act_from_float(W, 1)
    o1, o2, o3 = (torch.empty(M, N, device=DEV) for _ in range(3))
    ops.linear(xa, wa, M, N, K, drop=(0.1, 7, 1234), out_f32=o1)
    ops.linear(xa, wa, M, N, K, drop=(0.1, 7, 1234), out_f32=o2)
    ops.linear(xa, wa, M, N, K, drop=(0.1, 8, 1234), out_f32=o3)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)
    keep = (o1 != 0).float().mean().item()
    assert abs(keep - 0.9) < 0.005
    assert torch.allclose(o1[o1 != 0], torch.tensor(1 / 0.9, device=DEV), rtol=1e-5)
    # cast_act regenerates the same mask from (seed, site, row*N+col)
    ca = ops.Act(M, N, 1, DEV)
    ops.cast_act(torch.ones(M, N, device=DEV), M, N, out=ca, drop=(0.1, 7, 1234))
    assert torch.equal(ca.float() != 0, o1 != 0)


@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("M,P,Q", [(4096, 768, 256), (1000, 300, 200), (2048, 2827, 64), (992, 7, 256), (130, 256, 512),
                                   (640, 100, 30), (20000, 512, 256),    # Q % 4 != 0: scalar reductions; many row blocks
                                   # M >= 16384: the tall 256 x 256 tiles of the big path-level weight gradients (column sums
                                   # by the epilogue warps from shared memory)
                                   # >= 4 such tiles (the dispatch rule): 2827 x 256, 1536 x 512 (bias sums only from the
                                   # first Q tile), 512 x 512, ragged P with Q % 4 != 0 (scalar reductions), ragged P and Q
                                   (16500, 2827, 256), (20000, 1536, 512), (16500, 512, 512), (17000, 1000, 250),
                                   (33000, 600, 300),
                                   # fewer tiles at the same row counts: the 128 x 256 tiles
                                   (20000, 768, 256), (17000, 256, 512), (33000, 300, 200), (131072, 768, 256)])
def test_outer_matches_fp32(planes, M, P, Q):
    """planes = 1 runs the tall tiles where they apply; planes = 2 (parity mode) always the 128-row tiles."""
    ops = _ops()
    A, B = _rand(M, P, seed=1), _rand(M, Q, seed=2)
    lda, ldb = (P + 7) // 8 * 8, (Q + 7) // 8 * 8
    aa, ba = ops.act_from_float(A, planes, ld=lda), ops.act_from_float(B, planes, ld=ldb)
    Cout = torch.ones(P, Q, device=DEV)
    cs = torch.ones(P, device=DEV)
    sc = torch.tensor([2.0], device=DEV)
    ops.outer(aa, ba, M, P, Q, Cout, alpha=0.5, alpha_dev=sc, colsum=cs)
    ref = (aa.float().double().t() @ ba.float().double()).float() + 1.0
    assert _rel(Cout, ref) < 2e-5
    assert _rel(cs, aa.float().double().sum(0).float() + 1.0) < 2e-5   # fused bias-gradient column sums
    if planes == 2:
        assert _rel(Cout, (A.double().t() @ B.double()).float() + 1.0) < 3e-5


@pytest.mark.parametrize("M,d,ff", [(20000, 256, 512), (131072, 256, 512), (16500, 128, 256), (17000, 512, 512), (16400, 200, 300)])
def test_outer_group_matches_fp32(M, d, ff):
    """dsvg_outer_group: the four weight gradients of one block (in_proj, out_proj, linear1, linear2) in one launch, every
    problem with its bias column sums, accumulating into non-zero buffers; ragged tile edges at d = 128 / 200."""
    ops = _ops()
    shapes = [(3 * d, d), (d, d), (ff, d), (d, ff)]
    probs, refs = [], []
    for i, (P, Q) in enumerate(shapes):
        A, B = _rand(M, P, seed=10 + i), _rand(M, Q, seed=20 + i)
        aa = ops.act_from_float(A, 1, ld=(P + 7) // 8 * 8)
        ba = ops.act_from_float(B, 1, ld=(Q + 7) // 8 * 8)
        Cout = torch.full((P, Q), 0.5, device=DEV)
        cs = torch.full((P,), -1.0, device=DEV)
        probs.append((aa, ba, P, Q, Cout, cs))
        refs.append(((aa.float().double().t() @ ba.float().double()).float() + 0.5, aa.float().double().sum(0).float() - 1.0))
    ops.outer_group(probs, M)
    for (aa, ba, P, Q, Cout, cs), (rc, rs) in zip(probs, refs):
        assert _rel(Cout, rc) < 2e-5, (P, Q)
        assert _rel(cs, rs) < 2e-5, (P, Q)


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("D", [128, 256, 512])
def test_layernorm_fwd_bwd(D):
    ops = _ops()
    M = 777
    x, g, b = _rand(M, D, seed=1), 1 + 0.1 * _rand(D, seed=2), 0.1 * _rand(D, seed=3)
    y = ops.Act(M, D, 2, DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.ln_fwd(x, g, b, y, mean, rstd, M, D)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gr, br, 1e-5)
    assert _rel(y.float(), ref.detach()) < 1e-4
    dy, dx_in = _rand(M, D, seed=4), _rand(M, D, seed=5)
    dya = ops.act_from_float(dy, 2)
    ref.backward(dya.float())
    dx = torch.empty(M, D, device=DEV)
    dact = ops.Act(M, D, 2, DEV)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.ln_bwd(x, mean, rstd, g, M, D, dy=dya, dx_in=dx_in, dx_out=dx, dact=dact, dgamma=dg, dbeta=db)
    assert _rel(dx, xr.grad + dx_in) < 2e-5
    assert _rel(dact.float(), xr.grad + dx_in) < 1e-4
    assert _rel(dg, gr.grad) < 2e-5 and _rel(db, br.grad) < 2e-5


@pytest.mark.parametrize("M,K,rowvec,drop_p", [(16500, 256, False, 0.0), (16500, 512, True, 0.1), (33000, 256, True, 0.1),
                                               (20000, 512, False, 0.2)])
def test_linear_layernorm_fused_forward(M, K, rowvec, drop_p):
    """dsvg_linear_ln_fwd == dsvg_linear (residual-stream epilogue) followed by dsvg_ln_fwd: same fp32 residual stream bit
    for bit (same arithmetic, same dropout draws), LayerNorm output and statistics to rounding; and against plain fp32
    torch when dropout is off."""
    ops = _ops()
    N = 256
    assert ops.ln_fusable(M, N, 1) and not ops.ln_fusable(M, N, 2) and not ops.ln_fusable(4096, N, 1)
    X = ops.act_from_float(_rand(M, K, seed=1, scale=0.5), 1)
    W = ops.act_from_float(_rand(N, K, seed=2, scale=0.1), 1)
    bias, res = _rand(N, seed=3), _rand(M, N, seed=4, scale=2.0) + 1.5      # a non-zero row mean (variance by E[x^2]-m^2)
    g, b = 1 + 0.1 * _rand(N, seed=5), 0.1 * _rand(N, seed=6)
    L = 31
    rv = _rand((M + L - 1) // L, N, seed=7) if rowvec else None
    drop = (drop_p, 5, 1234) if drop_p > 0 else (0.0, 0, 0)
    kw = dict(bias=bias, drop=drop, rowvec=rv, rows_per_group=L if rowvec else 1, residual=res)
    x_a, x_b = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    y_a, y_b = ops.Act(M, N, 1, DEV), ops.Act(M, N, 1, DEV)
    m_a, r_a, m_b, r_b = (torch.empty(M, device=DEV) for _ in range(4))
    ops.linear(X, W, M, N, K, out_f32=x_a, **kw)
    ops.ln_fwd(x_a, g, b, y_a, m_a, r_a, M, N)
    ops.linear(X, W, M, N, K, out_f32=x_b, ln=(g, b, y_b, m_b, r_b), **kw)
    assert (x_a - x_b).abs().max().item() <= 2e-6 * x_a.abs().max().item()      # same arithmetic up to FMA contraction
    assert _rel(m_b, m_a) < 1e-5 and _rel(r_b, r_a) < 1e-4
    # the two outputs are bf16 roundings of values that agree to ~1e-6: at most one bf16 ulp apart
    assert ((y_a.float() - y_b.float()).abs() <= 2.0 ** -7 * y_a.float().abs() + 1e-6).all()
    assert (y_a.t != y_b.t).float().mean().item() < 0.02
    if drop_p == 0:
        ref = X.float() @ W.float().t() + bias + res
        if rowvec:
            ref = ref + rv[torch.arange(M, device=DEV) // L]
        assert _rel(x_b, ref) < 1e-5
        assert _rel(y_b.float(), F.layer_norm(ref, (N,), g, b, 1e-5)) < 1e-2


@pytest.mark.parametrize("M,K,drop_p,with_in,with_act", [(16500, 512, 0.0, True, True), (16500, 768, 0.1, True, True),
                                                         (33000, 512, 0.1, False, True), (20000, 768, 0.0, True, False)])
def test_linear_layernorm_fused_backward(M, K, drop_p, with_in, with_act):
    """dsvg_linear_ln_bwd: dy = dY . W^T stays on chip (fp32), LayerNorm backward in the epilogue -- against fp32 torch
    autograd of LayerNorm fed with the fp32 product, and the dropout zero pattern against the stand-alone ln_bwd kernel."""
    ops = _ops()
    N = 256
    dY = ops.act_from_float(_rand(M, K, seed=1, scale=0.5), 1)
    Wt = ops.act_from_float(_rand(N, K, seed=2, scale=0.1), 1)
    x = _rand(M, N, seed=3, scale=1.5) + 0.7
    g = 1 + 0.1 * _rand(N, seed=5)
    b = torch.zeros(N, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.ln_fwd(x, g, b, ops.Act(M, N, 1, DEV), mean, rstd, M, N)
    dx_in = _rand(M, N, seed=6) if with_in else None
    drop = (drop_p, 9, 77) if drop_p > 0 else (0.0, 0, 0)
    dx = torch.empty(M, N, device=DEV)
    dact = ops.Act(M, N, 1, DEV) if with_act else None
    dg, db = torch.zeros(N, device=DEV), torch.zeros(N, device=DEV)
    ops.linear_ln_bwd(dY, Wt, M, N, K, x, mean, rstd, g, dx_in=dx_in, dx_out=dx, dact=dact, drop=drop, dgamma=dg, dbeta=db)
    dy = dY.float() @ Wt.float().t()
    xr, gr, br = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xr, (N,), gr, br, 1e-5).backward(dy)
    want = xr.grad + (dx_in if with_in else 0)
    assert _rel(dx, want) < 5e-5
    assert _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4
    if with_act:
        dact2, dx2 = ops.Act(M, N, 1, DEV), torch.empty(M, N, device=DEV)
        ops.ln_bwd(x, mean, rstd, g, M, N, dy=ops.act_from_float(dy, 1), dx_in=dx_in, dx_out=dx2, dact=dact2, drop=drop)
        if drop_p > 0:
            z1, z2 = dact.float() == 0, dact2.float() == 0
            assert abs(z1.float().mean().item() - drop_p) < 0.01 and (z1 != z2).float().mean().item() < 1e-4
            keep = ~z1
            assert _rel(dact.float()[keep], (want / (1 - drop_p))[keep]) < 1e-2
        else:
            assert _rel(dact.float(), want) < 1e-2


def test_layernorm_pool_fwd_bwd():
    ops = _ops()
    nseq, L, D = 96, 31, 256
    M = nseq * L
    x, g, b = _rand(M, D, seed=1), 1 + 0.1 * _rand(D, seed=2), 0.1 * _rand(D, seed=3)
    lens = torch.randint(1, L + 1, (nseq,), generator=torch.Generator().manual_seed(9))
    valid = (torch.arange(L)[None, :] < lens[:, None]).to(torch.uint8).to(DEV).reshape(-1).contiguous()
    z = torch.empty(nseq, D, device=DEV)
    mean, rstd, ic = torch.empty(M, device=DEV), torch.empty(M, device=DEV), torch.empty(nseq, device=DEV)
    ops.ln_pool_fwd(x, g, b, valid, z, mean, rstd, ic, nseq, L, D)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    w = valid.float().reshape(nseq, L, 1)
    ref = (F.layer_norm(xr, (D,), gr, br, 1e-5).reshape(nseq, L, D) * w).sum(1) / w.sum(1)
    assert _rel(z, ref.detach()) < 2e-5
    dz = _rand(nseq, D, seed=4)
    ref.backward(dz)
    dx = torch.empty(M, D, device=DEV)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.ln_bwd(x, mean, rstd, g, M, D, dz=dz, valid=valid, inv_cnt=ic, L=L, dx_out=dx, dgamma=dg, dbeta=db)
    assert _rel(dx, xr.grad) < 2e-5
    assert _rel(dg, gr.grad) < 2e-5 and _rel(db, br.grad) < 2e-5


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("L,H,hd,masked,nseq", [(32, 8, 32, True, 37), (31, 8, 32, False, 37), (8, 8, 32, True, 37),
                                                (52, 4, 64, True, 37), (66, 8, 64, False, 37), (8, 4, 16, False, 37),
                                                (31, 8, 32, True, 1500), (66, 8, 64, True, 300), (52, 8, 32, True, 300),
                                                (16, 8, 64, True, 300)])
def test_attention_fwd_bwd(L, H, hd, masked, nseq):
    """Two-plane (parity mode) operands: 32 x 32 bf16x3 mma kernels (head_dim 32, L <= 32), general bf16x3 kernels
    (head_dim 32 / 64, L <= 80), fp32 SIMT for the rest (head_dim 16)."""
    ops = _ops()
    d = H * hd
    M = nseq * L
    qkv = _rand(M, 3 * d, seed=1, scale=0.7)
    qa = ops.act_from_float(qkv, 2)
    qv = qa.float().clone().requires_grad_(True)
    valid = None
    vmask = None
    if masked:
        lens = torch.randint(1, L + 1, (nseq,), generator=torch.Generator().manual_seed(3))
        vmask = (torch.arange(L)[None, :] < lens[:, None]).to(DEV)
        valid = vmask.to(torch.uint8).reshape(-1).contiguous()
    out = ops.Act(M, d, 2, DEV)
    ops.attn_fwd(qa, valid, out, nseq, L, H, hd, (0.0, 0, 0))
    q, k, v = (t.reshape(nseq, L, H, hd).transpose(1, 2) for t in qv.split(d, dim=-1))
    s = q @ k.transpose(-1, -2)
    if masked:
        s = s.masked_fill(~vmask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(M, d)
    assert _rel(out.float(), ref.detach()) < 1e-4
    do = _rand(M, d, seed=5)
    da = ops.act_from_float(do, 2)
    ref.backward(da.float())
    dqkv = ops.Act(M, 3 * d, 2, DEV)
    ops.attn_bwd(qa, valid, da, dqkv, nseq, L, H, hd, 1.0, (0.0, 0, 0))
    assert _rel(dqkv.float(), qv.grad) < 1e-4


@pytest.mark.parametrize("L,masked,nseq", [(32, True, 61), (31, False, 61), (8, True, 61), (17, True, 61),
                                           (31, True, 1500), (8, True, 2600)])
def test_attention_mma_fast_path(L, masked, nseq):
    """Single-plane bf16, head_dim 32, L <= 32 -> the mma.sync kernels; P and dS are rounded to bf16 inside.
    The large nseq cases make every CTA of the block-per-sequence kernels loop over several sequences."""
    ops = _ops()
    H, hd = 8, 32
    d, M = H * hd, nseq * L
    qa = ops.act_from_float(_rand(M, 3 * d, seed=1, scale=0.7), 1)
    qv = qa.float().clone().requires_grad_(True)
    valid = vmask = None
    if masked:
        lens = torch.randint(1, L + 1, (nseq,), generator=torch.Generator().manual_seed(3))
        vmask = (torch.arange(L)[None, :] < lens[:, None]).to(DEV)
        valid = vmask.to(torch.uint8).reshape(-1).contiguous()
    out = ops.Act(M, d, 1, DEV)
    ops.attn_fwd(qa, valid, out, nseq, L, H, hd, (0.0, 0, 0))
    q, k, v = (t.reshape(nseq, L, H, hd).transpose(1, 2) for t in qv.split(d, dim=-1))
    s = q @ k.transpose(-1, -2)
    if masked:
        s = s.masked_fill(~vmask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(M, d)
    assert _rel(out.float(), ref.detach()) < 1.5e-2
    da = ops.act_from_float(_rand(M, d, seed=5), 1)
    ref.backward(da.float())
    dqkv = ops.Act(M, 3 * d, 1, DEV)
    ops.attn_bwd(qa, valid, da, dqkv, nseq, L, H, hd, 0.5, (0.0, 0, 0))
    g = qv.grad.clone()
    g[:, :d] *= 0.5
    for lo, hi, nm in ((0, d, "dq"), (d, 2 * d, "dk"), (2 * d, 3 * d, "dv")):
        e = (dqkv.float()[:, lo:hi] - g[:, lo:hi]).norm() / g[:, lo:hi].norm()
        assert e.item() < 1.5e-2, (nm, e.item())


@pytest.mark.parametrize("L,H,hd,masked,nseq", [(52, 8, 32, True, 40), (51, 8, 32, False, 40), (66, 8, 64, True, 33),
                                                (65, 8, 64, False, 33), (16, 8, 64, True, 50), (33, 4, 32, True, 3000),
                                                (80, 2, 64, True, 7), (8, 8, 64, True, 64)])
def test_attention_general_tensor_core_path(L, H, hd, masked, nseq):
    """Single-plane bf16, head_dim 32 / 64, L <= 80 (one-stage fonts L = 52 / 51, scaled hierarchical L = 66 / 65 and
    16 group-level) -> attn_gmma kernels (CTA per (sequence, head), warp per 16-row query tile).  Same bound as the
    32 x 32 fast path: P and dS are rounded to bf16 inside."""
    ops = _ops()
    d, M = H * hd, nseq * L
    qa = ops.act_from_float(_rand(M, 3 * d, seed=1, scale=0.7), 1)
    qv = qa.float().clone().requires_grad_(True)
    valid = vmask = None
    if masked:
        lens = torch.randint(1, L + 1, (nseq,), generator=torch.Generator().manual_seed(3))
        vmask = (torch.arange(L)[None, :] < lens[:, None]).to(DEV)
        valid = vmask.to(torch.uint8).reshape(-1).contiguous()
    out = ops.Act(M, d, 1, DEV)
    ops.attn_fwd(qa, valid, out, nseq, L, H, hd, (0.0, 0, 0))
    q, k, v = (t.reshape(nseq, L, H, hd).transpose(1, 2) for t in qv.split(d, dim=-1))
    s = q @ k.transpose(-1, -2)
    if masked:
        s = s.masked_fill(~vmask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(M, d)
    assert _rel(out.float(), ref.detach()) < 1.5e-2
    da = ops.act_from_float(_rand(M, d, seed=5), 1)
    ref.backward(da.float())
    dqkv = ops.Act(M, 3 * d, 1, DEV)
    ops.attn_bwd(qa, valid, da, dqkv, nseq, L, H, hd, 0.5, (0.0, 0, 0))
    g = qv.grad.clone()
    g[:, :d] *= 0.5
    for lo, hi, nm in ((0, d, "dq"), (d, 2 * d, "dk"), (2 * d, 3 * d, "dv")):
        e = (dqkv.float()[:, lo:hi] - g[:, lo:hi]).norm() / g[:, lo:hi].norm()
        assert e.item() < 1.5e-2, (nm, e.item())


@pytest.mark.parametrize("planes,nseq", [(1, 9), (2, 9), (2, 700)])
def test_attention_general_path_dropout_consistent(planes, nseq):
    """planes = 2: the parity-mode (bf16x3) variant of the general kernels; nseq = 700: every CTA strides over several pairs."""
    ops = _ops()
    L, H, hd = 66, 8, 64
    d, M = H * hd, nseq * L
    qa = ops.act_from_float(_rand(M, 3 * d, seed=1, scale=0.5), planes)
    drop = (0.3, 11, 99)
    o1, o2, o0 = ops.Act(M, d, planes, DEV), ops.Act(M, d, planes, DEV), ops.Act(M, d, planes, DEV)
    ops.attn_fwd(qa, None, o1, nseq, L, H, hd, drop)
    ops.attn_fwd(qa, None, o2, nseq, L, H, hd, drop)
    ops.attn_fwd(qa, None, o0, nseq, L, H, hd, (0.0, 0, 0))
    assert torch.equal(o1.t, o2.t) and not torch.equal(o1.t, o0.t)
    ga = ops.act_from_float(_rand(M, d, seed=2), planes)
    dqkv = ops.Act(M, 3 * d, planes, DEV)
    ops.attn_bwd(qa, None, ga, dqkv, nseq, L, H, hd, 1.0, drop)
    lhs = (o1.float() * ga.float()).sum().item()
    rhs = (qa.float()[:, 2 * d:] * dqkv.float()[:, 2 * d:]).sum().item()
    assert abs(lhs - rhs) < (2e-3 if planes == 2 else 2e-2) * abs(lhs)
    # dropout keeps the mean: E[out] = out(no dropout)
    assert abs(o1.float().mean().item() - o0.float().mean().item()) < 5e-3 * o0.float().abs().mean().item() + 1e-4


@pytest.mark.parametrize("planes", [1, 2])
def test_attention_dropout_consistent_fwd_bwd(planes):
    """With dropout the backward must use the forward's mask: check d(out . w)/dv against finite structure."""
    ops = _ops()
    nseq, L, H, hd = 5, 32, 8, 32
    d, M = H * hd, nseq * L
    qkv = _rand(M, 3 * d, seed=1, scale=0.5)
    qa = ops.act_from_float(qkv, planes)
    drop = (0.3, 11, 99)
    o1, o2 = ops.Act(M, d, planes, DEV), ops.Act(M, d, planes, DEV)
    ops.attn_fwd(qa, None, o1, nseq, L, H, hd, drop)
    ops.attn_fwd(qa, None, o2, nseq, L, H, hd, drop)
    assert torch.equal(o1.t, o2.t)
    # out is linear in v for fixed probabilities+mask: out(v) . g == v . dv(g)
    g = _rand(M, d, seed=2)
    ga = ops.act_from_float(g, planes)
    dqkv = ops.Act(M, 3 * d, planes, DEV)
    ops.attn_bwd(qa, None, ga, dqkv, nseq, L, H, hd, 1.0, drop)
    lhs = (o1.float() * ga.float()).sum().item()
    rhs = (qa.float()[:, 2 * d:] * dqkv.float()[:, 2 * d:]).sum().item()
    assert abs(lhs - rhs) < (2e-3 if planes == 2 else 2e-2) * abs(lhs)
    keep = (o1.float().abs() > 0).float().mean().item()   # dropped probabilities thin the output but never zero a row
    assert keep > 0.99


# ------------------------------------------------------------------------------------------------ embedding
@pytest.mark.parametrize("use_grp", [False, True])
def test_embedding_fwd_bwd(use_grp):
    ops = _ops()
    from oracle import svg_oracle as O
    cfg = O.make_cfg("one_stage" if use_grp else "hierarchical", max_total_len=30)
    n = 6
    cmd, arg = O.synth_batch(cfg, n, seed=5)
    G, L = cmd.shape[1], cmd.shape[2]
    nseq, T, d, V, na = n * G, n * G * L, 256, 257, 11
    cmd, arg = cmd.to(DEV).contiguous(), arg.to(DEV).contiguous()
    Ec, Ea = _rand(7, d, seed=1), _rand(V, 64, seed=2)
    W, b = _rand(d, 64 * na, seed=3, scale=0.05), _rand(d, seed=4)
    Pt, Gt = _rand(L, d, seed=5), _rand(10, d, seed=6)
    grp = torch.empty(T, dtype=torch.uint8, device=DEV)
    ops.seq_prep(cmd, nseq, L, None, None, None, grp, None)
    table, base = torch.empty(na * V, d, device=DEV), torch.empty(d, device=DEV)
    ops.embed_fold(Ea, W, b, table, base, V, na, d)
    x = torch.empty(T, d, device=DEV)
    ops.embed_fwd(cmd, arg, grp if use_grp else None, Ec, table, base, Pt, Gt if use_grp else None, x, T, L, V, na, d,
                  (0.0, 0, 0))
    leaves = [t.clone().requires_grad_(True) for t in (Ec, Ea, W, b, Pt, Gt)]
    ec, ea, w, bb, pt, gt = leaves
    ci = cmd.long().reshape(-1)
    ref = ec[ci] + F.linear(ea[(arg + 1).long()].reshape(T, -1), w, bb) + pt.repeat(nseq, 1)
    if use_grp:
        ref = ref + gt[(cmd.long() == 0).cumsum(-1).reshape(-1)]
    assert _rel(x, ref.detach()) < 2e-5
    dx = _rand(T, d, seed=7)
    ref.backward(dx)
    grads = [torch.zeros_like(t) for t in (Ec, Ea, W, b, Pt, Gt)]
    scratch = torch.empty(na * V, d, device=DEV)
    ops.embed_bwd(cmd, arg, grp if use_grp else None, dx, Ea, W, grads[0], grads[4], grads[5] if use_grp else None,
                  grads[1], grads[2], grads[3], scratch, nseq, L, V, na, d, 10, (0.0, 0, 0))
    for name, got, leaf in zip("Ec Ea W b Pt Gt".split(), grads, leaves):
        if name == "Gt" and not use_grp:
            continue
        assert _rel(got, leaf.grad) < 5e-5, name


def test_seq_prep_matches_oracle_masks():
    ops = _ops()
    from oracle import svg_oracle as O
    cfg = O.make_cfg("hierarchical")
    cmd, _ = O.synth_batch(cfg, 16, seed=3)
    n, G, L = cmd.shape
    c = cmd.to(DEV).contiguous()
    nseq = n * G
    fe = torch.empty(nseq, dtype=torch.int32, device=DEV)
    vis = torch.empty(nseq, dtype=torch.uint8, device=DEV)
    kv = torch.empty(nseq * L, dtype=torch.uint8, device=DEV)
    grp = torch.empty(nseq * L, dtype=torch.uint8, device=DEV)
    counts = torch.zeros(2, device=DEV)
    ops.seq_prep(c, nseq, L, fe, vis, kv, grp, counts)
    ci = cmd.long()
    assert torch.equal(kv.cpu().bool().reshape(n, G, L), ~O.key_padding(ci))
    assert torch.equal(vis.cpu().bool().reshape(n, G), O.visibility(ci))
    assert torch.equal(grp.cpu().long().reshape(n, G, L), O.group_index(ci))
    wc = (O.extended_padding(ci) * O.visibility(ci).unsqueeze(-1).float())[..., 1:]
    wa = O.CMD_ARGS_MASK[ci[..., 1:]]
    assert counts[0].item() == wc.sum().item() and counts[1].item() == wa.sum().item()


def test_rows_embed_and_segsum():
    ops = _ops()
    nseq, L, d = 40, 31, 256
    R = nseq * L
    tab, add = _rand(L, d, seed=1), _rand(R, d, seed=2)
    x = torch.empty(R, d, device=DEV)
    ops.rows_embed_fwd(add, tab, x, R, L, d, (0.0, 0, 0))
    assert torch.allclose(x, add + tab.repeat(nseq, 1))
    dx = _rand(R, d, seed=3)
    dadd, dtab = torch.empty(R, d, device=DEV), torch.zeros(L, d, device=DEV)
    ops.rows_embed_bwd(dx, dadd, dtab, nseq, L, d, (0.0, 0, 0))
    assert torch.equal(dadd, dx) and _rel(dtab, dx.reshape(nseq, L, d).sum(0)) < 1e-5
    s = torch.empty(nseq, d, device=DEV)
    ops.seg_sum(dx, nseq, L, d, out_f32=s)
    assert _rel(s, dx.reshape(nseq, L, d).sum(1)) < 1e-5
    cs = torch.zeros(d, device=DEV)
    a = ops.act_from_float(dx, 2)
    ops.colsum(a, R, d, cs)
    assert _rel(cs, a.float().sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ loss
def test_cross_entropy_kernels():
    ops = _ops()
    from oracle import svg_oracle as O
    cfg = O.make_cfg("hierarchical")
    n = 8
    cmd, arg = O.synth_batch(cfg, n, seed=11)
    G, L = cmd.shape[1], cmd.shape[2]
    nseq, Ld = n * G, L - 1
    Md = nseq * Ld
    al = _rand(Md, 11 * 257, seed=1, scale=2.0)
    cl, vl = _rand(Md, 7, seed=2, scale=2.0), _rand(nseq, 2, seed=3)
    c, a = cmd.to(DEV).contiguous(), arg.to(DEV).contiguous()
    fe = torch.empty(nseq, dtype=torch.int32, device=DEV)
    vis = torch.empty(nseq, dtype=torch.uint8, device=DEV)
    counts, acc, out = torch.zeros(2, device=DEV), torch.zeros(8, device=DEV), torch.zeros(8, device=DEV)
    ops.seq_prep(c, nseq, L, fe, vis, None, None, counts)
    dla, dlc, dlv = ops.Act(Md, 2827, 2, DEV, ld=2880, zero=True), ops.Act(Md, 7, 2, DEV, ld=8), ops.Act(nseq, 2, 2, DEV, ld=8)
    ops.ce_args(al, 2827, c, a, counts, dla, acc, nseq, L, 11, 257)
    ops.ce_cmd(cl, c, fe, vis, counts, dlc, acc, nseq, L, 7)
    ops.ce_vis(vl, vis, dlv, acc, nseq, 1.0 / nseq)
    ops.loss_finalize(acc, counts, out, 1.0, 2.0, 1.0, 0.0, 0.1, 1.0 / nseq, 0.0, True, False)
    leaves = [t.clone().requires_grad_(True) for t in (al, cl, vl)]
    cfg.use_vae = False
    o = {"command_logits": leaves[1].reshape(n, G, Ld, 7), "args_logits": leaves[0].reshape(n, G, Ld, 11, 257),
         "visibility_logits": leaves[2].reshape(n, G, 1, 2), "tgt_commands": c, "tgt_args": a}
    O.CMD_ARGS_MASK = O.CMD_ARGS_MASK.to(DEV)
    try:
        ls = O.loss(o, cfg)
        for key, ten, leaf, dl in (("loss_args", al, leaves[0], dla), ("loss_cmd", cl, leaves[1], dlc),
                                   ("loss_visibility", vl, leaves[2], dlv)):
            g, = torch.autograd.grad(ls[key], leaf, retain_graph=True)
            assert _rel(dl.float(), g) < 1e-4, key
    finally:
        O.CMD_ARGS_MASK = O.CMD_ARGS_MASK.cpu()
    assert abs(out[0].item() - ls["loss"].item()) < 1e-5 * ls["loss"].item()
    assert abs(out[1].item() - ls["loss_cmd"].item()) < 1e-5 and abs(out[2].item() - ls["loss_args"].item()) < 1e-5
    assert abs(out[3].item() - ls["loss_visibility"].item()) < 1e-5
    assert dla.t[:, :, 2827:].abs().max().item() == 0  # K padding of the head dgrad operand stays zero


def test_vae_and_kl():
    ops = _ops()
    n, dz = 64, 256
    mu, ls, eps, dz_ = _rand(n, dz, seed=1), _rand(n, dz, seed=2, scale=0.3), _rand(n, dz, seed=3), _rand(n, dz, seed=4)
    z = torch.empty(n, dz, device=DEV)
    ops.vae_fwd(mu, ls, eps, z, n * dz)
    m, l = mu.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    zr = m + torch.exp(l / 2) * eps
    assert _rel(z, zr.detach()) < 1e-6
    kl = (-0.5 * torch.mean(1 + l - m.pow(2) - torch.exp(l))).clamp(min=0.1)
    ((zr * dz_).sum() + 3.0 * kl).backward()
    acc, out = torch.zeros(8, device=DEV), torch.zeros(8, device=DEV)
    ops.kl_sum(mu, ls, acc, n * dz)
    counts = torch.ones(2, device=DEV)
    ops.loss_finalize(acc, counts, out, 0.0, 0.0, 0.0, 1.0, 0.1, 0.0, 1.0 / (n * dz), False, True)
    assert abs(out[4].item() - kl.item()) < 1e-5
    dmu, dls = torch.empty(n, dz, device=DEV), torch.empty(n, dz, device=DEV)
    coef = torch.tensor([3.0], device=DEV)
    ops.vae_bwd(mu, ls, eps, dz_, coef, out, 1.0 / (n * dz), dmu, dls, n * dz)
    assert _rel(dmu, m.grad) < 1e-5 and _rel(dls, l.grad) < 1e-5


def test_cast_transpose_and_labels():
    ops = _ops()
    R, Cc = 258, 300
    x = _rand(R, Cc, seed=1)
    a, at = ops.Act(R, Cc, 2, DEV, ld=304), ops.Act(Cc, R, 2, DEV, ld=264)
    ops.cast_act(x, R, Cc, out=a, outT=at)
    assert _rel(a.float(), x) < 1e-5 and _rel(at.float(), x.t()) < 1e-5
    assert a.t[:, :, Cc:].abs().max().item() == 0 and at.t[:, :, R:].abs().max().item() == 0
    table = _rand(52, 64, seed=2)
    idx = torch.randint(0, 52, (40,), generator=torch.Generator().manual_seed(1)).to(DEV)
    g = ops.Act(40, 64, 2, DEV)
    ops.gather_rows(table, idx, 40, 64, g)
    assert _rel(g.float(), table[idx]) < 1e-5
    dt = torch.zeros(52, 64, device=DEV)
    gr = _rand(40, 64, seed=3)
    ops.scatter_rows(gr, idx, 40, 64, dt)
    assert _rel(dt, torch.zeros(52, 64, device=DEV).index_add_(0, idx, gr)) < 1e-5


def test_fused_adamw_matches_torch():
    """multi-tensor AdamW + global-norm clipping vs torch.optim.AdamW + clip_grad_norm_."""
    from deepsvg_b200 import FusedAdamW
    shapes = [(257, 64), (256,), (768, 256), (7, 256), (2827, 256), (1,)]
    ref = [torch.nn.Parameter(_rand(*s, seed=i)) for i, s in enumerate(shapes)]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.AdamW(ref, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    o_mine = FusedAdamW(mine, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0)
    for step in range(4):
        for i, (a, b) in enumerate(zip(ref, mine)):
            g = _rand(*a.shape, seed=100 * step + i, scale=0.3 if step % 2 else 3.0)
            a.grad, b.grad = g.clone(), g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step()
        o_mine.step()
        for a, b in zip(ref, mine):
            assert _rel(b.detach(), a.detach()) < 2e-6


# ------------------------------------------------------------------------------------------------ causal attention
@pytest.mark.parametrize("planes,L,H,hd", [(2, 31, 4, 32), (1, 31, 4, 32), (1, 32, 8, 32), (1, 51, 8, 32), (1, 66, 4, 64),
                                           (2, 51, 4, 32)])
def test_attention_causal_with_key_padding(planes, L, H, hd):
    """attn_mask = square_subsequent_mask plus key_padding_mask (the autoregressive decoder, model.py:264-269;
    functional.py:229-240) on all three kernels: fp32 SIMT (planes = 2), 32 x 32 mma (L <= 32, head_dim 32), general mma."""
    ops = _ops()
    nseq, d = 29, H * hd
    M = nseq * L
    qa = ops.act_from_float(_rand(M, 3 * d, seed=1, scale=0.7), planes)
    qv = qa.float().clone().requires_grad_(True)
    lens = torch.randint(1, L + 1, (nseq,), generator=torch.Generator().manual_seed(3))
    vmask = (torch.arange(L)[None, :] < lens[:, None]).to(DEV)
    valid = vmask.to(torch.uint8).reshape(-1).contiguous()
    out = ops.Act(M, d, planes, DEV)
    ops.attn_fwd(qa, valid, out, nseq, L, H, hd, (0.0, 0, 0), causal=True)
    q, k, v = (t.reshape(nseq, L, H, hd).transpose(1, 2) for t in qv.split(d, dim=-1))
    s = q @ k.transpose(-1, -2)
    s = s.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=DEV), 1), float("-inf"))
    s = s.masked_fill(~vmask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(M, d)
    tol = 1e-4 if planes == 2 else 1.5e-2
    assert _rel(out.float(), ref.detach()) < tol
    da = ops.act_from_float(_rand(M, d, seed=5), planes)
    ref.backward(da.float())
    dqkv = ops.Act(M, 3 * d, planes, DEV)
    ops.attn_bwd(qa, valid, da, dqkv, nseq, L, H, hd, 1.0, (0.0, 0, 0), causal=True)
    for lo, hi, nm in ((0, d, "dq"), (d, 2 * d, "dk"), (2 * d, 3 * d, "dv")):
        e = (dqkv.float()[:, lo:hi] - qv.grad[:, lo:hi]).norm() / qv.grad[:, lo:hi].norm()
        assert e.item() < tol, (nm, e.item())
