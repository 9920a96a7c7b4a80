"""Kernel-level parity on a real B200: every C-ABI kernel against a plain fp32 torch expression
(or, for integer work, exact equality).  Run with `pytest -m gpu`."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ark_b200 import ops  # noqa: E402
from ark_b200.layout import pack_layout  # noqa: E402

DEV = "cuda"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


# ----------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [
    (128, 128, 64), (128, 64, 128), (256, 384, 512), (16, 3072, 1024), (256, 3072, 1024),
    (200, 136, 72), (1000, 60943 // 8, 512), (37, 55, 512), (2560, 1024, 3072), (3072, 1024, 2560),
    (130, 24104, 128),
]


def _operand(rows, K, major, ld_pad=0, dtype=torch.bfloat16):
    """returns (storage tensor view with unit inner stride, logical [rows,K] fp32 matrix)"""
    if major == ops.MAJOR_K:
        buf = torch.randn(rows, K + ld_pad, device=DEV).to(dtype)
        view = buf[:, :K]
        return view, view.float()
    buf = torch.randn(K, rows + ld_pad, device=DEV).to(dtype)
    view = buf[:, :rows]
    return view, view.float().t()


@pytest.mark.parametrize("a_major", [ops.MAJOR_K, ops.MAJOR_MN])
@pytest.mark.parametrize("b_major", [ops.MAJOR_K, ops.MAJOR_MN])
@pytest.mark.parametrize("shape", GEMM_SHAPES)
def test_gemm_tc_matches_fp32(shape, a_major, b_major):
    M, N, K = shape
    torch.manual_seed(M * 7 + N * 3 + K)
    # MN-major operands need the M/N extent as leading dimension: pad to a multiple of 8
    A, Af = _operand(M, K, a_major, ld_pad=(-M) % 8 if a_major == ops.MAJOR_MN else (-K) % 8)
    B, Bf = _operand(N, K, b_major, ld_pad=(-N) % 8 if b_major == ops.MAJOR_MN else (-K) % 8)
    C = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(A, a_major, B, b_major, C, M, N, K, backend="tc")
    ref = Af @ Bf.t()
    assert torch.isfinite(C).all()
    assert _rel(C, ref) < 2e-3  # bf16 inputs are exact in fp32; only accumulation order differs


@pytest.mark.parametrize("epi", [ops.EPI_NONE, ops.EPI_GELU, ops.EPI_TANH])
@pytest.mark.parametrize("c_dtype", [torch.float32, torch.bfloat16])
def test_gemm_tc_epilogues(epi, c_dtype):
    M, N, K = 300, 200, 192
    torch.manual_seed(3)
    A, Af = _operand(M, K, ops.MAJOR_K)
    B, Bf = _operand(N, K, ops.MAJOR_K)
    bias = torch.randn(N, device=DEV)
    C = torch.zeros(M, N + 8, device=DEV, dtype=c_dtype)[:, :N]
    aux = torch.zeros(M, N + 8, device=DEV)[:, :N] if epi != ops.EPI_NONE else None
    ops.gemm(A, ops.MAJOR_K, B, ops.MAJOR_K, C, M, N, K, bias=bias, epilogue=epi, aux=aux, backend="tc")
    pre = Af @ Bf.t() / 1.0 + bias
    ref = {ops.EPI_NONE: pre, ops.EPI_GELU: torch.nn.functional.gelu(pre), ops.EPI_TANH: torch.tanh(pre)}[epi]
    tol = 1e-2 if c_dtype == torch.bfloat16 else 2e-3
    assert _rel(C, ref) < tol
    if aux is not None:
        assert _rel(aux, pre) < 2e-3


@pytest.mark.parametrize("M,N,with_bias", [(2048, 4736, True), (1100, 9000, False)])
def test_gemm_tc_persistent_bf16_tma_store_epilogue(M, N, with_bias):
    """>= 2 waves of 128x128 tiles with a plain bf16 output (the vocabulary projection): the persistent kernel whose
    full tiles leave through 128B-swizzled staging + TMA stores; ragged edges (M, N not multiples of 128) take the
    thread-written path inside the same launch."""
    K = 192
    g = torch.Generator(device=DEV).manual_seed(5)
    A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV, generator=g) if with_bias else None
    ldc = (N + 7) // 8 * 8
    C = torch.full((M, ldc), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, ops.MAJOR_K, B, ops.MAJOR_K, C, M, N, K, bias=bias, backend="tc")
    ref = A.float() @ B.float().t()
    if with_bias:
        ref = ref + bias
    torch.testing.assert_close(C[:, :N].float(), ref.to(torch.bfloat16).float(), rtol=2e-2, atol=2e-2)
    if ldc > N:
        assert (C[:, N:] == 7.0).all()          # the row padding is never written


def test_gemm_tc_accumulate_and_simt_agree():
    M, N, K = 130, 70, 1000
    torch.manual_seed(5)
    A, Af = _operand(M, K, ops.MAJOR_MN, ld_pad=(-M) % 8)
    B, Bf = _operand(N, K, ops.MAJOR_MN, ld_pad=(-N) % 8)
    base = torch.randn(M, N, device=DEV)
    C1, C2 = base.clone(), base.clone()
    ops.gemm(A, ops.MAJOR_MN, B, ops.MAJOR_MN, C1, M, N, K, accumulate=True, backend="tc")
    ops.gemm(A, ops.MAJOR_MN, B, ops.MAJOR_MN, C2, M, N, K, accumulate=True, backend="simt")
    ref = base + Af @ Bf.t()
    assert _rel(C1, ref) < 2e-3 and _rel(C2, ref) < 2e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_simt_awkward_shapes(dtype):
    for (M, N, K) in [(16, 512, 10), (256, 20, 1536), (5, 3, 1), (65, 65, 17)]:
        for am in (ops.MAJOR_K, ops.MAJOR_MN):
            for bm in (ops.MAJOR_K, ops.MAJOR_MN):
                A, Af = _operand(M, K, am, dtype=dtype)
                B, Bf = _operand(N, K, bm, dtype=dtype)
                bias = torch.randn(N, device=DEV)
                C = torch.empty(M, N, device=DEV)
                ops.gemm(A, am, B, bm, C, M, N, K, bias=bias, epilogue=ops.EPI_TANH, backend="simt")
                assert _rel(C, torch.tanh(Af @ Bf.t() + bias)) < 1e-5


# ----------------------------------------------------------------------------- encoder kernels
@pytest.mark.parametrize("pad", [False, True])
def test_gather_pool_fwd_bwd(pad):
    torch.manual_seed(0)
    B, T, d, nE, nR = 9, 7, 64, 50, 5
    pad_eid, pad_rid = (nE, nR) if pad else (None, None)
    E = torch.randn(nE + pad, d, device=DEV)
    R = torch.randn(nR + pad, d, device=DEV)
    tri = torch.stack([torch.randint(nE, (B, T)), torch.randint(nR, (B, T)), torch.randint(nE, (B, T))], -1)
    if pad:
        E[pad_eid] = 0
        R[pad_rid] = 0
        for b in range(B):
            n = (b * 3) % (T + 1)  # includes 0 valid triples and a full graph
            tri[b, n:] = torch.tensor([pad_eid, pad_rid, pad_eid])
    tri = tri.to(DEV)
    perm = torch.randperm(B).to(torch.int32).to(DEV)
    g = torch.empty(B, 3 * d, device=DEV)
    gb = torch.empty(B, 3 * d, device=DEV, dtype=torch.bfloat16)
    inv = torch.empty(B, device=DEV)
    ops.gather_pool_fwd(tri, perm, E, R, pad_rid, g, gb, inv)
    Er, Rr = E.clone().requires_grad_(), R.clone().requires_grad_()
    tp = tri[perm.long()]
    x = torch.cat([Er[tp[..., 0]], Rr[tp[..., 1]], Er[tp[..., 2]]], -1)
    if pad:
        m = tp[..., 1] != pad_rid
        ref = (x * m[..., None]).sum(1) / m.sum(1, keepdim=True).clamp(min=1)
    else:
        ref = x.mean(1)
    torch.testing.assert_close(g, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gb.float(), ref, rtol=1e-2, atol=1e-2)
    dg = torch.randn(B, 3 * d, device=DEV)
    ref.backward(dg)
    dE, dR = torch.zeros_like(E), torch.zeros_like(R)
    ops.gather_pool_bwd(dg, tri, perm, inv, pad_rid, pad_eid, dE, dR)
    gE, gR = Er.grad.clone(), Rr.grad.clone()
    if pad:
        gE[pad_eid] = 0
        gR[pad_rid] = 0
    torch.testing.assert_close(dE, gE, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dR, gR, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dz", [10, 24, 128])
def test_reparam_kl(dz):
    torch.manual_seed(1)
    B = 37
    heads = (torch.randn(B, 2 * dz, device=DEV) * 6).requires_grad_()  # logv beyond +-10 exercises the clamp
    eps = torch.randn(B, dz, device=DEV)
    perm = torch.randperm(B).to(torch.int32).to(DEV)
    z = torch.empty(B, dz, device=DEV)
    ldz = (dz + 7) // 8 * 8
    zb = torch.zeros(B, ldz, device=DEV, dtype=torch.bfloat16)
    kl = torch.zeros(1, device=DEV)
    kl_scale = 1.0 / (B * dz)
    ops.reparam_kl_fwd(heads.detach(), eps, perm, dz, True, kl_scale, z, zb, kl)
    mu, lv = heads[:, :dz], heads[:, dz:].clamp(-10, 10)
    e = eps[perm.long()]
    z_ref = mu + e * torch.exp(0.5 * lv)
    kl_ref = -0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp())
    torch.testing.assert_close(z, z_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(kl[0], kl_ref, rtol=1e-4, atol=1e-6)
    dz_in = torch.randn(B, dz, device=DEV)
    beta = 0.37
    ((z_ref * dz_in).sum() + beta * kl_ref).backward()
    dh = torch.empty(B, 2 * dz, device=DEV)
    ops.reparam_kl_bwd(heads.detach(), eps, perm, dz_in, dz, True, beta * kl_scale, dh, None)
    torch.testing.assert_close(dh, heads.grad, rtol=1e-4, atol=1e-5)


# ----------------------------------------------------------------------------- softmax CE
@pytest.mark.parametrize("V", [36, 55, 138, 24101, 60943])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_softmax_ce(V, dtype):
    torch.manual_seed(V)
    N = 67
    ldv = (V + 7) // 8 * 8
    logits = (torch.randn(N, ldv, device=DEV) * 3).to(dtype)
    tgt = torch.randint(1, V, (N,), device=DEV, dtype=torch.int32)
    tgt[0], tgt[1] = V - 1, 1
    ref_in = logits[:, :V].float().clone().requires_grad_()
    scale = 1.0 / 91.0
    loss_ref = torch.nn.functional.cross_entropy(ref_in, tgt.long(), reduction="sum") * scale
    loss_ref.backward()
    loss = torch.zeros(1, device=DEV)
    lse = torch.empty(N, device=DEV)
    work = logits.clone()
    ops.softmax_ce(work, V, tgt, scale, True, loss, lse)
    torch.testing.assert_close(loss[0], loss_ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lse, torch.logsumexp(ref_in.detach(), 1), rtol=1e-5, atol=1e-4)
    tol = 8e-3 if dtype == torch.bfloat16 else 1e-4
    assert _rel(work[:, :V], ref_in.grad) < tol
    assert (work[:, V:] == 0).all()
    # forward-only mode leaves the logits untouched
    work2 = logits.clone()
    loss2 = torch.zeros(1, device=DEV)
    ops.softmax_ce(work2, V, tgt, scale, False, loss2, None)
    assert torch.equal(work2, logits)
    torch.testing.assert_close(loss2, loss)


# ----------------------------------------------------------------------------- token ops (integer exact)
def test_pack_tokens_and_gather_scatter():
    torch.manual_seed(2)
    B, T, V, d = 11, 6, 40, 64
    seq = torch.zeros(B, 3 * T + 2, dtype=torch.int64)
    for b in range(B):
        n = (b % T) + 1
        toks = torch.randint(3, V, (3 * n,))
        seq[b, 0] = 1
        seq[b, 1:1 + 3 * n] = toks
        seq[b, 1 + 3 * n] = 2
    lay = pack_layout(seq).to(DEV)
    tok = torch.empty(lay.n_tok, dtype=torch.int32, device=DEV)
    tgt = torch.empty_like(tok)
    ops.pack_tokens(seq.to(DEV), lay.perm_dev, lay.bt_dev, lay.off_dev, lay.L, tok, tgt)
    exp_tok, exp_tgt = [], []
    for t in range(lay.L):
        for j in range(lay.bt[t]):
            exp_tok.append(int(seq[lay.perm[j], t]))
            exp_tgt.append(int(seq[lay.perm[j], t + 1]))
    assert tok.cpu().tolist() == exp_tok and tgt.cpu().tolist() == exp_tgt
    assert (tgt != 0).all() and lay.n_tok == int((seq[:, 1:] != 0).sum())
    W = torch.randn(V, d, device=DEV)
    Wb = W.to(torch.bfloat16)
    Xf = torch.empty(lay.n_tok, d, device=DEV)
    Xb = torch.empty(lay.n_tok, d, device=DEV, dtype=torch.bfloat16)
    ops.tok_gather_fwd(Wb, tok, Xf, Xb)
    assert torch.equal(Xb, Wb[tok.long()]) and torch.equal(Xf, Wb[tok.long()].float())
    ops.tok_gather_fwd(W, tok, Xf, None)
    assert torch.equal(Xf, W[tok.long()])
    dX = torch.randn(lay.n_tok, d, device=DEV)
    dW = torch.ones(V, d, device=DEV)
    ops.tok_scatter_add(dX, tok, dW)
    ref = torch.ones(V, d, device=DEV).index_add_(0, tok.long(), dX)
    torch.testing.assert_close(dW, ref, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------- elementwise + Adam
def test_elementwise_helpers():
    torch.manual_seed(4)
    n = 1003
    a, b = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    pre = a.clone().requires_grad_()
    torch.nn.functional.gelu(pre).backward(b)
    out = torch.empty(n, device=DEV)
    ops.gelu_bwd(b, a, out, None)
    torch.testing.assert_close(out, pre.grad, rtol=1e-4, atol=1e-5)
    h = torch.tanh(a)
    ops.tanh_bwd(b, h, out, None)
    torch.testing.assert_close(out, b * (1 - h * h), rtol=1e-5, atol=1e-6)
    X = torch.randn(777, 130, device=DEV)
    cs = torch.empty(130, device=DEV)
    ops.colsum(X, 777, 130, cs)
    torch.testing.assert_close(cs, X.sum(0), rtol=1e-4, atol=1e-4)
    Xb = X.to(torch.bfloat16)
    ops.colsum(Xb, 777, 130, cs)
    torch.testing.assert_close(cs, Xb.float().sum(0), rtol=1e-4, atol=1e-4)
    # ragged N inside a padded row (the [N_tok, ceil(V/8)*8] logits): vector path, NaN padding must not leak
    Xp = torch.full((1500, 136), float("nan"), device=DEV)
    Xp[:, :131] = torch.randn(1500, 131, device=DEV)
    cs = torch.empty(131, device=DEV)
    ops.colsum(Xp, 1500, 131, cs)
    torch.testing.assert_close(cs, Xp[:, :131].sum(0), rtol=1e-4, atol=1e-4)
    Xpb = Xp.to(torch.bfloat16)
    ops.colsum(Xpb, 1500, 131, cs)
    torch.testing.assert_close(cs, Xpb[:, :131].float().sum(0), rtol=1e-4, atol=2e-4)
    y = torch.empty(n, device=DEV)
    m = torch.empty(n, device=DEV, dtype=torch.uint8)
    ops.dropout_fwd(a, 0.25, 1234, 0, y, None, m)
    keep = m.bool()
    assert 0.65 < keep.float().mean().item() < 0.85
    torch.testing.assert_close(y, torch.where(keep, a / 0.75, torch.zeros_like(a)))
    dx = torch.empty(n, device=DEV)
    ops.dropout_bwd(b, m, 0.25, dx)
    torch.testing.assert_close(dx, torch.where(keep, b / 0.75, torch.zeros_like(b)))
    # vector paths (n % 4 == 0, aligned): the bf16 dropout draws the SAME mask as the f32 kernel (shared Philox
    # stream: the fused GRU epilogues rely on it) and the 4-wide backward agrees with the scalar one
    n4 = 4096
    a4, b4 = torch.randn(n4, device=DEV), torch.randn(n4, device=DEV)
    m32 = torch.empty(n4, device=DEV, dtype=torch.uint8)
    ops.dropout_fwd(a4, 0.25, 99, 7, torch.empty(n4, device=DEV), None, m32)
    xb = a4.to(torch.bfloat16)
    yb = torch.empty_like(xb)
    mb = torch.empty(n4, device=DEV, dtype=torch.uint8)
    ops.dropout_bf16(xb, 0.25, 99, 7, yb, mb)
    assert torch.equal(mb, m32)
    ref = torch.where(mb.bool(), (xb.float() / 0.75).to(torch.bfloat16), torch.zeros_like(xb))
    torch.testing.assert_close(yb.float(), ref.float(), rtol=1e-2, atol=1e-3)
    ops.dropout_bf16(xb[:1001], 0.25, 99, 7, yb[:1001], mb[:1001])          # ragged tail: scalar path, same draw
    assert torch.equal(mb[:1001], m32[:1001])
    dx4 = torch.empty(n4, device=DEV)
    ops.dropout_bwd(b4, mb, 0.25, dx4)
    torch.testing.assert_close(dx4, torch.where(mb.bool(), b4 / 0.75, torch.zeros_like(b4)))


def test_adam_flat_matches_torch():
    torch.manual_seed(6)
    n = 4099
    p0 = torch.randn(n, device=DEV)
    p = p0.clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sh = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    pt = p0.clone().requires_grad_()
    opt = torch.optim.Adam([pt], lr=1e-3)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV)
        pt.grad = g.clone()
        opt.step()
        ops.adam_flat(p, g, m, v, sh, 1e-3, 0.9, 0.999, 1e-8, step)
    torch.testing.assert_close(p, pt.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(sh, p.to(torch.bfloat16))


# ----------------------------------------------------------------------------- GRU layer
@pytest.mark.parametrize("use_tc", [0, 1])
def test_gru_layer_fwd_bwd_matches_torch(use_tc):
    torch.manual_seed(8)
    d, B = 64, 10
    lens = np.array([9, 9, 7, 7, 7, 4, 4, 2, 1, 1], dtype=np.int32)  # already sorted
    L = int(lens.max())
    bt = np.array([(lens > t).sum() for t in range(L)], dtype=np.int32)
    off = np.zeros(L + 1, dtype=np.int32)
    off[1:] = np.cumsum(bt)
    N = int(off[-1])
    gru = torch.nn.GRU(d, d, 1, batch_first=True).to(DEV)
    Wih, Whh = gru.weight_ih_l0.detach(), gru.weight_hh_l0.detach()
    # make weights/activations bf16-representable so both paths see the same numbers
    with torch.no_grad():
        for prm in gru.parameters():
            prm.copy_(prm.to(torch.bfloat16).float())
    x = torch.randn(B, L, d, device=DEV).to(torch.bfloat16).float().requires_grad_()
    h0 = torch.tanh(torch.randn(B, d, device=DEV)).to(torch.bfloat16).float().requires_grad_()
    y_ref, _ = gru(x, h0[None])
    dy_dense = torch.randn(B, L, d, device=DEV)
    mask = torch.from_numpy(lens).to(DEV)[:, None] > torch.arange(L, device=DEV)[None]
    (y_ref * dy_dense * mask[..., None]).sum().backward()

    rows = [(b, t) for t in range(L) for b in range(bt[t])]
    bi = torch.tensor([r[0] for r in rows], device=DEV)
    ti = torch.tensor([r[1] for r in rows], device=DEV)
    xp = x.detach()[bi, ti]
    gi = xp @ Wih.t() + gru.bias_ih_l0.detach()
    hp_f = torch.zeros(N, d, device=DEV)
    hp_b = torch.zeros(N, d, device=DEV, dtype=torch.bfloat16)
    hp_f[:B] = h0.detach()
    hp_b[:B] = h0.detach().to(torch.bfloat16)
    y = torch.empty(N, d, device=DEV)
    yb = torch.empty(N, d, device=DEV, dtype=torch.bfloat16)
    gates = tuple(torch.empty(N, d, device=DEV) for _ in range(4))
    ws = torch.empty(B, 3 * d, device=DEV)
    Wb = Whh.to(torch.bfloat16).contiguous()
    ops.gru_layer_fwd(hp_b, hp_f, Wb, gi.contiguous(), gru.bias_hh_l0.detach(), bt, off, L, d, y, yb, gates, ws, use_tc)
    assert _rel(y, y_ref.detach()[bi, ti]) < 1e-2   # h is re-rounded to bf16 each step on this path
    dy = dy_dense[bi, ti].contiguous()
    dgi = torch.empty(N, 3 * d, device=DEV, dtype=torch.bfloat16)
    dgh = torch.empty(N, 3 * d, device=DEV, dtype=torch.bfloat16)
    dha, dhb = torch.empty(B, d, device=DEV), torch.empty(B, d, device=DEV)
    dh0 = ops.gru_layer_bwd(dy, gates, hp_f, Wb, bt, off, L, d, dgi, dgh, dha, dhb, use_tc)
    assert _rel(dh0, h0.grad) < 2e-2
    dx = dgi.float() @ Wih
    assert _rel(dx, x.grad[bi, ti]) < 2e-2
    assert _rel(dgi.float().t() @ xp, gru.weight_ih_l0.grad) < 2e-2
    assert _rel(dgh.float().t() @ hp_f, gru.weight_hh_l0.grad) < 2e-2


# ----------------------------------------------------------------------------- persistent GRU (cooperative)
@pytest.mark.parametrize("d,lens", [
    (64, [9, 9, 7, 7, 7, 4, 4, 2, 1, 1]),                      # 4 slices, one ragged batch tile
    (128, [5] * 130),                                           # two batch tiles (130 rows), dense
    (256, list(range(40, 0, -1)) * 4),                          # 160 graphs, very ragged, tile 1 retires early
    (512, [30, 22, 22, 9, 3, 3, 3, 1]),                         # wd-articles-like small batch
    (1024, [10] * 256),                                         # syn-types shape: 64 slices x 2 tiles = 128 CTAs
    (512, [1 + (i * 7) % 12 for i in range(200)]),              # half-tile kernels, ragged: halves retire at different steps
    (512, [6] * 70 + [3] * 30),                                 # half-tile kernels: second half with 6 live rows at first
])
def test_gru_persist_fwd_bwd_matches_torch(d, lens):
    torch.manual_seed(d)
    lens = np.array(sorted(lens, reverse=True), dtype=np.int32)
    B, L = len(lens), int(lens.max())
    assert ops.gru_persist_supported(d, B) > 0
    bt = np.array([(lens > t).sum() for t in range(L)], dtype=np.int32)
    off = np.zeros(L + 1, dtype=np.int32)
    off[1:] = np.cumsum(bt)
    N = int(off[-1])
    gru = torch.nn.GRU(d, d, 1, batch_first=True).to(DEV)
    with torch.no_grad():
        for prm in gru.parameters():
            prm.copy_(prm.to(torch.bfloat16).float())
    Wih, Whh = gru.weight_ih_l0.detach(), gru.weight_hh_l0.detach()
    x = torch.randn(B, L, d, device=DEV).to(torch.bfloat16).float().requires_grad_()
    h0 = torch.tanh(torch.randn(B, d, device=DEV)).to(torch.bfloat16).float().requires_grad_()
    y_ref, _ = gru(x, h0[None])
    dy_dense = torch.randn(B, L, d, device=DEV)
    mask = torch.from_numpy(lens).to(DEV)[:, None] > torch.arange(L, device=DEV)[None]
    (y_ref * dy_dense * mask[..., None]).sum().backward()

    bi = torch.cat([torch.arange(int(bt[t])) for t in range(L)]).to(DEV)
    ti = torch.cat([torch.full((int(bt[t]),), t) for t in range(L)]).to(DEV)
    xp = x.detach()[bi, ti]
    gi = (xp @ Wih.t() + gru.bias_ih_l0.detach()).contiguous()
    bf = torch.bfloat16
    hp_b = torch.zeros(N, d, device=DEV, dtype=bf)
    hp_b[:B] = h0.detach().to(bf)
    y_b = torch.empty(N, d, device=DEV, dtype=bf)
    gates = tuple(torch.empty(N, d, device=DEV, dtype=bf) for _ in range(4))
    sync = torch.empty(2 * ((B + 127) // 128), device=DEV, dtype=torch.int32)
    bt_d, off_d = torch.from_numpy(bt).to(DEV), torch.from_numpy(off[:-1].copy()).to(DEV)
    Wb = Whh.to(bf).contiguous()
    ops.gru_persist_fwd(hp_b, h0.detach().contiguous(), Wb, gi, gru.bias_hh_l0.detach(), bt_d, off_d, L, B, d, y_b,
                        gates, sync)
    torch.cuda.synchronize()
    assert _rel(y_b, y_ref.detach()[bi, ti]) < 1e-2
    # the packed h_prev rows of step t+1 are the outputs of step t (prefix of the sorted batch)
    for t in range(L - 1):
        n = int(bt[t + 1])
        assert torch.equal(hp_b[off[t + 1]:off[t + 1] + n], y_b[off[t]:off[t] + n])
    WT = torch.empty(d, 3 * d, device=DEV, dtype=bf)
    ops.transpose_bf16(Wb, WT)
    assert torch.equal(WT, Wb.t().contiguous())
    dy = dy_dense[bi, ti].contiguous()
    dgi = torch.empty(N, 3 * d, device=DEV, dtype=bf)
    dgh = torch.empty(N, 3 * d, device=DEV, dtype=bf)
    dh0 = torch.full((B, d), 7.0, device=DEV)
    ops.gru_persist_bwd(dy, gates, hp_b, WT, bt_d, off_d, L, B, d, dgi, dgh, dh0, False, sync, Whh_b=Wb)
    torch.cuda.synchronize()
    assert _rel(dh0, h0.grad) < 3e-2
    assert _rel(dgi.float() @ Wih, x.grad[bi, ti]) < 3e-2
    assert _rel(dgi.float().t() @ xp, gru.weight_ih_l0.grad) < 3e-2
    assert _rel(dgh.float().t() @ hp_b.float(), gru.weight_hh_l0.grad) < 3e-2
    # accumulate mode adds to what is there
    dh0b = dh0.clone()
    ops.gru_persist_bwd(dy, gates, hp_b, WT, bt_d, off_d, L, B, d, dgi, dgh, dh0b, True, sync, Whh_b=Wb)
    torch.testing.assert_close(dh0b, 2 * dh0, rtol=1e-5, atol=1e-6)
    if d >= 512:     # both backward kernels (K-split clusters with untransposed W_hh / N-sliced with W_hh^T) agree
        assert ops.gru_persist_bwd_ksplit(d, B)
        dgi2, dgh2, dh02 = torch.empty_like(dgi), torch.empty_like(dgh), torch.empty_like(dh0)
        ops.gru_persist_bwd(dy, gates, hp_b, WT, bt_d, off_d, L, B, d, dgi2, dgh2, dh02, False, sync)   # no Whh_b -> N-sliced
        assert _rel(dgh2, dgh) < 5e-3 and _rel(dgi2, dgi) < 5e-3 and _rel(dh02, dh0) < 5e-3
    # fused inter-layer dropout == the stand-alone kernels with the same Philox stream
    y_d, mask = torch.empty_like(y_b), torch.empty(N, d, device=DEV, dtype=torch.uint8)
    hp2 = hp_b.clone()
    ops.gru_persist_fwd(hp2, h0.detach().contiguous(), Wb, gi, gru.bias_hh_l0.detach(), bt_d, off_d, L, B, d, y_d,
                        gates, sync, mask=mask, p_drop=0.25, seed=77, offset=1000)
    y_ref2, mask_ref = torch.empty_like(y_b), torch.empty_like(mask)
    ops.dropout_bf16(y_b, 0.25, 77, 1000, y_ref2, mask_ref)
    assert torch.equal(mask, mask_ref) and torch.equal(y_d, y_ref2) and torch.equal(hp2, hp_b)
    dy_m = torch.empty_like(dy)
    ops.dropout_bwd(dy, mask, 0.25, dy_m)
    dgi_a, dgh_a, dh0_a = torch.empty_like(dgi), torch.empty_like(dgh), torch.empty_like(dh0)
    dgi_b_, dgh_b_, dh0_b_ = torch.empty_like(dgi), torch.empty_like(dgh), torch.empty_like(dh0)
    ops.gru_persist_bwd(dy_m, gates, hp_b, WT, bt_d, off_d, L, B, d, dgi_a, dgh_a, dh0_a, False, sync, Whh_b=Wb)
    ops.gru_persist_bwd(dy, gates, hp_b, WT, bt_d, off_d, L, B, d, dgi_b_, dgh_b_, dh0_b_, False, sync, Whh_b=Wb,
                        dy_mask=mask, p_drop=0.25)
    assert torch.equal(dgi_a, dgi_b_) and torch.equal(dgh_a, dgh_b_) and torch.equal(dh0_a, dh0_b_)
