"""Thin torch-tensor wrappers over the C ABI (include/arkb200.h).

Each function checks dtype/device/contiguity, extracts raw device pointers and the CURRENT stream, and
enqueues one library call.  torch is used only for memory and streams.  No fallbacks: a non-CUDA tensor
is an error.
"""
from __future__ import annotations

import ctypes

import torch

from . import _C
from ._C import BF16, EPI_GELU, EPI_NONE, EPI_RELU, EPI_TANH, F32, MAJOR_K, MAJOR_MN  # noqa: F401

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _ptr(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise _C.ArkError("ark_b200 kernels need CUDA tensors (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise _C.ArkError(f"expected {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())


def _host_i32(a):
    """HOST int32 array (numpy) -> pointer; the C side reads it synchronously during the call."""
    import numpy as np
    if not (isinstance(a, np.ndarray) and a.dtype == np.int32 and a.flags.c_contiguous):
        raise _C.ArkError("host metadata must be a contiguous numpy int32 array")
    return ctypes.c_void_p(a.ctypes.data)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _contig(*ts):
    for t in ts:
        if t is not None and not t.is_contiguous():
            raise _C.ArkError("tensor must be contiguous")


def gather_pool_fwd(triples, perm, E, R, pad_rid, g=None, g_bf16=None, inv_cnt=None):
    B, T, _ = triples.shape
    d = E.shape[1]
    _contig(triples, perm, E, R, g, g_bf16, inv_cnt)
    _C.lib().call("ark_gather_pool_fwd", _ptr(triples, torch.int64), _ptr(perm, torch.int32), _ptr(E, torch.float32),
                  _ptr(R, torch.float32), B, T, d, -1 if pad_rid is None else int(pad_rid), _ptr(g, torch.float32),
                  _ptr(g_bf16, torch.bfloat16), _ptr(inv_cnt, torch.float32), _stream())


def gather_pool_bwd(dg, triples, perm, inv_cnt, pad_rid, pad_eid, dE, dR):
    B, T, _ = triples.shape
    d = dE.shape[1]
    _contig(dg, triples, perm, inv_cnt, dE, dR)
    _C.lib().call("ark_gather_pool_bwd", _ptr(dg, torch.float32), _ptr(triples, torch.int64), _ptr(perm, torch.int32),
                  _ptr(inv_cnt, torch.float32), B, T, d, -1 if pad_rid is None else int(pad_rid),
                  -1 if pad_eid is None else int(pad_eid), dR.shape[0], _ptr(dE, torch.float32), _ptr(dR, torch.float32),
                  _stream())


def pack_tokens(seq, perm, bt, off, L, tok_in, tgt, row_t=None):
    B, seq_len = seq.shape
    _contig(seq, perm, bt, off, tok_in, tgt, row_t)
    _C.lib().call("ark_pack_tokens", _ptr(seq, torch.int64), _ptr(perm, torch.int32), _ptr(bt, torch.int32),
                  _ptr(off, torch.int32), B, seq_len, L, _ptr(tok_in, torch.int32), _ptr(tgt, torch.int32),
                  _ptr(row_t, torch.int32), _stream())


def tok_pos_gather_fwd(W, P, tok, pos, X_bf16):
    _contig(W, P, tok, pos, X_bf16)
    _C.lib().call("ark_tok_pos_gather_fwd", _ptr(W, torch.bfloat16), _ptr(P, torch.bfloat16), _ptr(tok, torch.int32),
                  _ptr(pos, torch.int32), tok.numel(), W.shape[1], _ptr(X_bf16, torch.bfloat16), _stream())


def tok_gather_fwd(W, tok, X_f32=None, X_bf16=None):
    V, d = W.shape
    _contig(W, tok, X_f32, X_bf16)
    _C.lib().call("ark_tok_gather_fwd", _ptr(W), _DT[W.dtype], _ptr(tok, torch.int32), tok.numel(), d, V,
                  _ptr(X_f32, torch.float32), _ptr(X_bf16, torch.bfloat16), _stream())


def tok_scatter_add(dX, tok, dW):
    V, d = dW.shape
    _contig(dX, tok, dW)
    _C.lib().call("ark_tok_scatter_add", _ptr(dX, torch.float32), _ptr(tok, torch.int32), tok.numel(), d, V,
                  _ptr(dW, torch.float32), _stream())


def reparam_kl_fwd(heads, eps, perm, dz, clamp, kl_scale, z, z_bf16, kl_acc):
    B = heads.shape[0]
    _contig(eps, perm, z, z_bf16)
    _C.lib().call("ark_reparam_kl_fwd", _ptr(heads, torch.float32), heads.stride(0), _ptr(eps, torch.float32),
                  _ptr(perm, torch.int32), B, dz, int(clamp), float(kl_scale), _ptr(z, torch.float32),
                  _ptr(z_bf16, torch.bfloat16), 0 if z_bf16 is None else z_bf16.stride(0),
                  _ptr(kl_acc, torch.float32), _stream())


def reparam_kl_bwd(heads, eps, perm, dz_in, dz, clamp, beta_kl_scale, dheads, dheads_bf16, beta_dev=None, dmu_ext=None,
                   dlogv_ext=None):
    B = heads.shape[0]
    _contig(eps, perm, dz_in, dmu_ext, dlogv_ext)
    ld = dheads.stride(0) if dheads is not None else dheads_bf16.stride(0)
    if dheads is not None and dheads_bf16 is not None and dheads.stride(0) != dheads_bf16.stride(0):
        raise _C.ArkError("dheads and dheads_bf16 must share a row stride")
    _C.lib().call("ark_reparam_kl_bwd", _ptr(heads, torch.float32), heads.stride(0), _ptr(eps, torch.float32),
                  _ptr(perm, torch.int32), _ptr(dz_in, torch.float32), B, dz, int(clamp), float(beta_kl_scale),
                  _ptr(beta_dev, torch.float32), _ptr(dmu_ext, torch.float32), _ptr(dlogv_ext, torch.float32),
                  _ptr(dheads, torch.float32), _ptr(dheads_bf16, torch.bfloat16), ld, _stream())


def softmax_ce(logits, V, tgt, grad_scale, write_grad, loss_acc=None, lse=None):
    N, ldv = logits.shape[0], logits.stride(0)
    _contig(tgt, lse)
    _C.lib().call("ark_softmax_ce", _ptr(logits), _DT[logits.dtype], N, V, ldv, _ptr(tgt, torch.int32),
                  float(grad_scale), int(write_grad), _ptr(loss_acc, torch.float32), _ptr(lse, torch.float32), _stream())


def _ld(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise _C.ArkError("GEMM operands must be 2-D with unit inner stride")
    return t.stride(0)


def gemm(A, a_major, B, b_major, C, M, N, K, bias=None, epilogue=EPI_NONE, accumulate=False, aux=None,
         backend="tc"):
    """C[M,N] = epi(A.B^T + bias).  A is [M,K] (MAJOR_K) or [K,M] (MAJOR_MN); B is [N,K] or [K,N].

    backend "tc" = tcgen05/TMA kernel (bf16 operands, 16-byte aligned, ld % 8 == 0); "simt" = fp32-FMA
    kernel (any dtype pair / alignment).  The choice is the caller's; nothing is substituted silently.
    """
    if A.dtype != B.dtype:
        raise _C.ArkError("GEMM operands must share a dtype")
    if aux is not None and (aux.stride(0) != C.stride(0)):
        raise _C.ArkError("aux must share C's row stride")
    lda, ldb, ldc = _ld(A), _ld(B), _ld(C)
    if backend == "tc":
        _C.lib().call("ark_gemm_bf16_tc", _ptr(A, torch.bfloat16), a_major, lda, _ptr(B, torch.bfloat16), b_major, ldb,
                      _ptr(C), _DT[C.dtype], ldc, M, N, K, _ptr(bias, torch.float32), epilogue, int(accumulate),
                      _ptr(aux, torch.float32), _stream())
    elif backend == "simt":
        _C.lib().call("ark_gemm_simt", _ptr(A), a_major, lda, _ptr(B), b_major, ldb, _DT[A.dtype], _ptr(C),
                      _DT[C.dtype], ldc, M, N, K, _ptr(bias, torch.float32), epilogue, int(accumulate),
                      _ptr(aux, torch.float32), _stream())
    else:
        raise _C.ArkError(f"unknown GEMM backend {backend!r}")


def tc_eligible(A, B) -> bool:
    """True when both operands satisfy the TMA constraints of the tensor-core GEMM."""
    return (A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.stride(0) % 8 == 0 and B.stride(0) % 8 == 0
            and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0)


def gru_layer_fwd(hp_bf16, hp_f32, Whh, gi, b_hh, bt_host, off_host, L, d, y, y_bf16, gates, gh_ws, use_tc):
    r, z, n, ghn = gates if gates is not None else (None, None, None, None)
    _C.lib().call("ark_gru_layer_fwd", _ptr(hp_bf16, torch.bfloat16), _ptr(hp_f32, torch.float32), _ptr(Whh),
                  _DT[Whh.dtype], _ptr(gi, torch.float32), _ptr(b_hh, torch.float32), _host_i32(bt_host), _host_i32(off_host), L, d,
                  _ptr(y, torch.float32), _ptr(y_bf16, torch.bfloat16), _ptr(r), _ptr(z), _ptr(n), _ptr(ghn),
                  _ptr(gh_ws, torch.float32), int(use_tc), _stream())


def gru_layer_bwd(dy, gates, hp_f32, Whh, bt_host, off_host, L, d, dgi, dgh, dh_a, dh_b, use_tc):
    """Returns the tensor (dh_a or dh_b) that holds d(loss)/d(h0) [bt[0], d]."""
    r, z, n, ghn = gates
    out = ctypes.c_void_p()
    _C.lib().call("ark_gru_layer_bwd", _ptr(dy, torch.float32), _ptr(r), _ptr(z), _ptr(n), _ptr(ghn),
                  _ptr(hp_f32, torch.float32), _ptr(Whh), _DT[Whh.dtype], _host_i32(bt_host), _host_i32(off_host), L, d, _ptr(dgi), _ptr(dgh),
                  _ptr(dh_a, torch.float32), _ptr(dh_b, torch.float32), ctypes.byref(out), int(use_tc), _stream())
    return dh_a if out.value == dh_a.data_ptr() else dh_b


def gelu_bwd(dact, pre, dpre=None, dpre_bf16=None):
    _contig(dact, pre, dpre, dpre_bf16)
    _C.lib().call("ark_gelu_bwd", _ptr(dact, torch.float32), _ptr(pre, torch.float32), dact.numel(),
                  _ptr(dpre, torch.float32), _ptr(dpre_bf16, torch.bfloat16), _stream())


def tanh_bwd(dh, h, dpre=None, dpre_bf16=None):
    _contig(dh, h, dpre, dpre_bf16)
    _C.lib().call("ark_tanh_bwd", _ptr(dh, torch.float32), _ptr(h, torch.float32), dh.numel(),
                  _ptr(dpre, torch.float32), _ptr(dpre_bf16, torch.bfloat16), _stream())


def colsum(X, M, N, out, accumulate=False, deterministic=False):
    """deterministic: fixed summation order (ranks that redo the same sum must agree bitwise)."""
    _C.lib().call("ark_colsum", _ptr(X), _DT[X.dtype], M, N, X.stride(0), _ptr(out, torch.float32),
                  2 if deterministic else int(accumulate), _stream())


def add_(a, b, y=None, y_bf16=None):
    _contig(a, b, y, y_bf16)
    _C.lib().call("ark_add_f32", _ptr(a, torch.float32), _ptr(b, torch.float32), a.numel(), _ptr(y, torch.float32),
                  _ptr(y_bf16, torch.bfloat16), _stream())


def cast_bf16(x, y):
    _contig(x, y)
    _C.lib().call("ark_cast_f32_to_bf16", _ptr(x, torch.float32), x.numel(), _ptr(y, torch.bfloat16), _stream())


def dropout_fwd(x, p, seed, offset, y=None, y_bf16=None, mask=None, offset_dev=None):
    _contig(x, y, y_bf16, mask)
    _C.lib().call("ark_dropout_fwd", _ptr(x, torch.float32), x.numel(), float(p), int(seed), int(offset),
                  _ptr(y, torch.float32), _ptr(y_bf16, torch.bfloat16), _ptr(mask, torch.uint8),
                  _ptr(offset_dev, torch.int64), _stream())


def dropout_bwd(dy, mask, p, dx):
    _contig(dy, mask, dx)
    _C.lib().call("ark_dropout_bwd", _ptr(dy, torch.float32), _ptr(mask, torch.uint8), dy.numel(), float(p),
                  _ptr(dx, torch.float32), _stream())


def adam_flat(p, g, m, v, shadow, lr, beta1, beta2, eps, step, grad_scale=1.0):
    _contig(p, g, m, v, shadow)
    _C.lib().call("ark_adam_flat", _ptr(p, torch.float32), _ptr(g, torch.float32), _ptr(m, torch.float32),
                  _ptr(v, torch.float32), _ptr(shadow, torch.bfloat16), p.numel(), float(lr), float(beta1),
                  float(beta2), float(eps), int(step), float(grad_scale), _stream())


def gru_persist_supported(d, bt0) -> int:
    return int(_C.lib().raw("ark_gru_persist_supported")(int(d), int(bt0)))


def transpose_bf16(x, out):
    R, C = x.shape
    _contig(x, out)
    _C.lib().call("ark_transpose_bf16", _ptr(x, torch.bfloat16), R, C, _ptr(out, torch.bfloat16), _stream())


def gru_persist_fwd(hp_b, h0, Whh_b, gi, b_hh, bt_dev, off_dev, L, bt0, d, y_b, gates, sync_ws, mask=None, p_drop=0.0,
                    seed=0, offset=0, offset_dev=None):
    r, z, n, ghn = gates if gates is not None else (None, None, None, None)
    _contig(hp_b, h0, Whh_b, gi, y_b, mask)
    _C.lib().call("ark_gru_persist_fwd", _ptr(hp_b, torch.bfloat16), _ptr(h0, torch.float32), _ptr(Whh_b, torch.bfloat16),
                  _ptr(gi, torch.float32), _ptr(b_hh, torch.float32), _ptr(bt_dev, torch.int32), _ptr(off_dev, torch.int32),
                  L, bt0, hp_b.shape[0], d, _ptr(y_b, torch.bfloat16), _ptr(r, torch.bfloat16), _ptr(z, torch.bfloat16),
                  _ptr(n, torch.bfloat16), _ptr(ghn, torch.bfloat16), _ptr(mask, torch.uint8), float(p_drop), int(seed),
                  int(offset), _ptr(offset_dev, torch.int64), _ptr(sync_ws, torch.int32), _stream())


def gru_persist_bwd_ksplit(d, bt0) -> bool:
    """True when the backward runs the K-split cluster kernel, which reads W_hh untransposed."""
    return bool(_C.lib().raw("ark_gru_persist_bwd_ksplit")(int(d), int(bt0)))


def gru_persist_bwd(dy, gates, hp_b, WhhT_b, bt_dev, off_dev, L, bt0, d, dgi_b, dgh_b, dh0, accumulate, sync_ws, Whh_b=None,
                    dy_mask=None, p_drop=0.0):
    r, z, n, ghn = gates
    _contig(dy, hp_b, WhhT_b, Whh_b, dgi_b, dgh_b, dh0, dy_mask)
    _C.lib().call("ark_gru_persist_bwd", _ptr(dy, torch.float32), _ptr(r, torch.bfloat16), _ptr(z, torch.bfloat16),
                  _ptr(n, torch.bfloat16), _ptr(ghn, torch.bfloat16), _ptr(hp_b, torch.bfloat16),
                  _ptr(WhhT_b, torch.bfloat16), _ptr(bt_dev, torch.int32), _ptr(off_dev, torch.int32), L, bt0,
                  hp_b.shape[0], d, _ptr(dgi_b, torch.bfloat16), _ptr(dgh_b, torch.bfloat16), _ptr(dh0, torch.float32),
                  int(accumulate), _ptr(Whh_b, torch.bfloat16), _ptr(dy_mask, torch.uint8), float(p_drop),
                  _ptr(sync_ws, torch.int32), _stream())


def dropout_bf16(x, p, seed, offset, y, mask=None, offset_dev=None):
    _contig(x, y, mask)
    _C.lib().call("ark_dropout_bf16", _ptr(x, torch.bfloat16), x.numel(), float(p), int(seed), int(offset),
                  _ptr(y, torch.bfloat16), _ptr(mask, torch.uint8), _ptr(offset_dev, torch.int64), _stream())


def adam_flat_dyn(p, g, m, v, shadow, hyper, beta1, beta2, eps, grad_scale=1.0):
    _contig(p, g, m, v, shadow, hyper)
    _C.lib().call("ark_adam_flat_dyn", _ptr(p, torch.float32), _ptr(g, torch.float32), _ptr(m, torch.float32),
                  _ptr(v, torch.float32), _ptr(shadow, torch.bfloat16), p.numel(), _ptr(hyper, torch.float32),
                  float(beta1), float(beta2), float(eps), float(grad_scale), _stream())


def dp_reduce_adam(sym, spans, m, v, mode, lr=0.0, beta1=0.9, beta2=0.999, eps=1e-8, step=1, hyper=None, ctas=0):
    """K11 (csrc/dp_reduce.cu): switch-reduced gradients -> Adam on this rank's 1/world slice -> multicast parameters.
    sym: ark_b200.symm.SymmFlat; spans: [(begin, end)] in elements (multiples of 4); mode 0 = gradient all-reduce only."""
    import numpy as np
    sb = np.ascontiguousarray([s_ for s_, _ in spans], dtype=np.int64)
    se = np.ascontiguousarray([e_ for _, e_ in spans], dtype=np.int64)
    flags = (ctypes.c_void_p * sym.world)(*sym.peer_flags)
    _contig(m, v, hyper)
    _C.lib().call("ark_dp_reduce_adam", ctypes.c_void_p(sym.mc_grad), ctypes.c_void_p(sym.mc_param),
                  ctypes.c_void_p(sym.mc_shadow), _ptr(sym.param, torch.float32), _ptr(m, torch.float32),
                  _ptr(v, torch.float32), flags, sym.rank, sym.world, sb.ctypes.data_as(ctypes.c_void_p),
                  se.ctypes.data_as(ctypes.c_void_p), len(spans), int(mode), float(lr), float(beta1), float(beta2),
                  float(eps), int(step), _ptr(hyper, torch.float32), int(ctas), _stream())


def dp_allgather_mc(sym, ws, items, ctas=0):
    """All-gather by multicast store (csrc/dp_reduce.cu).  ws: ark_b200.symm.SymmBuf holding the gathered buffers;
    items: [(src tensor, byte offset of THIS rank's slot in ws)]."""
    import numpy as np
    n = len(items)
    src = (ctypes.c_void_p * n)(*[t.data_ptr() for t, _ in items])
    dst = (ctypes.c_void_p * n)(*[ws.mc + int(o) for _, o in items])
    nb = np.ascontiguousarray([t.numel() * t.element_size() for t, _ in items], dtype=np.int64)
    _contig(*[t for t, _ in items])
    flags = (ctypes.c_void_p * sym.world)(*sym.peer_flags)
    _C.lib().call("ark_dp_allgather_mc", src, dst, nb.ctypes.data_as(ctypes.c_void_p), n, flags, sym.rank, sym.world,
                  int(ctas), _stream())


def gru_wave_supported(d, bt0, nl) -> int:
    return int(_C.lib().raw("ark_gru_wave_supported")(int(d), int(bt0), int(nl)))


def _ptr_array(tensors, dtype):
    """HOST array of device pointers (const T* const*)."""
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = _ptr(t, dtype) if t is not None else None
    return arr


def gru_wave_fwd(x_b, hp_b, out_b, h0, Wih, Whh, b_ih, b_hh, bt_dev, off_dev, L, bt0, d, gates, mask, p_drop, seed,
                 offset, offset_dev, sync_ws):
    """All layers x all steps of the GRU stack in one cooperative launch (csrc/gru_wave.cu)."""
    nl, N = hp_b.shape[0], hp_b.shape[1]
    r, z, n, ghn = gates if gates is not None else (None, None, None, None)
    _contig(x_b, hp_b, out_b, h0, r, z, n, ghn, mask, *Wih, *Whh)
    a_ih, a_hh = _ptr_array(Wih, torch.bfloat16), _ptr_array(Whh, torch.bfloat16)
    a_bi, a_bh = _ptr_array(b_ih, torch.float32), _ptr_array(b_hh, torch.float32)
    _C.lib().call("ark_gru_wave_fwd", _ptr(x_b, torch.bfloat16), _ptr(hp_b, torch.bfloat16), _ptr(out_b, torch.bfloat16),
                  _ptr(h0, torch.float32), a_ih, a_hh, a_bi, a_bh, _ptr(bt_dev, torch.int32), _ptr(off_dev, torch.int32),
                  L, bt0, N, d, nl, _ptr(r, torch.bfloat16), _ptr(z, torch.bfloat16), _ptr(n, torch.bfloat16),
                  _ptr(ghn, torch.bfloat16), _ptr(mask, torch.uint8), float(p_drop), int(seed), int(offset),
                  _ptr(offset_dev, torch.int64), _ptr(sync_ws, torch.int32), _stream())


def gru_wave_bwd(dy_top, gates, hp_b, mask, p_drop, WhhT, WihT, bt_dev, off_dev, L, bt0, d, dgi_b, dgh_b, dh0, sync_ws):
    nl, N = hp_b.shape[0], hp_b.shape[1]
    r, z, n, ghn = gates
    _contig(dy_top, r, z, n, ghn, hp_b, mask, dgi_b, dgh_b, dh0, *WhhT, *[w for w in WihT if w is not None])
    a_hh, a_ih = _ptr_array(WhhT, torch.bfloat16), _ptr_array(WihT, torch.bfloat16)
    _C.lib().call("ark_gru_wave_bwd", _ptr(dy_top, torch.float32), _ptr(r, torch.bfloat16), _ptr(z, torch.bfloat16),
                  _ptr(n, torch.bfloat16), _ptr(ghn, torch.bfloat16), _ptr(hp_b, torch.bfloat16), _ptr(mask, torch.uint8),
                  float(p_drop), a_hh, a_ih, _ptr(bt_dev, torch.int32), _ptr(off_dev, torch.int32), L, bt0, N, d, nl,
                  _ptr(dgi_b, torch.bfloat16), _ptr(dgh_b, torch.bfloat16), _ptr(dh0, torch.float32),
                  _ptr(sync_ws, torch.int32), _stream())


def gru_cluster_supported(d, bt0, nl, L) -> int:
    """Batch-tile rows (16/32/64) of the cluster GRU stack kernel (csrc/gru_cluster.cu) or 0."""
    return int(_C.lib().raw("ark_gru_cluster_supported")(int(d), int(bt0), int(nl), int(L)))


def gru_cluster_workspace_bytes(L, bt0, d, nl) -> int:
    return int(_C.lib().raw("ark_gru_cluster_workspace_bytes")(int(L), int(bt0), int(d), int(nl)))


def gru_cluster_fwd(x_b, hp_b, out_b, h0, Wih, Whh, b_ih, b_hh, bt_dev, off_dev, L, bt0, d, gates, mask, p_drop, seed,
                    offset, offset_dev, sync_ws, ws):
    """gru_wave_fwd's contract through thread-block clusters + distributed shared memory (csrc/gru_cluster.cu)."""
    nl, N = hp_b.shape[0], hp_b.shape[1]
    r, z, n, ghn = gates if gates is not None else (None, None, None, None)
    _contig(x_b, hp_b, out_b, h0, r, z, n, ghn, mask, ws, *Wih, *Whh)
    a_ih, a_hh = _ptr_array(Wih, torch.bfloat16), _ptr_array(Whh, torch.bfloat16)
    a_bi, a_bh = _ptr_array(b_ih, torch.float32), _ptr_array(b_hh, torch.float32)
    _C.lib().call("ark_gru_cluster_fwd", _ptr(x_b, torch.bfloat16), _ptr(hp_b, torch.bfloat16), _ptr(out_b, torch.bfloat16),
                  _ptr(h0, torch.float32), a_ih, a_hh, a_bi, a_bh, _ptr(bt_dev, torch.int32), _ptr(off_dev, torch.int32),
                  L, bt0, N, d, nl, _ptr(r, torch.bfloat16), _ptr(z, torch.bfloat16), _ptr(n, torch.bfloat16),
                  _ptr(ghn, torch.bfloat16), _ptr(mask, torch.uint8), float(p_drop), int(seed), int(offset),
                  _ptr(offset_dev, torch.int64), _ptr(sync_ws, torch.int32), _ptr(ws, torch.uint8), ws.numel(), _stream())


def gru_cluster_bwd(dy_top, gates, hp_b, mask, p_drop, WhhT, WihT, bt_dev, off_dev, L, bt0, d, dgi_b, dgh_b, dh0, sync_ws,
                    ws):
    nl, N = hp_b.shape[0], hp_b.shape[1]
    r, z, n, ghn = gates
    _contig(dy_top, r, z, n, ghn, hp_b, mask, dgi_b, dgh_b, dh0, ws, *WhhT, *[w for w in WihT if w is not None])
    a_hh, a_ih = _ptr_array(WhhT, torch.bfloat16), _ptr_array(WihT, torch.bfloat16)
    _C.lib().call("ark_gru_cluster_bwd", _ptr(dy_top, torch.float32), _ptr(r, torch.bfloat16), _ptr(z, torch.bfloat16),
                  _ptr(n, torch.bfloat16), _ptr(ghn, torch.bfloat16), _ptr(hp_b, torch.bfloat16), _ptr(mask, torch.uint8),
                  float(p_drop), a_hh, a_ih, _ptr(bt_dev, torch.int32), _ptr(off_dev, torch.int32), L, bt0, N, d, nl,
                  _ptr(dgi_b, torch.bfloat16), _ptr(dgh_b, torch.bfloat16), _ptr(dh0, torch.float32),
                  _ptr(sync_ws, torch.int32), _ptr(ws, torch.uint8), ws.numel(), _stream())


# ------------------------------------------------------------------ t-SAIL (Transformer) blocks: csrc/attn_ops.cu
TOK, SQ = 0, 1


def _opnd(t, kind):
    """(ptr, is_f32, ld) of a bgemm operand; SQ buffers are flat."""
    if t.dtype not in (torch.float32, torch.bfloat16):
        raise _C.ArkError("attention operands must be f32 or bf16")
    return _ptr(t), int(t.dtype == torch.float32), (t.stride(0) if kind == TOK else 0)


def attn_bgemm(A, a_kind, a_trans, a_col0, B, b_kind, b_trans, b_col0, C, c_kind, c_col0, seg, H, hd, mode, causal, alpha):
    """C_p = alpha * A_p . B_p for every (graph, head); `seg` is a layout.Segments (cu, sq_off, n_max)."""
    ap, af, ald = _opnd(A, a_kind)
    bp, bfl, bld = _opnd(B, b_kind)
    cp, cf, cld = _opnd(C, c_kind)
    _C.lib().call("ark_attn_bgemm", ap, a_kind, int(a_trans), af, ald, a_col0, bp, b_kind, int(b_trans), bfl, bld, b_col0,
                  cp, c_kind, cf, cld, c_col0, _ptr(seg.cu_dev, torch.int32), _ptr(seg.sq_dev, torch.int64), seg.n_graphs,
                  seg.n_max, H, hd, mode, int(causal), float(alpha), _stream())


def attn_softmax_fwd(S, seg, H, causal, p_drop, seed, offset, offset_dev, P, P_drop):
    _C.lib().call("ark_attn_softmax_fwd", _ptr(S, torch.float32), _ptr(seg.cu_dev, torch.int32), _ptr(seg.sq_dev, torch.int64),
                  _ptr(seg.graph_dev, torch.int32), seg.n_rows, H, int(causal), float(p_drop), int(seed), int(offset),
                  _ptr(offset_dev, torch.int64), _ptr(P, torch.bfloat16), _ptr(P_drop, torch.bfloat16), _stream())


def attn_softmax_inplace(S, seg, H, causal):
    _C.lib().call("ark_attn_softmax_inplace", _ptr(S, torch.float32), _ptr(seg.cu_dev, torch.int32),
                  _ptr(seg.sq_dev, torch.int64), _ptr(seg.graph_dev, torch.int32), seg.n_rows, H, int(causal), _stream())


def attn_softmax_bwd(P, P_drop, dP, seg, H, causal, p_drop, alpha, dS):
    _C.lib().call("ark_attn_softmax_bwd", _ptr(P, torch.bfloat16), _ptr(P_drop, torch.bfloat16), _ptr(dP, torch.float32),
                  _ptr(seg.cu_dev, torch.int32), _ptr(seg.sq_dev, torch.int64), _ptr(seg.graph_dev, torch.int32),
                  seg.n_rows, H, int(causal), float(p_drop), float(alpha), _ptr(dS, torch.bfloat16), _stream())


def add_layernorm_fwd(branch, res, gamma, beta, eps, p_drop, seed, offset, offset_dev, mask, y, y_bf16, mean, rstd):
    n, D = branch.shape
    _contig(branch, res, gamma, beta, mask, y, y_bf16, mean, rstd)
    _C.lib().call("ark_add_layernorm_fwd", _ptr(branch, torch.float32), _ptr(res, torch.float32), _ptr(gamma, torch.float32),
                  _ptr(beta, torch.float32), n, D, float(eps), float(p_drop), int(seed), int(offset),
                  _ptr(offset_dev, torch.int64), _ptr(mask, torch.uint8), _ptr(y, torch.float32),
                  _ptr(y_bf16, torch.bfloat16), _ptr(mean, torch.float32), _ptr(rstd, torch.float32), _stream())


def add_layernorm_bwd(dy, s, mean, rstd, gamma, p_drop, mask, d_res, d_branch_f32, d_branch_bf16, dgamma, dbeta):
    n, D = dy.shape
    _contig(dy, s, mean, rstd, gamma, mask, d_res, d_branch_f32, d_branch_bf16, dgamma, dbeta)
    _C.lib().call("ark_add_layernorm_bwd", _ptr(dy, torch.float32), _ptr(s, torch.float32), _ptr(mean, torch.float32),
                  _ptr(rstd, torch.float32), _ptr(gamma, torch.float32), n, D, float(p_drop), _ptr(mask, torch.uint8),
                  _ptr(d_res, torch.float32), _ptr(d_branch_f32, torch.float32), _ptr(d_branch_bf16, torch.bfloat16),
                  _ptr(dgamma, torch.float32), _ptr(dbeta, torch.float32), _stream())


def triple_embed_fwd(idx, E, R, X, X_bf16):
    _contig(idx, E, R, X, X_bf16)
    _C.lib().call("ark_triple_embed_fwd", _ptr(idx, torch.int32), _ptr(E, torch.float32), _ptr(R, torch.float32),
                  idx.shape[0], E.shape[1], _ptr(X, torch.float32), _ptr(X_bf16, torch.bfloat16), _stream())


def triple_embed_bwd(idx, dX, dE, dR):
    _contig(idx, dX, dE, dR)
    _C.lib().call("ark_triple_embed_bwd", _ptr(idx, torch.int32), _ptr(dX, torch.float32), idx.shape[0], dE.shape[1],
                  _ptr(dE, torch.float32), _ptr(dR, torch.float32), _stream())


def embed_sum_fwd(W, P, tok, pos, X, X_bf16):
    _contig(W, P, tok, pos, X, X_bf16)
    _C.lib().call("ark_embed_sum_fwd", _ptr(W, torch.float32), _ptr(P, torch.float32), _ptr(tok, torch.int32),
                  _ptr(pos, torch.int32), tok.numel(), W.shape[1], _ptr(X, torch.float32), _ptr(X_bf16, torch.bfloat16),
                  _stream())


def seg_reduce(X, seg, mean, wts, H, out, out_bf16):
    _contig(X, wts, out, out_bf16)
    _C.lib().call("ark_seg_reduce", _ptr(X, torch.float32), _ptr(seg.cu_dev, torch.int32), seg.n_graphs, X.shape[1], int(mean),
                  _ptr(wts, torch.float32), H, _ptr(out, torch.float32), _ptr(out_bf16, torch.bfloat16), _stream())


def seg_broadcast(src, seg, mean, wts, H, out, out_bf16):
    _contig(src, wts, out, out_bf16)
    _C.lib().call("ark_seg_broadcast", _ptr(src, torch.float32), _ptr(seg.cu_dev, torch.int32), _ptr(seg.graph_dev, torch.int32),
                  seg.n_rows, src.shape[1], int(mean), _ptr(wts, torch.float32), H, _ptr(out, torch.float32),
                  _ptr(out_bf16, torch.bfloat16), _stream())


def xattn_weights(n_items, n_keys, p_drop, seed, offset, offset_dev, wts):
    _C.lib().call("ark_xattn_weights", n_items, n_keys, float(p_drop), int(seed), int(offset), _ptr(offset_dev, torch.int64),
                  _ptr(wts, torch.float32), _stream())


def relu_bwd(d, out, scale, dpre):
    _contig(d, out, dpre)
    _C.lib().call("ark_relu_bwd", _ptr(d, torch.float32), _ptr(out, torch.bfloat16), d.numel(), float(scale),
                  _ptr(dpre, torch.bfloat16), _stream())
