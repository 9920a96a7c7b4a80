"""Host-side batch layout: PAD-skipping, length-sorted, time-major packing of decoder positions.

The reference pads every graph to ``max_edges`` and runs the decoder over all ``seq_len-1`` positions
(kgvae/model/utils.py:131-146, kgvae/model/models.py:136-142) although PAD targets are ignored by the loss
(ablation_study.py:65-69).  Because the decoder is a forward GRU, dropping those positions is exact for
the loss and every gradient (SURVEY.md finding 7).  Graph b has ``len_b = 3*n_b + 1`` live decoder
positions (BOS + 3 tokens per triple as inputs; the last target is EOS).

Rows are ordered time-major over graphs sorted by decreasing length: row(t, j) = off[t] + j where j is
the rank of the graph in the sorted order and bt[t] = #{b : len_b > t}.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

PAD = 0


@dataclass
class PackedLayout:
    perm: np.ndarray        # int32 [B]   sorted rank -> original graph index
    lens: np.ndarray        # int32 [B]   live decoder positions per ORIGINAL graph
    bt: np.ndarray          # int32 [L]   active graphs per step (non-increasing)
    off: np.ndarray         # int32 [L+1] row offset of step t; off[L] = n_tok
    n_tok: int
    n_triples: int          # real (non-PAD) triples in the batch
    L: int                  # number of steps with at least one active graph
    perm_dev: torch.Tensor = None
    bt_dev: torch.Tensor = None
    off_dev: torch.Tensor = None

    def to(self, device):
        device = torch.device(device)
        if self.perm_dev is not None and self.perm_dev.device == device:
            return self                 # (the cached uniform layouts of fixed-length datasets: nothing to copy again)
        self.perm_dev = torch.from_numpy(self.perm).to(device, non_blocking=True)
        self.bt_dev = torch.from_numpy(self.bt).to(device, non_blocking=True)
        self.off_dev = torch.from_numpy(self.off[:-1].copy()).to(device, non_blocking=True)
        return self


# Fixed-length datasets (syn-*: every graph has the same number of triples): the layout depends on (B, len) only —
# identity permutation, B active graphs at every step.  Deriving it again for every batch (argsort, per-step counts,
# three small H2D copies) is ~70 us of exposed host time in front of a ~1.5 ms step, so it is built once.  The device
# tensors of a cached layout are read-only for every consumer (the graph replay copies perm_dev OUT of it).
_UNIFORM = {}


def _uniform_layout(B: int, n: int) -> PackedLayout:
    lay = _UNIFORM.get((B, n))
    if lay is None:
        if len(_UNIFORM) > 64:
            _UNIFORM.clear()
        lay = PackedLayout(perm=np.arange(B, dtype=np.int32), lens=np.full(B, n, dtype=np.int32),
                           bt=np.full(n, B, dtype=np.int32), off=(np.arange(n + 1, dtype=np.int64) * B).astype(np.int32),
                           n_tok=B * n, n_triples=B * ((n - 1) // 3), L=n)
        _UNIFORM[(B, n)] = lay
    return lay


def pack_layout(seq_cpu: torch.Tensor) -> PackedLayout:
    """seq_cpu: int64 [B, seq_len] HOST tensor in the reference token layout (utils.py:102-108)."""
    if seq_cpu.is_cuda:
        raise ValueError("pack_layout works on the host copy of the batch (lengths are host metadata)")
    s = seq_cpu.numpy()
    return pack_layout_from_lens((s[:, 1:] != PAD).sum(1).astype(np.int32))          # targets that are not PAD


def pack_layout_from_lens(lens: np.ndarray) -> PackedLayout:
    """Layout from the live decoder positions per graph (3 n_b + 1) alone — what a loader that keeps the tensorised
    split on the GPU knows on the host without reading any token back."""
    lens = np.asarray(lens, dtype=np.int32)
    B = len(lens)
    if B and lens[0] > 0 and (lens == lens[0]).all():
        return _uniform_layout(B, int(lens[0]))
    perm = np.argsort(-lens, kind="stable").astype(np.int32)
    L = int(lens.max()) if B else 0
    steps = np.arange(L, dtype=np.int32)
    bt = (lens[None, :] > steps[:, None]).sum(1).astype(np.int32)
    off = np.zeros(L + 1, dtype=np.int32)
    np.cumsum(bt, out=off[1:])
    n_tok = int(off[-1])
    return PackedLayout(perm=perm, lens=lens, bt=bt, off=off, n_tok=n_tok,
                        n_triples=int(((lens - 1) // 3).sum()), L=L)


def unpack_rows(packed: torch.Tensor, lay: PackedLayout, B: int, L_full: int, fill=0.0) -> torch.Tensor:
    """[n_tok, ...] packed rows -> dense [B, L_full, ...] in ORIGINAL graph order (test/inference helper)."""
    out = packed.new_full((B, L_full) + tuple(packed.shape[1:]), fill)
    perm = torch.from_numpy(lay.perm.astype(np.int64)).to(packed.device)
    for t in range(lay.L):
        n, o = int(lay.bt[t]), int(lay.off[t])
        out[perm[:n], t] = packed[o:o + n]
    return out


# ------------------------------------------------------------------------------------------------------------
# t-SAIL: graph-major ragged rows (a graph's rows are contiguous: attention needs all of them together)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class Segments:
    lens: np.ndarray        # int32 [B] rows per graph
    cu: np.ndarray          # int32 [B+1] row offsets
    sq_off: np.ndarray      # int64 [B+1] prefix sums of lens^2 (offset of the graph's score blocks / H)
    graph: np.ndarray       # int32 [n_rows] graph of every row
    n_rows: int
    n_graphs: int
    n_max: int
    sq_total: int
    cu_dev: torch.Tensor = None
    sq_dev: torch.Tensor = None
    graph_dev: torch.Tensor = None

    def to(self, device):
        self.cu_dev = torch.from_numpy(self.cu).to(device, non_blocking=True)
        self.sq_dev = torch.from_numpy(self.sq_off).to(device, non_blocking=True)
        self.graph_dev = torch.from_numpy(self.graph).to(device, non_blocking=True)
        return self


def segments_from_lens(lens) -> Segments:
    lens = np.asarray(lens, dtype=np.int32)
    cu = np.zeros(len(lens) + 1, dtype=np.int32)
    np.cumsum(lens, out=cu[1:])
    sq = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens.astype(np.int64) ** 2, out=sq[1:])
    graph = np.repeat(np.arange(len(lens), dtype=np.int32), lens)
    return Segments(lens=lens, cu=cu, sq_off=sq, graph=graph, n_rows=int(cu[-1]), n_graphs=len(lens),
                    n_max=int(lens.max()) if len(lens) else 0, sq_total=int(sq[-1]))


@dataclass
class TLayout:
    """PAD-free batch of the Transformer KG-VAE: encoder rows = real triples, decoder rows = real positions.
    Exact for loss and gradients: PAD triples are masked keys whose own outputs the masked mean-pool drops
    (models.py:85-89); PAD decoder positions sit behind the causal mask of every real position and are ignored
    by the loss (models.py:112, ablation_study.py:65-69)."""
    enc: Segments
    dec: Segments
    idx: np.ndarray         # int32 [N_e, 3] (h, r, t) of the real triples
    tok: np.ndarray         # int32 [N]  decoder input token
    pos: np.ndarray         # int32 [N]  position
    tgt: np.ndarray         # int32 [N]  target token (never PAD)
    n_tok: int
    n_triples: int
    L_pad: int              # seq_len - 1: the reference's (padded) memory length
    idx_dev: torch.Tensor = None
    tok_dev: torch.Tensor = None
    pos_dev: torch.Tensor = None
    tgt_dev: torch.Tensor = None

    def to(self, device):
        self.enc.to(device)
        self.dec.to(device)
        for k in ("idx", "tok", "pos", "tgt"):
            setattr(self, k + "_dev", torch.from_numpy(getattr(self, k)).to(device, non_blocking=True))
        return self


def pack_tlayout(triples_cpu: torch.Tensor, seq_cpu: torch.Tensor, pad_rid) -> TLayout:
    tri, s = triples_cpu.numpy(), seq_cpu.numpy()
    B, T, _ = tri.shape
    live = (tri[:, :, 1] != pad_rid) if pad_rid is not None else np.ones((B, T), dtype=bool)
    idx = np.ascontiguousarray(tri[live].astype(np.int32))
    enc = segments_from_lens(live.sum(1))
    L = s.shape[1] - 1
    lens = (s[:, 1:] != PAD).sum(1).astype(np.int32)
    m = np.arange(L)[None, :] < lens[:, None]
    tok = np.ascontiguousarray(s[:, :-1][m].astype(np.int32))
    tgt = np.ascontiguousarray(s[:, 1:][m].astype(np.int32))
    pos = np.ascontiguousarray(np.broadcast_to(np.arange(L, dtype=np.int32)[None, :], (B, L))[m])
    dec = segments_from_lens(lens)
    return TLayout(enc=enc, dec=dec, idx=idx, tok=tok, pos=pos, tgt=tgt, n_tok=dec.n_rows, n_triples=enc.n_rows, L_pad=L)
