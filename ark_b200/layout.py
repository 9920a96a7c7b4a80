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
        self.perm_dev = torch.from_numpy(self.perm).to(device, non_blocking=True)
        self.bt_dev = torch.from_numpy(self.bt).to(device, non_blocking=True)
        self.off_dev = torch.from_numpy(self.off[:-1].copy()).to(device, non_blocking=True)
        return self


def pack_layout(seq_cpu: torch.Tensor) -> PackedLayout:
    """seq_cpu: int64 [B, seq_len] HOST tensor in the reference token layout (utils.py:102-108)."""
    if seq_cpu.is_cuda:
        raise ValueError("pack_layout works on the host copy of the batch (lengths are host metadata)")
    s = seq_cpu.numpy()
    B, seq_len = s.shape
    lens = (s[:, 1:] != PAD).sum(1).astype(np.int32)          # targets that are not PAD
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
