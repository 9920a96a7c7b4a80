"""Synthetic IntelliGraphs-shaped triple batches (no network in this environment, so no dataset download).

Shape constants (entities / relations / triples per graph) are the IntelliGraphs figures recalled in
SURVEY.md §8(d); they parameterise the generator and are not taken from the reference repository.  Everything
derived from them — vocabulary layout, sequence length, padding — follows the reference's rules
(kgvae/experiments/ablation_study.py:436-454, kgvae/model/utils.py:102-146).
"""
from __future__ import annotations

import numpy as np
import torch

from .layout import pack_layout

# dataset -> (n_entities, n_relations, min_edges, max_edges, use_padding)
DATASET_SHAPES = {
    "syn-paths": (49, 3, 3, 3, False),
    "syn-types": (30, 3, 3, 3, False),
    "syn-tipr": (130, 5, 5, 5, False),
    "wd-movies": (24093, 3, 2, 23, True),
    "wd-articles": (60932, 6, 4, 212, True),
}
# dataset -> YAML model hyper-parameters (configs/autoreg_<ds>.yaml)
MODEL_SHAPES = {
    "syn-paths": dict(d_model=512, d_latent=10, n_heads=4, n_layers=3, batch_size=256),
    "syn-types": dict(d_model=1024, d_latent=24, n_heads=4, n_layers=3, batch_size=256),
    "syn-tipr": dict(d_model=1024, d_latent=32, n_heads=16, n_layers=3, batch_size=256),
    "wd-movies": dict(d_model=128, d_latent=64, n_heads=8, n_layers=3, batch_size=256),
    "wd-articles": dict(d_model=512, d_latent=128, n_heads=8, n_layers=3, batch_size=16),
}


def vocab_layout(dataset, use_padding=None):
    """ablation_study.py:436-454 applied to the synthetic shape of `dataset`."""
    nE, nR, lo, hi, pad = DATASET_SHAPES[dataset]
    if use_padding is not None:
        pad = bool(use_padding)
    pad_eid = pad_rid = None
    if pad:
        pad_eid, pad_rid = nE, nR
        nE, nR = nE + 1, nR + 1
    ent_base = 3
    rel_base = ent_base + nE
    return {"n_entities": nE, "n_relations": nR, "pad_eid": pad_eid, "pad_rid": pad_rid,
            "special_tokens": {"PAD": 0, "BOS": 1, "EOS": 2}, "ENT_BASE": ent_base, "REL_BASE": rel_base,
            "vocab_size": rel_base + nR, "seq_len": 3 * hi + 2, "max_edges": hi, "min_edges": lo,
            "use_padding": pad, "n_entities_raw": DATASET_SHAPES[dataset][0], "n_relations_raw": DATASET_SHAPES[dataset][1]}


def model_config(dataset, **overrides):
    cfg = dict(vocab_layout(dataset), model_type="SAIL", dataset=dataset, **MODEL_SHAPES[dataset])
    cfg.update(overrides)
    return cfg


def synth_batch(lay, batch, seed, dense=False):
    """One batch as HOST tensors in the reference's format: triples int64 [B,T,3], seq int64 [B,seq_len].

    torch.Generator(seed); n_b ~ U[min,max] (max if `dense`); h,t ~ U[0,nE_raw), r ~ U[0,nR_raw) — real ids
    only; padding triples (pad_eid,pad_rid,pad_eid) and PAD tokens exactly as GraphSeqDataset would emit.
    """
    g = torch.Generator().manual_seed(int(seed))
    lo, hi = lay["min_edges"], lay["max_edges"]
    nE, nR = lay["n_entities_raw"], lay["n_relations_raw"]
    n = torch.full((batch,), hi, dtype=torch.int64) if (dense or lo == hi) else \
        torch.randint(lo, hi + 1, (batch,), generator=g)
    h = torch.randint(0, nE, (batch, hi), generator=g)
    r = torch.randint(0, nR, (batch, hi), generator=g)
    t = torch.randint(0, nE, (batch, hi), generator=g)
    live = torch.arange(hi)[None, :] < n[:, None]
    T = hi
    tri = torch.stack([h, r, t], -1)
    if lay["use_padding"]:
        pad = torch.tensor([lay["pad_eid"], lay["pad_rid"], lay["pad_eid"]])
        tri = torch.where(live[..., None], tri, pad.expand(batch, T, 3))
    seq = torch.zeros(batch, lay["seq_len"], dtype=torch.int64)
    seq[:, 0] = 1
    toks = torch.stack([h + lay["ENT_BASE"], r + lay["REL_BASE"], t + lay["ENT_BASE"]], -1)
    toks = torch.where(live[..., None], toks, torch.zeros_like(toks)).reshape(batch, 3 * T)
    seq[:, 1:1 + 3 * T] = toks
    seq[torch.arange(batch), 1 + 3 * n] = 2
    return tri.contiguous(), seq.contiguous(), int(n.sum())


class DeviceBatch:
    """A batch resident in HBM (what `bench.py`'s kernel-only `value` is measured on)."""

    def __init__(self, tri, seq, n_triples, device, seed):
        self.n_triples = n_triples
        self.layout = pack_layout(seq).to(device)
        self.triples = tri.to(device)
        self.seq = seq.to(device)
        g = torch.Generator().manual_seed(int(seed) + 7)
        self.host = (tri.pin_memory() if device != "cpu" else tri, seq.pin_memory() if device != "cpu" else seq)
        self.eps_seed = int(seed)

    def eps(self, dz, device):
        g = torch.Generator(device=device).manual_seed(self.eps_seed)
        return torch.randn(self.triples.shape[0], dz, device=device, generator=g)
