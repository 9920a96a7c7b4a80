"""Bit-exact integer parity of the drop-in kgvae.model.utils with the reference (fixtures in tests/golden)."""
import json
import os

import torch

from conftest import GOLDEN
from kgvae.model import utils as U


def _recs():
    with open(os.path.join(GOLDEN, "utils_indexing.json")) as f:
        return json.load(f)


def test_triples_seq_roundtrip_matches_reference():
    for r in _recs():
        lay = r["layout"]
        sp = lay["special_tokens"]
        for g, s, back in zip(r["graphs"], r["seqs"], r["seq_to_triples"]):
            got = U.triples_to_seq([tuple(t) for t in g], sp, lay["ENT_BASE"], lay["REL_BASE"], lay["seq_len"])
            assert got.dtype == torch.long and got.tolist() == s
            assert [list(t) for t in U.seq_to_triples(got, sp, lay["ENT_BASE"], lay["REL_BASE"])] == back
        for s, back in zip(r["odd"], r["odd_back"]):
            assert [list(t) for t in U.seq_to_triples(s, sp, lay["ENT_BASE"], lay["REL_BASE"])] == back


def test_dataset_and_vectorised_batch_match_reference():
    for r in _recs():
        lay = r["layout"]
        graphs = [[tuple(t) for t in g] for g in r["graphs"]]
        kw = dict(special_tokens=lay["special_tokens"], ent_base=lay["ENT_BASE"], rel_base=lay["REL_BASE"],
                  seq_len=lay["seq_len"], use_padding=r["use_padding"], pad_eid=lay["pad_eid"], pad_rid=lay["pad_rid"])
        ds = U.GraphSeqDataset(graphs, None, None, max_triples=lay["max_edges"], **kw)
        if r["use_padding"] or len({len(g) for g in graphs}) == 1:
            tri, seq = next(iter(torch.utils.data.DataLoader(ds, batch_size=len(graphs))))
            assert tri.tolist() == r["batch_triples"] and seq.tolist() == r["batch_seq"]
            tri2, seq2 = U.build_batch(graphs, max_triples=lay["max_edges"], **kw)
            assert torch.equal(tri2, tri) and torch.equal(seq2, seq)


def test_build_batch_edge_cases():
    sp = {"PAD": 0, "BOS": 1, "EOS": 2}
    # empty graph with padding: only BOS, EOS
    tri, seq = U.build_batch([[], [(0, 0, 1)]], special_tokens=sp, ent_base=3, rel_base=6, seq_len=8, max_triples=2,
                             use_padding=True, pad_eid=2, pad_rid=1)
    assert seq.tolist() == [[1, 2, 0, 0, 0, 0, 0, 0], [1, 3, 6, 4, 2, 0, 0, 0]]
    assert tri.tolist() == [[[2, 1, 2], [2, 1, 2]], [[0, 0, 1], [2, 1, 2]]]
