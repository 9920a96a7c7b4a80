"""CPU-side tests: C-ABI exports, packing layout, flat parameter storage, batch sharding, and the
data-parallel normaliser logic (2-process gloo) checked with the numpy oracle.  No GPU compute."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from ark_b200 import _C
from ark_b200.flat import FlatParams, sail_param_order
from ark_b200.layout import pack_layout, unpack_rows
from ark_b200.synthetic import model_config, synth_batch, vocab_layout
from oracle import sail_oracle as O


def test_library_loads_and_exports_every_declared_symbol():
    lib = _C.lib()
    header = open(_C.HEADER).read()
    declared = set(re.findall(r"\b(ark_\w+)\s*\(", header))
    assert declared and declared == set(lib.protos)
    dll = ctypes.CDLL(_C.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), name
    assert dll.ark_abi_version() == 1


def test_argument_validation_happens_before_any_launch():
    lib = _C.lib()
    with pytest.raises(_C.ArkError, match="null pointer"):
        lib.call("ark_softmax_ce", None, 1, 4, 10, 16, None, 1.0, 1, None, None, None)
    with pytest.raises(_C.ArkError, match="multiple of 8"):
        lib.call("ark_softmax_ce", ctypes.c_void_p(256), 1, 4, 10, 10, ctypes.c_void_p(256), 1.0, 1, None, None, None)
    with pytest.raises(_C.ArkError, match="TMA needs"):
        lib.call("ark_gemm_bf16_tc", ctypes.c_void_p(256), 0, 10, ctypes.c_void_p(512), 0, 16, ctypes.c_void_p(1024), 0,
                 16, 4, 4, 10, None, 0, 0, None, None)


def test_product_path_refuses_cpu_tensors():
    from kgvae.model.models import SAIL
    cfg = model_config("syn-paths", d_model=16, d_latent=4, n_layers=1)
    m = SAIL(cfg)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.enc(torch.zeros(2, 3, 3, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.engine()


def test_pack_layout_properties():
    lay_cfg = vocab_layout("wd-movies")
    tri, seq, n_tri = synth_batch(lay_cfg, 37, seed=3)
    lay = pack_layout(seq)
    lens = (seq[:, 1:] != 0).sum(1).numpy()
    assert lay.n_tok == lens.sum() and lay.n_triples == n_tri == int(((lens - 1) // 3).sum())
    assert (np.diff(lay.bt) <= 0).all() and lay.bt[0] == 37 and lay.L == lens.max()
    assert (lens[lay.perm][:-1] >= lens[lay.perm][1:]).all()
    assert (lay.off[1:] - lay.off[:-1] == lay.bt).all()
    # pack -> unpack round trip of the target tokens
    rows = []
    for t in range(lay.L):
        rows += [seq[lay.perm[j], t + 1].item() for j in range(lay.bt[t])]
    dense = unpack_rows(torch.tensor(rows), lay, 37, seq.shape[1] - 1, fill=0)
    assert torch.equal(dense, seq[:, 1:])
    # valid triples <-> mask agreement with the reference rule (relation != pad_rid)
    assert int((tri[:, :, 1] != lay_cfg["pad_rid"]).sum()) == n_tri


def test_synthetic_batches_follow_reference_token_layout():
    for ds in ("syn-paths", "wd-articles"):
        lay = vocab_layout(ds)
        tri, seq, _ = synth_batch(lay, 5, seed=1)
        olay = O.vocab_layout(lay["n_entities_raw"], lay["n_relations_raw"], lay["max_edges"], lay["use_padding"])
        for k in ("vocab_size", "seq_len", "ENT_BASE", "REL_BASE", "pad_eid", "pad_rid", "n_entities", "n_relations"):
            assert lay[k] == olay[k]
        for b in range(5):
            g = [tuple(int(x) for x in t) for t in tri[b] if lay["pad_rid"] is None or t[1] != lay["pad_rid"]]
            assert O.triples_to_seq(g, olay).tolist() == seq[b].tolist()


def test_flat_params_alias_module_and_keep_state_dict():
    from kgvae.model.models import SAIL
    cfg = model_config("wd-movies", d_model=16, d_latent=4, n_layers=2)
    cfg.update(vocab_layout("wd-movies"))
    cfg.update(n_entities=50, n_relations=4, vocab_size=57, pad_eid=49, pad_rid=3)
    torch.manual_seed(0)
    m = SAIL(cfg)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat = FlatParams(sail_param_order(m), "cpu")
    after = m.state_dict()
    assert list(before) == list(after)
    for k in before:
        assert torch.equal(before[k], after[k])
    assert m.dec.out.weight is m.dec.tok_emb.weight
    # module parameters alias the flat buffer; mu/logv are adjacent (one fused [2dz, 3d] operand)
    flat.param.add_(1.0)
    assert torch.equal(m.enc.mu.weight.data, before["enc.mu.weight"] + 1)
    fused = flat.fused(flat.param, "enc.mu.weight", "enc.logv.weight", (8, 48))
    assert torch.equal(fused[:4], m.enc.mu.weight.data) and torch.equal(fused[4:], m.enc.logv.weight.data)
    for name, (off, n, shape) in flat.slots.items():
        assert off % 4 == 0
    assert flat.slots["dec.out.bias"][0] == 0        # first gradient to become final


def test_batch_loader_shards_and_global_counts():
    from kgvae.experiments.train import BatchLoader
    rng = np.random.default_rng(0)
    graphs = [[(int(rng.integers(9)), int(rng.integers(2)), int(rng.integers(9))) for _ in range(int(rng.integers(1, 5)))]
              for _ in range(37)]
    v = {"ENT_BASE": 3, "REL_BASE": 13, "seq_len": 14, "max_edges": 4, "use_padding": True, "pad_eid": 9, "pad_rid": 2}
    seen = []
    for r in range(2):
        for tri, seq, ntg, bg in BatchLoader(graphs, v, 4, rank=r, world=2):
            assert tri.shape == (4, 4, 3) and seq.shape == (4, 14) and bg == 8
            seen.append((r, seq, ntg))
    assert len(seen) == 2 * (37 // 8)
    for g in range(37 // 8):
        a, b = seen[g], seen[37 // 8 + g]
        local = int((a[1][:, 1:] != 0).sum() + (b[1][:, 1:] != 0).sum())
        assert a[2] == b[2] == local                    # every rank derives the same GLOBAL token count


_DDP_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import sail_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
rng = np.random.default_rng(0)
lay = O.vocab_layout(30, 3, 5, True)
graphs = [[(int(rng.integers(30)), int(rng.integers(3)), int(rng.integers(30))) for _ in range(int(rng.integers(1, 6)))] for _ in range(8)]
tri, seq = O.build_batch(graphs, lay)
cfg = dict(lay, model_type="SAIL", d_model=8, d_latent=4, n_layers=2)
prng = np.random.default_rng(1)
shapes = {"enc.e_emb.weight": (31, 8), "enc.r_emb.weight": (4, 8), "enc.mlp.0.weight": (24, 24), "enc.mlp.0.bias": (24,),
          "enc.mlp.2.weight": (24, 24), "enc.mlp.2.bias": (24,), "enc.mu.weight": (4, 24), "enc.mu.bias": (4,),
          "enc.logv.weight": (4, 24), "enc.logv.bias": (4,), "dec.tok_emb.weight": (38, 8), "dec.z_proj.weight": (8, 4),
          "dec.z_proj.bias": (8,), "dec.out.bias": (38,)}
for k in range(2):
    shapes.update({f"dec.gru.weight_ih_l{k}": (24, 8), f"dec.gru.weight_hh_l{k}": (24, 8), f"dec.gru.bias_ih_l{k}": (24,), f"dec.gru.bias_hh_l{k}": (24,)})
params = {k: prng.standard_normal(s) * 0.3 for k, s in shapes.items()}
eps = prng.standard_normal((8, 4))
whole_l, whole_g, _ = O.elbo_step(params, cfg, tri, seq, eps, 0.5)
sl = slice(rank * 4, rank * 4 + 4)
n_tok_global = int((seq[:, 1:] != 0).sum())
loc_l, loc_g, _ = O.elbo_step(params, cfg, tri[sl], seq[sl], eps[sl], 0.5, n_tok_global=n_tok_global, batch_global=8)
flat = torch.from_numpy(np.concatenate([loc_g[k].ravel() for k in sorted(loc_g)]))
dist.all_reduce(flat)                                   # ncclSum in production; gradients are SUMMED, never averaged
ref = np.concatenate([whole_g[k].ravel() for k in sorted(whole_g)])
st = torch.tensor([loc_l["ce"], loc_l["kl"]], dtype=torch.float64)
dist.all_reduce(st)
assert np.allclose(flat.numpy(), ref, rtol=1e-9, atol=1e-12), np.abs(flat.numpy() - ref).max()
assert np.allclose(st.numpy(), [whole_l["ce"], whole_l["kl"]], rtol=1e-12)
dist.destroy_process_group()
print("OK", rank)
'''


def test_two_rank_gloo_gradient_sum_equals_whole_batch(tmp_path):
    script = tmp_path / "ddp_worker.py"
    script.write_text(_DDP_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2


def test_merge_span_handles_adjacent_and_out_of_order_slots():
    """Gradient slots arrive in the order the backward pass finishes them, which is NOT the layout order once the
    leaf weight-gradient GEMMs are deferred behind the critical chain: only a slot directly behind the previous one
    may extend it; everything else is its own slice (nothing is dropped, nothing overlaps)."""
    from ark_b200.flat import merge_span
    p = []
    merge_span(p, 0, 100)
    merge_span(p, 128, 300)          # behind the 64-element alignment gap: same slice
    assert p == [(0, 300)]
    merge_span(p, 1000, 1200)        # a hole: new slice
    merge_span(p, 400, 600)          # out of order (earlier in the layout): new slice, never a negative range
    merge_span(p, 640, 700)
    assert p == [(0, 300), (1000, 1200), (400, 700)]
    assert all(e > s for s, e in p)
    covered = sorted(p)
    assert all(covered[i][1] <= covered[i + 1][0] for i in range(len(covered) - 1))


_FACTOR_WORKER = r'''
import sys
import numpy as np
import torch
import torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.Generator().manual_seed(100 + rank)
B, n = 6, 40                                        # B << n: dW = dY^T X has rank <= world * B
X = torch.randn(B, n, generator=g, dtype=torch.float64)
dY = torch.randn(B, n, generator=g, dtype=torch.float64)
# what the engine used to do: all-reduce the [n, n] weight gradient (and the bias gradient)
dW = dY.t() @ X
db = dY.sum(0)
dist.all_reduce(dW)
dist.all_reduce(db)
# what it does now (ark_b200/elbo.py, dp_factor_gather): all-gather the two [B, n] factors, ONE K = world*B product
X_all, dY_all = torch.empty(world * B, n, dtype=torch.float64), torch.empty(world * B, n, dtype=torch.float64)
dist.all_gather_into_tensor(X_all, X)
dist.all_gather_into_tensor(dY_all, dY)
assert torch.allclose(dY_all.t() @ X_all, dW, rtol=1e-12, atol=1e-12)
assert torch.allclose(dY_all.sum(0), db, rtol=1e-12, atol=1e-12)
# every rank computed the SAME bits from the same gathered operands (no drift between replicas)
ref = (dY_all.t() @ X_all).clone()
dist.broadcast(ref, 0)
assert torch.equal(ref, dY_all.t() @ X_all)
assert 2 * world * B * n < n * n                   # and it moves fewer numbers than the all-reduce
dist.destroy_process_group()
print("OK", rank)
'''


def test_two_rank_gloo_factor_gather_equals_weight_gradient_all_reduce(tmp_path):
    script = tmp_path / "factor_worker.py"
    script.write_text(_FACTOR_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29633", str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2


def test_bf16_partial_sum_reduce_scatter_error_is_operand_rounding_sized():
    """Numerical contract of the cluster GRU backward (ark_b200/csrc/gru_cluster.cu): every CTA multiplies its 96 gate
    columns (fp32 accumulate over K = 96 from bf16 operands), ROUNDS its partial sum to bf16 for the exchange through
    distributed shared memory, and the owner adds the CS partials in fp32.  The extra error w.r.t. one fp32
    accumulation over K = 3d stays of the order of the bf16 rounding of the operands themselves (so the 3e-2
    gradient tolerance of the parity tests is not consumed by it)."""
    rng = np.random.default_rng(0)
    d, nb, cs = 512, 16, 16

    def bf16(x):
        x = np.asarray(x, dtype=np.float32)
        u = x.view(np.uint32)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000       # round to nearest even
        return u.view(np.float32)

    W = bf16(rng.standard_normal((3 * d, d)) / np.sqrt(d))     # W_hh [3d, d]
    dgh = bf16(rng.standard_normal((nb, 3 * d)) * 1e-2)        # dgh_t [NB, 3d]
    exact = dgh.astype(np.float64) @ W.astype(np.float64)       # what one long accumulation of the bf16 operands gives
    # CTA c owns gate rows {g*d + 32c .. 32c+31}: its partial, rounded to bf16 before the exchange
    total = np.zeros((nb, d), dtype=np.float32)
    for c in range(cs):
        rows = np.concatenate([g * d + 32 * c + np.arange(32) for g in range(3)])
        part = (dgh[:, rows].astype(np.float64) @ W[rows].astype(np.float64)).astype(np.float32)
        total += bf16(part)
    rel = np.linalg.norm(total - exact) / np.linalg.norm(exact)
    # reference point: the error the bf16 rounding of the OPERANDS already causes w.r.t. fp32 operands
    W32 = rng.standard_normal((3 * d, d)).astype(np.float32) / np.sqrt(d)
    g32 = (rng.standard_normal((nb, 3 * d)) * 1e-2).astype(np.float32)
    op_rel = (np.linalg.norm(bf16(g32).astype(np.float64) @ bf16(W32).astype(np.float64) - g32.astype(np.float64) @ W32)
              / np.linalg.norm(g32.astype(np.float64) @ W32))
    assert rel < 4e-3 and rel < 2.0 * op_rel, (rel, op_rel)


def test_bench_kernel_table_and_roofline_selection():
    """bench.py picks the single kernel with the largest in-graph time as `roofline` (never an aggregate of families),
    computes achieved / peak from the algorithmic work and attaches the ncu DRAM traffic of that kernel."""
    import bench
    pk = {"hbm_gbs": 6552.0, "bf16_tflops": 1641.5, "bf16_tflops_sustained": 1386.4, "src": "measured"}
    agg = {"gru_persist_bwd": {"ms": 0.9, "calls": 9, "flops": 9 * 16.1e9, "bytes": 0.0},
           "gemm_tc:gru_gi": {"ms": 0.27, "calls": 9, "flops": 9 * 16.1e9, "bytes": 0.0},
           "softmax_ce": {"ms": 0.05, "calls": 3, "flops": 0.0, "bytes": 3 * 1.2e6},
           "nccl_all_reduce": {"ms": 1.2, "calls": 15, "flops": 0.0, "bytes": 15 * 4e7}}
    agg["adam_flat"] = {"ms": 1.0, "calls": 24, "flops": 0.0, "bytes": 24 * 1.8e8}
    rows = bench.kernel_table(agg, 3, 1.25, pk, "syn-types", {"adam_flat": "side", "nccl_all_reduce": "side"})
    assert [r["name"] for r in rows] == ["nccl_all_reduce", "adam_flat", "gru_persist_bwd", "gemm_tc:gru_gi", "softmax_ce"]
    roof = bench.roofline_from(rows, pk, "graph-replay event nodes")
    assert roof["kernel"] == "gru_persist_bwd" and roof["bound"] == "tensor"   # side-stream Adam / NCCL overlap the chain
    assert abs(roof["achieved"] - 16.1e9 / (0.1e-3) / 1e12) < 1e-6 and abs(roof["frac"] - roof["achieved"] / 1386.4) < 1e-9
    assert roof["launches_per_step"] == 3.0 and abs(roof["share_of_step"] - 0.3 / 1.25) < 1e-9
    ce = [r for r in rows if r["name"] == "softmax_ce"][0]
    assert ce["bound"] == "hbm" and abs(ce["achieved"] - 1.2e6 / (0.05e-3 / 3) / 1e9) < 1e-3
    if ("syn-types", "gru_persist_bwd") in bench.NCU_TRAFFIC:
        assert roof["traffic"] == bench.NCU_TRAFFIC[("syn-types", "gru_persist_bwd")][0]


def test_multicast_exchange_slices_partition_every_span():
    """ark_b200.symm.SymmFlat.owned mirrors the slice rule of csrc/dp_reduce.cu (rank r owns the r-th ceil(n4/world) vec4 block
    of a span): the slices are disjoint, cover the span, stay 16-byte aligned — for ragged sizes and more ranks than vec4s."""
    from ark_b200.symm import SymmFlat
    for world in (2, 3, 4, 8, 16):
        sym = SymmFlat.__new__(SymmFlat)
        sym.world = world
        for s, e in ((0, 4), (64, 64 + 4 * 7), (128, 128 + 4 * 1000003), (4096, 4096)):
            sl = sym.owned(s, e)
            assert len(sl) == world and sl[0][0] == s and sl[-1][1] == e
            for (a, b), (c, _) in zip(sl, sl[1:] + [(e, e)]):
                assert a <= b == c and a % 4 == 0 and b % 4 == 0
            per = ((e - s) // 4 + world - 1) // world
            assert all(b - a <= 4 * per for a, b in sl)


_SHARDED_ADAM_WORKER = r'''
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from ark_b200.symm import SymmFlat
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sym = SymmFlat.__new__(SymmFlat)
sym.world, sym.rank = world, rank
n, spans = 4 * 257, [(0, 4 * 100), (4 * 128, 4 * 257)]       # ragged vec4 counts; a gap no span covers
g0 = torch.Generator().manual_seed(1)
p = torch.randn(n, generator=g0, dtype=torch.float64)
m, v = torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
p_ref, m_ref, v_ref = p.clone(), m.clone(), v.clone()
lr, b1, b2, eps = 1e-2, 0.9, 0.999, 1e-8
def adam(p, g, m, v, t):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_((lr / (1 - b1 ** t)) * m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + eps))
gr = torch.Generator().manual_seed(10 + rank)
for t in (1, 2, 3):
    g = torch.randn(n, generator=gr, dtype=torch.float64)
    # reference: all-reduce + the full update on every rank
    gs = g.clone()
    dist.all_reduce(gs)
    for s, e in spans:
        adam(p_ref[s:e], gs[s:e], m_ref[s:e], v_ref[s:e], t)
    # the exchange of csrc/dp_reduce.cu, spelled with gloo: the owner reduces its slice, updates it, broadcasts p
    for s, e in spans:
        for r, (lo, hi) in enumerate(sym.owned(s, e)):
            part = g[lo:hi].clone()
            dist.reduce(part, dst=r)                       # multimem.ld_reduce by the owner
            if r == rank:
                adam(p[lo:hi], part, m[lo:hi], v[lo:hi], t)
            dist.broadcast(p[lo:hi], src=r)                # multimem.st to every rank
assert torch.allclose(p, p_ref, rtol=1e-12, atol=1e-14)
own = torch.zeros(n, dtype=torch.bool)
for s, e in spans:
    lo, hi = sym.owned(s, e)[rank]
    own[lo:hi] = True
assert torch.allclose(m[own], m_ref[own]) and torch.allclose(v[own], v_ref[own])
cov = torch.zeros(n, dtype=torch.bool)
for s, e in spans:
    cov[s:e] = True
assert not torch.allclose(m[cov & ~own], m_ref[cov & ~own])      # the state IS sharded until it is gathered ...
for s, e in spans:                                                # ... SailEngine.gather_adam_state()
    for r, (lo, hi) in enumerate(sym.owned(s, e)):
        dist.broadcast(m[lo:hi], src=r)
        dist.broadcast(v[lo:hi], src=r)
assert torch.allclose(m, m_ref, rtol=1e-12, atol=1e-14) and torch.allclose(v, v_ref, rtol=1e-12, atol=1e-16)
dist.destroy_process_group()
print("OK", rank)
'''


def test_two_rank_gloo_sharded_adam_exchange_equals_allreduce_plus_adam(tmp_path):
    """The data-parallel protocol of csrc/dp_reduce.cu (owner reduces its 1/world slice, applies Adam there, broadcasts
    the parameters; Adam state sharded until gather_adam_state) restated with gloo collectives == all-reduce + full Adam."""
    script = tmp_path / "sharded_adam_worker.py"
    script.write_text(_SHARDED_ADAM_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29633", str(script), ROOT],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2
