#!/usr/bin/env python
"""Data-parallel exactness on real GPUs (run under torchrun, world >= 2):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_check.py

Checks, with injected eps and dec_dropout 0:
  1. eager DP step (bucketed NCCL all-reduce overlapped with backward): summed gradients == the gradient of the
     single-process step on the CONCATENATED batch (computed on rank 0 with the same engine, world=1);
  2. graphed DP step (graph segments + eager NCCL between them) == eager DP step, for three consecutive
     optimiser steps (parameters compared in aggregate: Adam's first steps are sign-like).
Prints one 'DP_CHECK ok' line per rank and exits non-zero on failure.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ark_b200.layout import pack_layout  # noqa: E402
from ark_b200.synthetic import model_config, synth_batch  # noqa: E402
from kgvae.model.models import SAIL  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = model_config("wd-movies", d_model=64, d_latent=16, dec_dropout=0.0)
    B = 24
    tri_all, seq_all, _ = synth_batch(cfg, B * world, 99)
    g = torch.Generator().manual_seed(5)
    eps_all = torch.randn(B * world, cfg["d_latent"], generator=g)
    sl = slice(rank * B, (rank + 1) * B)
    n_tok_g = int((seq_all[:, 1:] != 0).sum())
    beta = 0.5

    def fresh(group):
        torch.manual_seed(0)
        m = SAIL(dict(cfg)).to(dev)
        e = m.engine(lr=2e-3, dist_group=group, bucket_mb=0.25)           # small buckets: several all-reduces
        if "DP_EMB_MIN_BYTES" in os.environ:      # 0: force the (opt-in) embedding factor gather
            e.dp_emb_min_bytes = int(os.environ["DP_EMB_MIN_BYTES"])
        e.dp_factor_gather = os.environ.get("DP_FACTOR_GATHER", "1") != "0"
        return m, e

    tri, seq, eps = tri_all[sl].contiguous().to(dev), seq_all[sl].contiguous(), eps_all[sl].contiguous().to(dev)
    lay = pack_layout(seq).to(dev)
    seq = seq.to(dev)

    # ---- 1. eager DP gradients vs the whole batch on one rank
    m_dp, e_dp = fresh(dist.group.WORLD)
    out = e_dp.forward_backward(tri, seq, lay, eps, beta, n_tok_global=n_tok_g, batch_global=B * world)
    e_dp._sync_grads()
    torch.cuda.synchronize()
    g_dp = e_dp.flat.grad.clone()
    st = out.clone()
    dist.all_reduce(st)
    m_1, e_1 = fresh(None)
    lay_all = pack_layout(seq_all).to(dev)
    out1 = e_1.forward_backward(tri_all.to(dev), seq_all.to(dev), lay_all, eps_all.to(dev), beta)
    rel = ((g_dp - e_1.flat.grad).norm() / e_1.flat.grad.norm()).item()
    assert rel < 2e-2, f"DP gradient != whole-batch gradient: rel {rel}"
    torch.testing.assert_close(st, out1, rtol=2e-3, atol=1e-5)

    # ---- 2. graphed DP == eager DP over three optimiser steps
    res = []
    for graphed in (False, True):
        m, e = fresh(dist.group.WORLD)
        outs = []
        for s in range(3):
            fn = e.train_step_graphed if graphed else e.train_step
            outs.append(fn(tri, seq, lay, eps, beta, n_tok_global=n_tok_g, batch_global=B * world).clone())
        torch.cuda.synchronize()
        res.append((torch.stack(outs), e.flat.param.clone()))
        if e.mm is not None:        # multicast exchange: the Adam state is sharded until the owners broadcast it
            e.gather_adam_state()
            for buf in (e.flat.exp_avg, e.flat.exp_avg_sq):
                ref = buf.clone()
                dist.broadcast(ref, 0)
                assert torch.equal(ref, buf), "Adam state differs between ranks after gather_adam_state()"
            assert e.flat.exp_avg.abs().sum() > 0
        if graphed:
            nseg = len(next(iter(e._graphs.values()))["segs"])
            if e.capture_nccl:
                assert nseg == 1, f"NCCL is captured into the step graph: expected ONE graph, got {nseg} segments"
            else:
                assert nseg >= 3, f"expected several graph segments, got {nseg}"
            e.release_graphs()      # captured NCCL kernels must be gone before the process group is destroyed
    torch.testing.assert_close(res[0][0], res[1][0], rtol=5e-3, atol=1e-5)
    bad = ((res[0][1] - res[1][1]).abs() > 0.25 * 2e-3).float().mean().item()
    assert bad < 0.02, bad
    # every rank holds the same parameters after DP steps
    p = res[1][1].clone()
    dist.broadcast(p, 0)
    if "DP_EMB_MIN_BYTES" not in os.environ:     # (the embedding gather's atomics are not bitwise reproducible)
        assert torch.equal(p, res[1][1]), "ranks diverged"
    print(f"DP_CHECK ok rank {rank}/{world} ({'multicast' if e.mm is not None else 'NCCL'} exchange): grad rel {rel:.2e}, segments {nseg}, graphed-vs-eager mismatches {bad:.4f}",
          flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
