#!/usr/bin/env python
"""K11 (csrc/dp_reduce.cu) on real GPUs, under torchrun (world >= 2):
  1. the symmetric allocation rendezvous + multicast mapping (prints what torch's plumbing returned);
  2. mode 0: switch-reduced gradient == NCCL all-reduce;
  3. mode 1: parameters / bf16 shadow after fused reduce + sharded Adam + multicast == NCCL all-reduce + adam_flat on every
     rank (bitwise at world 2, where the sum of two floats has one order), identical on all ranks; the Adam state of the
     OWNED slices matches;
  4. 200 back-to-back launches replayed from a CUDA graph (epoch protocol), timed.
Prints 'DP_MM ok' per rank; exits non-zero on failure."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ark_b200 import ops  # noqa: E402
from ark_b200.symm import SymmFlat  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 6_300_000 // 64 * 64 + 64
    sym = SymmFlat(n, dev, dist.group.WORLD)
    h = sym.hdl
    if rank == 0:
        print("hdl:", {k: getattr(h, k, None) for k in ("rank", "world_size", "buffer_size", "offset", "multicast_ptr",
                                                        "has_multicast_support", "signal_pad_size")}, flush=True)
        print("buffer_ptrs", [hex(int(p)) for p in h.buffer_ptrs], "data_ptr", hex(sym.buf.data_ptr()), flush=True)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    gc = torch.Generator(device=dev).manual_seed(7)
    p0 = torch.randn(n, device=dev, generator=gc)
    m0 = 0.01 * torch.randn(n, device=dev, generator=gc)
    v0 = (0.01 * torch.randn(n, device=dev, generator=gc)) ** 2
    grad = torch.randn(n, device=dev, generator=g)
    spans = [(0, 1024), (4096, 3_000_000), (3_000_064, n - 60)]
    spans = [(s, (e + 3) // 4 * 4) for s, e in spans]

    # ---- reference: NCCL all-reduce + replicated Adam
    g_ref = grad.clone()
    dist.all_reduce(g_ref)
    p_ref, m_ref, v_ref = p0.clone(), m0.clone(), v0.clone()
    s_ref = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    for s, e in spans:
        ops.adam_flat(p_ref[s:e], g_ref[s:e], m_ref[s:e], v_ref[s:e], s_ref[s:e], 1e-3, 0.9, 0.999, 1e-8, 3)

    # ---- mode 0
    sym.grad.copy_(grad)
    torch.cuda.synchronize()
    ops.dp_reduce_adam(sym, spans, m0, v0, 0)
    torch.cuda.synchronize()
    for s, e in spans:
        if world == 2:
            assert torch.equal(sym.grad[s:e], g_ref[s:e]), "mode 0: switch sum != NCCL sum"
        else:
            torch.testing.assert_close(sym.grad[s:e], g_ref[s:e], rtol=1e-5, atol=1e-5)
    untouched = torch.ones(n, dtype=torch.bool, device=dev)
    for s, e in spans:
        untouched[s:e] = False
    assert torch.equal(sym.grad[untouched], grad[untouched]), "mode 0 wrote outside the spans"

    # ---- mode 1
    sym.grad.copy_(grad)
    sym.param.copy_(p0)
    sym.shadow.zero_()
    m, v = m0.clone(), v0.clone()
    torch.cuda.synchronize()
    dist.barrier()
    ops.dp_reduce_adam(sym, spans, m, v, 1, 1e-3, 0.9, 0.999, 1e-8, step=3)
    torch.cuda.synchronize()
    for s, e in spans:
        # (not bitwise: the two kernels' FMA contraction may differ in the last bit)
        torch.testing.assert_close(sym.param[s:e], p_ref[s:e], rtol=2e-6 if world == 2 else 1e-4, atol=1e-7 if world == 2 else 1e-5)
        torch.testing.assert_close(sym.shadow[s:e].float(), s_ref[s:e].float(), rtol=1e-2, atol=1e-3)
        if rank == 0:
            print("mode 1 span", (s, e), "max |dp|", (sym.param[s:e] - p_ref[s:e]).abs().max().item(), flush=True)
        lo, hi = sym.owned(s, e)[rank]
        torch.testing.assert_close(m[lo:hi], m_ref[lo:hi], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(v[lo:hi], v_ref[lo:hi], rtol=1e-5, atol=1e-9)
    assert torch.equal(sym.param[untouched], p0[untouched]), "mode 1 wrote outside the spans"
    pb = sym.param.clone()
    dist.broadcast(pb, 0)
    assert torch.equal(pb, sym.param), "ranks hold different parameters"

    # ---- graph replay + timing (one 25 MB bucket: a GRU layer's gradients)
    hyper = torch.tensor([1e-3, 1.0], device=dev)
    big = [(0, n)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            ops.dp_reduce_adam(sym, big, m, v, 1, hyper=hyper)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(10):
                ops.dp_reduce_adam(sym, big, m, v, 1, hyper=hyper)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pb = sym.param.clone()
    dist.broadcast(pb, 0)
    assert torch.equal(pb, sym.param), "ranks diverged over 200 replayed launches"
    assert torch.isfinite(sym.param).all()
    x = torch.zeros(n, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"NCCL all_reduce of the same bucket: {e0.elapsed_time(e1) / 20:.4f} ms (+ adam_flat on every rank)", flush=True)
    for ctas in (16, 32, 64, 128):
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(20):
            ops.dp_reduce_adam(sym, big, m, v, 1, hyper=hyper, ctas=ctas)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"ctas {ctas}: {tt.item():.4f} ms per {4 * n / 1e6:.1f} MB bucket = {4 * n / 1e6 / tt.item():.0f} GB/s algbw", flush=True)
    print(f"DP_MM ok rank {rank}/{world}: {t.item():.4f} ms per {4 * n / 1e6:.1f} MB bucket (graph replay)", flush=True)
    del graph
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
