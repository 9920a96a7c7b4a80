#!/usr/bin/env python
"""Stand-alone bandwidth of the HBM-bound kernels of the ELBO step at sizes where a roofline fraction is
meaningful (in situ several of them move < 1 MB and are launch-latency bound; SURVEY.md §8d asks for both views).

    python tools/bench_hbm_kernels.py            # prints one JSON line per kernel

Algorithmic bytes follow DESIGN.md §4; time = CUDA events on the launching stream, median of 7 after 3 warm-ups,
inputs larger than the 126 MB L2 (or a fresh buffer per iteration where the kernel works in place).
Peak = MEASURED_PEAKS.json hbm_gbs (burst: kernels timed alone).
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ark_b200 import ops  # noqa: E402

DEV = "cuda"


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured"
    except OSError:
        return 6650.0, "fallback"


def timed(fn, n_iter=7, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(n_iter):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(warm + i)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(name, nbytes, ms, note):
    pk, src = peak()
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GB/s": round(gbs),
                      "frac_of_hbm_peak": round(gbs / pk, 3), "peak_GB/s": pk, "peak_src": src, "shape": note}))


def main():
    torch.manual_seed(0)
    # ---- K1/K2 gather + masked mean-pool: wd-articles-shaped table, large batch (256 graphs x 212 triples)
    nE, nR, d, B, T = 60933, 7, 512, 256, 212
    E, R = torch.randn(nE, d, device=DEV), torch.randn(nR, d, device=DEV)
    tri = torch.stack([torch.randint(0, nE - 1, (B, T)), torch.randint(0, nR - 1, (B, T)), torch.randint(0, nE - 1, (B, T))], -1).to(DEV)
    g_b, inv = torch.empty(B, 3 * d, device=DEV, dtype=torch.bfloat16), torch.empty(B, device=DEV)
    ms = timed(lambda i: ops.gather_pool_fwd(tri, None, E, R, nR - 1, None, g_b, inv))
    report("gather_pool_fwd", 3.0 * B * T * d * 4 + 3.0 * B * T * 8 + B * 3 * d * 2, ms, f"B={B} T={T} d={d} nE={nE}")
    dg = torch.randn(B, 3 * d, device=DEV)
    dE, dR = torch.zeros(nE, d, device=DEV), torch.zeros(nR, d, device=DEV)
    ms = timed(lambda i: ops.gather_pool_bwd(dg, tri, None, inv, nR - 1, nE - 1, dE, dR))
    report("gather_pool_bwd", B * 3 * d * 4 + 3.0 * B * T * d * 8, ms, f"B={B} T={T} d={d} (fp32 RMW of gradient rows)")

    # ---- K4 reparameterisation + KL at B*dz = 2^25 (in situ it is <= 262 KB: launch-bound)
    Bz, dz = 1 << 17, 256
    heads, eps = torch.randn(Bz, 2 * dz, device=DEV), torch.randn(Bz, dz, device=DEV)
    z, z_b, kl = torch.empty(Bz, dz, device=DEV), torch.empty(Bz, dz, device=DEV, dtype=torch.bfloat16), torch.zeros(1, device=DEV)
    ms = timed(lambda i: ops.reparam_kl_fwd(heads, eps, None, dz, True, 1.0 / (Bz * dz), z, z_b, kl))
    report("reparam_kl_fwd", Bz * dz * (3 * 4 + 4 + 2), ms, f"B={Bz} dz={dz}")
    dz_in = torch.randn(Bz, dz, device=DEV)
    dh, dh_b = torch.empty(Bz, 2 * dz, device=DEV), torch.empty(Bz, 2 * dz, device=DEV, dtype=torch.bfloat16)
    ms = timed(lambda i: ops.reparam_kl_bwd(heads, eps, None, dz_in, dz, True, 0.5 / (Bz * dz), dh, dh_b))
    report("reparam_kl_bwd", Bz * dz * (4 * 4 + 2 * 4 + 2 * 2), ms, f"B={Bz} dz={dz}")

    # ---- K5 token gather / scatter-add (wd-articles large batch: 163072 rows, d = 512)
    V, N = 60943, 163072
    Wb = torch.randn(V, d, device=DEV).to(torch.bfloat16)
    tok = torch.randint(0, V, (N,), device=DEV, dtype=torch.int32)
    xb = torch.empty(N, d, device=DEV, dtype=torch.bfloat16)
    ms = timed(lambda i: ops.tok_gather_fwd(Wb, tok, None, xb))
    report("tok_gather_fwd", N * d * 2 * 2 + N * 4, ms, f"N={N} d={d} V={V} (bf16 rows)")
    dX, dW = torch.randn(N, d, device=DEV), torch.zeros(V, d, device=DEV)
    ms = timed(lambda i: ops.tok_scatter_add(dX, tok, dW))
    report("tok_scatter_add", N * d * 12.0, ms, f"N={N} d={d} V={V}")

    # ---- K7 softmax cross-entropy, forward+backward in place (wd-articles ragged / dense-sized rows)
    for (Nr, Vv) in [(4966, 60943), (20000, 60943)]:
        ldv = (Vv + 7) // 8 * 8
        src = torch.randn(Nr, ldv, device=DEV).to(torch.bfloat16)
        bufs = [src.clone() for _ in range(10)]
        tgt = torch.randint(1, Vv, (Nr,), device=DEV, dtype=torch.int32)
        loss = torch.zeros(1, device=DEV)
        ms = timed(lambda i: ops.softmax_ce(bufs[i], Vv, tgt, 1e-4, True, loss, None))
        report("softmax_ce", 2.0 * Nr * Vv * 2 + 12.0 * Nr, ms, f"N={Nr} V={Vv} bf16 in place")

    # ---- K10 Adam over a flat buffer (wd-articles: 74.7M parameters)
    P = 74734848
    p, g, m, v = (torch.randn(P, device=DEV) for _ in range(4))
    v.abs_()
    sh = torch.empty(P, device=DEV, dtype=torch.bfloat16)
    ms = timed(lambda i: ops.adam_flat(p, g, m, v, sh, 1e-3, 0.9, 0.999, 1e-8, i + 1))
    report("adam_flat", 30.0 * P, ms, f"P={P} (28 B/param + bf16 shadow)")


if __name__ == "__main__":
    main()
