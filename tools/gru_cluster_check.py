"""A/B check of the cluster GRU stack kernel (csrc/gru_cluster.cu) against the wavefront kernel on one GPU:
loss / gradient agreement and per-launch times.   python tools/gru_cluster_check.py [d B nl hi]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from ark_b200 import ops  # noqa: E402
from ark_b200.layout import pack_layout  # noqa: E402
from kgvae.model.models import SAIL  # noqa: E402
from oracle import sail_oracle as O  # noqa: E402  (batch builder only)


def main():
    d, B, nl, hi = (int(a) for a in (sys.argv[1:5] + ["512", "16", "3", "40"][len(sys.argv) - 1:]))
    p_drop = float(os.environ.get("P_DROP", "0.0"))
    rng = np.random.default_rng(13)
    nE, nR = 300, 6
    layv = O.vocab_layout(nE, nR, hi, True)
    graphs = [[(int(rng.integers(nE)), int(rng.integers(nR)), int(rng.integers(nE))) for _ in range(int(rng.integers(2, hi + 1)))]
              for _ in range(B)]
    tri, seq = O.build_batch(graphs, layv)
    cfg = dict(layv, model_type="SAIL", d_model=d, d_latent=16, n_heads=2, n_layers=nl, dec_dropout=p_drop)
    print("cluster supported NB =", ops.gru_cluster_supported(d, B, nl, seq.shape[1] - 1), " wave dj =", ops.gru_wave_supported(d, B, nl), flush=True)
    eps = torch.from_numpy(rng.standard_normal((B, 16)).astype(np.float32)).cuda()
    seq_t = torch.from_numpy(seq)
    lay = pack_layout(seq_t).to("cuda")
    res = {}
    for mode in ("wave", "cluster"):
        torch.manual_seed(4)
        model = SAIL(dict(cfg)).cuda()
        eng = model.engine(seed=11)
        eng.gru_mode = mode
        t0 = time.time()
        out = eng.forward_backward(torch.from_numpy(tri).cuda(), seq_t.cuda(), lay, eps, 0.5).clone()
        torch.cuda.synchronize()
        print(mode, "out", out.tolist(), f"{time.time() - t0:.2f}s", flush=True)
        res[mode] = (out, eng.flat.grad.clone(), eng)
        eng.forward_backward(torch.from_numpy(tri).cuda(), seq_t.cuda(), lay, eps, 0.5)
        eng.prof = []
        for _ in range(3):
            eng.forward_backward(torch.from_numpy(tri).cuda(), seq_t.cuda(), lay, eps, 0.5)
        for kname, v in eng.profile_summary().items():
            if "gru_" in kname and "gemm" not in kname:
                print(f"    {kname}: {v['ms'] / v['calls']:.4f} ms/launch ({lay.L} steps)", flush=True)
        eng.prof = None
    if os.environ.get("ARK_GRU_CLUSTER_DBG"):
        import ctypes
        from ark_b200 import _C
        n = 2 * 8 * 3 * 4 * 16
        buf = (ctypes.c_int64 * n)()
        _C.lib().call("ark_gru_cluster_debug_dump", ctypes.cast(buf, ctypes.c_void_p), n)
        tl = np.frombuffer(buf, dtype=np.int64).reshape(2, 8, 3, 4, 16)
        for di, dname in enumerate(("fwd", "bwd")):
            for zi in range(8):
                blk = tl[di, zi]
                if not blk.any():
                    continue
                t0 = blk[blk > 0].min()
                stage = ("rec" if zi < nl else "proj") + str(zi if zi < nl else zi - nl)
                for ri, rname in enumerate(("loader", "mma", "epi")):
                    for it in range(4):
                        pts = blk[ri, it]
                        if pts.any():
                            print(f"  {dname} {stage} {rname:6s} it+{it}: " +
                                  " ".join(f"{int(v - t0):6d}" if v > 0 else "     ." for v in pts[:13]))
    a, b = res["cluster"], res["wave"]
    print("loss diff", (a[0] - b[0]).abs().tolist())
    print("grad rel (all)", ((a[1] - b[1]).norm() / b[1].norm()).item())
    for name in b[2].flat.order:
        ga, gb = a[2].flat.g(name), b[2].flat.g(name)
        if gb.norm() > 0:
            print(f"   {name:32s} rel {((ga - gb).norm() / gb.norm()).item():.3e}")


if __name__ == "__main__":
    main()
