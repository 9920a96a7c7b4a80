"""Stand-alone timing + clock64 timeline of the per-layer persistent GRU kernels (csrc/gru_persist.cu).
python tools/gru_persist_bench.py [d B L]   (env: ARK_GRU_STAGES, ARK_GRU_CLUSTER, ARK_GRU_PERSIST_DBG=1)"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ark_b200 import _C, ops  # noqa: E402

d, B, L = (int(a) for a in (sys.argv[1:4] + ["1024", "256", "10"][len(sys.argv) - 1:]))
DEV, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
N = B * L
bt = np.full(L, B, dtype=np.int32)
off = (np.arange(L + 1) * B).astype(np.int32)
bt_d, off_d = torch.from_numpy(bt).to(DEV), torch.from_numpy(off[:-1].copy()).to(DEV)
Wb = (torch.randn(3 * d, d, device=DEV) / d ** 0.5).to(bf).contiguous()
WT = torch.empty(d, 3 * d, device=DEV, dtype=bf)
ops.transpose_bf16(Wb, WT)
gi = torch.randn(N, 3 * d, device=DEV)
b_hh = torch.zeros(3 * d, device=DEV)
h0 = torch.tanh(torch.randn(B, d, device=DEV))
hp_b = torch.zeros(N, d, device=DEV, dtype=bf)
hp_b[:B] = h0.to(bf)
y_b = torch.empty(N, d, device=DEV, dtype=bf)
gates = tuple(torch.empty(N, d, device=DEV, dtype=bf) for _ in range(4))
sync = torch.empty((B + 127) // 128 * 64, device=DEV, dtype=torch.int32)
dy = torch.randn(N, d, device=DEV)
dgi = torch.empty(N, 3 * d, device=DEV, dtype=bf)
dgh = torch.empty(N, 3 * d, device=DEV, dtype=bf)
dh0 = torch.zeros(B, d, device=DEV)


def fwd():
    ops.gru_persist_fwd(hp_b, h0, Wb, gi, b_hh, bt_d, off_d, L, B, d, y_b, gates, sync)


def bwd():
    ops.gru_persist_bwd(dy, gates, hp_b, WT, bt_d, off_d, L, B, d, dgi, dgh, dh0, False, sync, Whh_b=Wb)


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


tf, tb = timeit(fwd), timeit(bwd)
chk = (y_b.float().abs().mean().item(), dgh.float().abs().mean().item())
print(f"d={d} B={B} L={L} stages={os.environ.get('ARK_GRU_STAGES', 'def')} cluster={os.environ.get('ARK_GRU_CLUSTER', 'def')} "
      f"variant={os.environ.get('ARK_GRU_VARIANT', 'def')}: "
      f"fwd {tf:.1f} us ({tf / L:.2f}/step)  bwd {tb:.1f} us ({tb / L:.2f}/step)  chk {chk[0]:.5f} {chk[1]:.5f}", flush=True)
if os.environ.get("ARK_GRU_PERSIST_DBG"):
    buf = (ctypes.c_int64 * 64)()
    _C.lib().call("ark_gru_persist_debug_dump", ctypes.cast(buf, ctypes.c_void_p), 64)
    tl = np.frombuffer(buf, dtype=np.int64).reshape(2, 4, 8)
    names = ["ctr_ok", "tma_issued", "first_full", "mma_issued", "tmem_full", "ld_done", "math_done", "released"]
    for di, dn in enumerate(("fwd", "bwd")):
        blk = tl[di]
        if not blk.any():
            continue
        t0 = blk[blk > 0].min()
        print(f"  {dn} timeline (cycles from first event; CTA (0,0), 4 consecutive steps):")
        for it in range(4):
            print("    " + "  ".join(f"{n}={int(v - t0):6d}" if v > 0 else f"{n}=     ." for n, v in zip(names, blk[it])))
