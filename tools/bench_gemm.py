"""Micro-benchmark of the tcgen05 GEMM on the step's shapes:  python tools/bench_gemm.py [name ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ark_b200 import ops
K_, MN = ops.MAJOR_K, ops.MAJOR_MN
SHAPES = {  # name: (M, N, K, a_major, b_major, c_dtype, bias)
    "wdm_vocab_fwd": (10333, 24101, 128, K_, K_, torch.bfloat16, True),
    "wda_vocab_fwd": (4966, 60943, 512, K_, K_, torch.bfloat16, True),
    "wda_vocab_dY": (4966, 512, 60943, K_, MN, torch.float32, False),
    "wda_vocab_dW": (60943, 512, 4966, MN, MN, torch.float32, False),
    "syn_gru_gi": (2560, 3072, 1024, K_, K_, torch.float32, True),
    "syn_gru_dW": (3072, 1024, 2560, MN, MN, torch.float32, False),
    "syn_mlp_dW": (3072, 3072, 256, MN, MN, torch.float32, False),
    "syn_mlp_fwd": (256, 3072, 3072, K_, K_, torch.bfloat16, True),
}
names = sys.argv[1:] or list(SHAPES)
for nm in names:
    M, N, K, am, bm, cdt, bias = SHAPES[nm]
    up8 = lambda x: (x + 7) // 8 * 8
    A = (torch.randn(M, up8(K), device="cuda") if am == K_ else torch.randn(K, up8(M), device="cuda")).to(torch.bfloat16)
    B = (torch.randn(N, up8(K), device="cuda") if bm == K_ else torch.randn(K, up8(N), device="cuda")).to(torch.bfloat16)
    A = A[:, :K] if am == K_ else A[:, :M]
    B = B[:, :K] if bm == K_ else B[:, :N]
    C = torch.empty(M, up8(N), device="cuda", dtype=cdt)[:, :N]
    b = torch.randn(N, device="cuda") if bias else None
    for _ in range(3):
        ops.gemm(A, am, B, bm, C, M, N, K, bias=b, backend="tc")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(A, am, B, bm, C, M, N, K, bias=b, backend="tc")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out_b = M * N * (2 if cdt == torch.bfloat16 else 4)
    print(f"{nm:16s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s  out {out_b/ms/1e6:6.0f} GB/s")
