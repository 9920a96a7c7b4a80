// K7: fused softmax cross-entropy forward + backward over packed (non-PAD) token rows.
// Replaces F.cross_entropy(logits.reshape(-1,V), tgt, ignore_index=PAD) and its autograd backward
// (kgvae/experiments/ablation_study.py:64-69): log_softmax + nll_loss forward and two backward passes
// over [N,V] become ONE DRAM read and ONE in-place DRAM write; the second read of the row is served
// by L2 (a row is at most 122 KB; 148 resident rows = 18 MB of the 126 MB L2).  The probability
// matrix never exists in HBM.  HBM-bound: algorithmic bytes = 2*N*V*sizeof(logit).
// (Tried and dropped: holding the whole bf16 row in registers between the passes — one 512-thread CTA per SM at
// 128 registers/thread — removes the L2 re-read but also the overlap between a row's load and its neighbour's
// compute/store: 3.65-3.84 TB/s stand-alone against 4.26-4.53 TB/s for this two-pass kernel at V = 60943.)
#include "common.cuh"
#include <stdlib.h>

namespace ark {

constexpr int kCeThreadsDefault = 512;

constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2(float x) {   // 2^x on the MUFU pipe; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The kernel is as much instruction-issue bound as HBM bound (a bf16 logit is 2 bytes of traffic), so the
// inner loops are written to cost ~5.6 (pass 1) + ~3.6 (pass 2) instructions per element: one FFMA feeds the
// MUFU directly (x*log2e - max*log2e), the gradient scale is folded into the exponent.
struct OnlineLse {
  float m, s;  // running max, running sum of exp(x - m)
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void push8(const float* x) {
    float mx = x[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, x[i]);
    const float mn = fmaxf(m, mx);
    const float nb = -mn * kLog2e;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += ex2(fmaf(x[i], kLog2e, nb));
    s = fmaf(s, ex2(fmaf(m, kLog2e, nb)), acc);  // ex2(-inf) = 0 on the first push
    m = mn;
  }
  __device__ __forceinline__ void push(float x) {
    const float mn = fmaxf(m, x);
    const float nb = -mn * kLog2e;
    s = fmaf(s, ex2(fmaf(m, kLog2e, nb)), ex2(fmaf(x, kLog2e, nb)));
    m = mn;
  }
};

__device__ __forceinline__ void load8(const uint16_t* p, float* x) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void load8(const float* p, float* x) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void store8(uint16_t* p, const float* x) {
  uint4 v;
  v.x = pack_bf16x2(x[0], x[1]); v.y = pack_bf16x2(x[2], x[3]);
  v.z = pack_bf16x2(x[4], x[5]); v.w = pack_bf16x2(x[6], x[7]);
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void store8(float* p, const float* x) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
}
// Cache policy of the in-place gradient pass (measured with ncu on wd-articles, 592 resident rows x 122 KB: with default
// policies the gradient rows being written evicted logits rows still waiting for their second read — 1.05 GB of DRAM reads
// for 0.6 GB of logits): the second read of a row is its LAST use and the gradient row is not read again by this kernel, so
// both are streaming accesses (evict-first); only the first read keeps the default policy.
__device__ __forceinline__ void load8_stream(const uint16_t* p, float* x) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void load8_stream(const float* p, float* x) {
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "l"(p));
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7]) : "l"(p + 4));
}
__device__ __forceinline__ void store8_stream(uint16_t* p, const float* x) {
  const uint32_t a = pack_bf16x2(x[0], x[1]), b = pack_bf16x2(x[2], x[3]), c = pack_bf16x2(x[4], x[5]), d = pack_bf16x2(x[6], x[7]);
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void store8_stream(float* p, const float* x) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]) : "memory");
}
__device__ __forceinline__ float ld1(const uint16_t* p) { return bf16_bits_to_f32(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(uint16_t* p, float x) { *p = f32_to_bf16_bits(x); }
__device__ __forceinline__ void st1(float* p, float x) { *p = x; }

// one CTA per row (grid-stride over rows).  Requires ldv % 8 == 0 and a 16B-aligned base so that every
// row starts on a 16-byte boundary: all accesses inside [0, V8) are 128-bit.
// U = independent 16-byte loads per thread in flight (4 for long rows, 2 for short ones)
template <typename T, int U, int kCeThreads = kCeThreadsDefault, int kMinCtas = (U == 4 ? 2 : 4)>
__global__ void __launch_bounds__(kCeThreads, kMinCtas) softmax_ce_kernel(
    T* __restrict__ logits, int64_t N, int V, int64_t ldv, const int32_t* __restrict__ tgt, float grad_scale,
    int write_grad, float* __restrict__ loss_acc, float* __restrict__ lse_out) {
  __shared__ float red[33];
  const int V8 = V & ~7;
  float loss_local = 0.f;  // only meaningful in thread 0
  for (int64_t row = blockIdx.x; row < N; row += gridDim.x) {
    T* x = logits + row * ldv;
    OnlineLse st;
    st.init();
    {
      // bytes in flight decide the DRAM rate here: 4 independent 16-byte loads per thread before any math
      constexpr int S = kCeThreads * 8;
      int c = threadIdx.x * 8;
      for (; c + (U - 1) * S < V8; c += U * S) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) load8(x + c + u * S, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) st.push8(v[u]);
      }
      for (; c < V8; c += S) {
        float v[8];
        load8(x + c, v);
        st.push8(v);
      }
    }
    for (int c = V8 + threadIdx.x; c < V; c += kCeThreads) st.push(ld1(x + c));
    const float M = block_max(st.m, red);
    const float contrib = (st.m == -INFINITY) ? 0.f : st.s * ex2((st.m - M) * kLog2e);
    const float S = block_sum(contrib, red);
    const float lse = M + __logf(S);
    const int t = tgt[row];
    if (threadIdx.x == 0) {
      loss_local += lse - ld1(x + t);
      if (lse_out) lse_out[row] = lse;
    }
    if (write_grad) {
      __syncthreads();  // thread 0 has read x[t] before anyone overwrites it
      // grad = scale*exp(x - lse) = 2^(x*log2e + kk), kk = log2(scale) - lse*log2e
      const float kk = fmaf(-lse, kLog2e, __log2f(grad_scale));
      constexpr int S = kCeThreads * 8;
      int c = threadIdx.x * 8;
      for (; c + (U - 1) * S < V8; c += U * S) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) load8_stream(x + c + u * S, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[u][i] = ex2(fmaf(v[u][i], kLog2e, kk));
          if ((unsigned)(t - (c + u * S)) < 8u) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i == t - (c + u * S)) v[u][i] -= grad_scale;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) store8_stream(x + c + u * S, v[u]);
      }
      for (; c < V8; c += kCeThreads * 8) {
        float v[8];
        load8_stream(x + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ex2(fmaf(v[i], kLog2e, kk));
        if ((unsigned)(t - c) < 8u) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == t - c) v[i] -= grad_scale;
        }
        store8_stream(x + c, v);
      }
      for (int c = V8 + threadIdx.x; c < (int)ldv; c += kCeThreads) {
        float g = 0.f;
        if (c < V) {
          g = ex2(fmaf(ld1(x + c), kLog2e, kk));
          if (c == t) g -= grad_scale;
        }
        st1(x + c, g);
      }
    }
  }
  if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, loss_local * grad_scale);
}

}  // namespace ark

using namespace ark;

extern "C" int ark_softmax_ce(void* logits, int dtype, int64_t N, int64_t V, int64_t ldv, const int32_t* tgt,
                              float grad_scale, int write_grad, float* loss_acc, float* lse, void* stream) {
  ARK_REQUIRE(logits && tgt, ARK_E_BADARG, "softmax_ce: null pointer");
  ARK_REQUIRE(N >= 0 && V > 0 && ldv >= V, ARK_E_BADARG, "softmax_ce: bad sizes");
  ARK_REQUIRE(ldv % 8 == 0 && aligned16(logits), ARK_E_ALIGN,
              "softmax_ce: ldv=%lld must be a multiple of 8 and the base 16-byte aligned", (long long)ldv);
  if (N == 0) return 0;
  // 2 CTAs of 512 threads per SM keep one row loading while another computes/stores
  const bool deep = V >= 40000;   // long rows: 4 loads in flight per thread, 2 CTAs/SM; short rows: 2 loads, 4 CTAs/SM
  cudaStream_t s = (cudaStream_t)stream;
  // tuning knob (tools/bench_ce.py): ARK_CE_VARIANT = threads*100 + U*10 + CTAs/SM, e.g. 25643 = 256 threads, U 4, 3/SM
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("ARK_CE_VARIANT"); variant = e ? atoi(e) : 0; }
#define ARK_CE_V(TT, UU, TH, PS)                                                                                        \
  {                                                                                                                     \
    const unsigned g_ = (unsigned)(N < (int64_t)(PS) * kNumSMs ? N : (int64_t)(PS) * kNumSMs);                          \
    softmax_ce_kernel<TT, UU, TH, PS><<<g_, TH, 0, s>>>((TT*)logits, N, (int)V, ldv, tgt, grad_scale, write_grad, loss_acc, lse); \
    return launched("softmax_ce");                                                                                      \
  }
  if (variant && dtype == ARK_BF16) {
    switch (variant) {
      case 51242: ARK_CE_V(uint16_t, 4, 512, 2)
      case 51243: ARK_CE_V(uint16_t, 4, 512, 3)
      case 51224: ARK_CE_V(uint16_t, 2, 512, 4)
      case 25644: ARK_CE_V(uint16_t, 4, 256, 4)
      case 25646: ARK_CE_V(uint16_t, 4, 256, 6)
      case 25628: ARK_CE_V(uint16_t, 2, 256, 8)
      case 25626: ARK_CE_V(uint16_t, 2, 256, 6)
      case 102441: ARK_CE_V(uint16_t, 4, 1024, 1)
      case 102422: ARK_CE_V(uint16_t, 2, 1024, 2)
      default: break;
    }
  }
  // measured stand-alone on B200 (tools/bench_ce.py, bf16, GB/s of algorithmic bytes; 70 % of the 6552 GB/s copy peak = 4586):
  //   V 60943, N 4966 : 512 thr U4 2/SM 4435 | 512 thr U2 4/SM 4767 | 1024 thr U2 2/SM 4636
  //   V 60943, N 20000: 512 thr U4 2/SM 4583 | 512 thr U2 4/SM 4808 | 1024 thr U2 2/SM 5050
  //   V 24101, N 10333: 512 thr U2 4/SM 4928 | 256 thr U2 8/SM 5042
  // -> more resident rows per SM beat more loads in flight per thread: a row's block reductions and pass switch are
  //    covered by its neighbours
  if (dtype == ARK_BF16 && !variant) {
    if (!deep) ARK_CE_V(uint16_t, 2, 256, 8)
    if (N >= 16384) ARK_CE_V(uint16_t, 2, 1024, 2)
    ARK_CE_V(uint16_t, 2, 512, 4)
  }
  const int per_sm = deep ? 2 : 4;
  const unsigned grid = (unsigned)(N < per_sm * kNumSMs ? N : per_sm * kNumSMs);
#define ARK_CE_GO(TT, UU) softmax_ce_kernel<TT, UU><<<grid, kCeThreadsDefault, 0, s>>>((TT*)logits, N, (int)V, ldv, tgt, grad_scale, write_grad, loss_acc, lse)
  if (dtype == ARK_BF16) {
    if (deep) ARK_CE_GO(uint16_t, 4); else ARK_CE_GO(uint16_t, 2);
  } else if (dtype == ARK_F32) {
    if (deep) ARK_CE_GO(float, 4); else ARK_CE_GO(float, 2);
  } else {
    return fail(ARK_E_BADARG, "softmax_ce: unknown dtype %d", dtype);
  }
#undef ARK_CE_GO
#undef ARK_CE_V
  return launched("softmax_ce");
}
