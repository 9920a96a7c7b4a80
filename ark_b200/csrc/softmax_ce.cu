// K7: fused softmax cross-entropy forward + backward over packed (non-PAD) token rows.
// Replaces F.cross_entropy(logits.reshape(-1,V), tgt, ignore_index=PAD) and its autograd backward
// (kgvae/experiments/ablation_study.py:64-69): log_softmax + nll_loss forward and two backward passes
// over [N,V] become ONE DRAM read and ONE in-place DRAM write; the second read of the row is served
// by L2 (a row is at most 122 KB; 148 resident rows = 18 MB of the 126 MB L2).  The probability
// matrix never exists in HBM.  HBM-bound: algorithmic bytes = 2*N*V*sizeof(logit).
#include "common.cuh"

namespace ark {

constexpr int kCeThreads = 512;

struct OnlineLse {
  float m, s;  // running max, running sum of exp(x - m)
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  __device__ __forceinline__ void push8(const float* x) {
    float mx = x[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, x[i]);
    const float mn = fmaxf(m, mx);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __expf(x[i] - mn);
    s = s * __expf(m - mn) + acc;  // exp(-inf - finite) = 0 on the first push
    m = mn;
  }
  __device__ __forceinline__ void push(float x) {
    const float mn = fmaxf(m, x);
    s = s * __expf(m - mn) + __expf(x - mn);
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
__device__ __forceinline__ float ld1(const uint16_t* p) { return bf16_bits_to_f32(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(uint16_t* p, float x) { *p = f32_to_bf16_bits(x); }
__device__ __forceinline__ void st1(float* p, float x) { *p = x; }

// one CTA per row (grid-stride over rows).  Requires ldv % 8 == 0 and a 16B-aligned base so that every
// row starts on a 16-byte boundary: all accesses inside [0, V8) are 128-bit.
template <typename T>
__global__ void __launch_bounds__(kCeThreads) softmax_ce_kernel(
    T* __restrict__ logits, int64_t N, int V, int64_t ldv, const int32_t* __restrict__ tgt, float grad_scale,
    int write_grad, float* __restrict__ loss_acc, float* __restrict__ lse_out) {
  __shared__ float red[33];
  const int V8 = V & ~7;
  float loss_local = 0.f;  // only meaningful in thread 0
  for (int64_t row = blockIdx.x; row < N; row += gridDim.x) {
    T* x = logits + row * ldv;
    OnlineLse st;
    st.init();
    for (int c = threadIdx.x * 8; c < V8; c += kCeThreads * 8) {
      float v[8];
      load8(x + c, v);
      st.push8(v);
    }
    for (int c = V8 + threadIdx.x; c < V; c += kCeThreads) st.push(ld1(x + c));
    const float M = block_max(st.m, red);
    const float contrib = (st.m == -INFINITY) ? 0.f : st.s * __expf(st.m - M);
    const float S = block_sum(contrib, red);
    const float lse = M + __logf(S);
    const int t = tgt[row];
    if (threadIdx.x == 0) {
      loss_local += lse - ld1(x + t);
      if (lse_out) lse_out[row] = lse;
    }
    if (write_grad) {
      __syncthreads();  // thread 0 has read x[t] before anyone overwrites it
      for (int c = threadIdx.x * 8; c < V8; c += kCeThreads * 8) {
        float v[8];
        load8(x + c, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = grad_scale * __expf(v[i] - lse);
        if (t >= c && t < c + 8) v[t - c] -= grad_scale;
        store8(x + c, v);
      }
      for (int c = V8 + threadIdx.x; c < (int)ldv; c += kCeThreads) {
        float g = 0.f;
        if (c < V) {
          g = grad_scale * __expf(ld1(x + c) - lse);
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
  const unsigned grid = (unsigned)(N < 4 * kNumSMs ? N : 4 * kNumSMs);
  if (dtype == ARK_BF16)
    softmax_ce_kernel<uint16_t><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>(
        (uint16_t*)logits, N, (int)V, ldv, tgt, grad_scale, write_grad, loss_acc, lse);
  else if (dtype == ARK_F32)
    softmax_ce_kernel<float><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>((float*)logits, N, (int)V, ldv, tgt,
                                                                           grad_scale, write_grad, loss_acc, lse);
  else
    return fail(ARK_E_BADARG, "softmax_ce: unknown dtype %d", dtype);
  return launched("softmax_ce");
}
