// K11 — data-parallel gradient exchange FUSED with the optimiser, over NVSwitch multicast (NVLink SHARP).
//
// The reference trains on one GPU (ablation_study.py:571-600: loss.backward(); optimizer.step()).  Under data parallelism
// the same update needs the gradient SUMMED over ranks.  The library way is all-reduce(grad) -> Adam on every rank: every
// rank receives the whole summed gradient over NVLink and then streams the whole p/g/m/v state through its HBM.  Here
// the flat param / grad / shadow buffers of every rank live in ONE symmetric allocation that is also mapped as an
// NVSwitch MULTICAST object, and one kernel per gradient bucket does
//
//     g_sum = multimem.ld_reduce(grad)          the switch adds the 1/world slice this rank owns over all ranks
//     p, m, v <- Adam(p, g_sum, m, v)           on that slice only: the Adam state is SHARDED (1/world of the HBM traffic)
//     multimem.st(param) ; multimem.st(shadow)  the switch writes the new fp32 parameters + bf16 operand copies to ALL ranks
//
// i.e. reduce-scatter + sharded Adam + all-gather with no intermediate buffer; per rank and parameter the NVLink carries
// 4 B out + 4/world B in for the reduction and 6/world B out + 6 B in for the broadcast (all-reduce: 4 out + 4 in, then
// 28 B of HBM traffic on every rank).  Every rank ends with bit-identical parameters (one owner computed them).
//
// Cross-rank ordering is two epoch-stamped flag exchanges through peer-mapped memory inside the kernel:
//   ready: "my gradients of this bucket are final" (the launch is stream-ordered behind their producers);
//   done : "I have read your gradients and my broadcast has been fenced" — the kernel does not end before every peer
//          said so, hence anything stream-ordered after it may rewrite the gradients / read the parameters.
// Epochs live in device memory, so the launch can be captured in and replayed from a CUDA graph.
#include "common.cuh"
#include "ptx.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace ark {

constexpr int kDpMaxRanks = 16;
constexpr int kDpMaxSpans = 16;
// words of a rank's flag block (uint32): [0,16) ready[src], [32,48) done[src], 64 = epoch of the last finished launch,
// 65 = blocks of the running launch that have finished
constexpr int kDpReady = 0, kDpDone = 32, kDpEpoch = 64, kDpBlocks = 65;

struct DpArgs {
  uint32_t* flags[kDpMaxRanks];      // every rank's flag block (peer-mapped addresses; [rank] = this rank's own)
  int64_t s[kDpMaxSpans], e[kDpMaxSpans];   // element ranges of the flat buffers, multiples of 4
  int n_spans, rank, world, mode;
  float* grad_mc;                    // multicast addresses
  float* param_mc;
  uint16_t* shadow_mc;
  const float* param;                // this rank's own buffers
  float* m;
  float* v;
  const float* hyper;
  float step_size, inv_sqrt_bc2, beta1, beta2, eps;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add_f32x4(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st_bf16x4(uint16_t* mc, uint32_t lo, uint32_t hi) {
  asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(mc), "r"(lo), "r"(hi) : "memory");
}

__device__ __forceinline__ void dp_wait_flags(const uint32_t* flag, uint32_t epoch, const char* what) {
  uint64_t t0 = 0;
  uint32_t n = 0;
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if ((++n & 0x3FFu) == 0) {
      const uint64_t now = ptx::globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 60000000000ull) {
        printf("arkb200: dp_reduce_adam: a peer did not signal '%s' within 60 s (block %d thread %d epoch %u)\n", what,
               blockIdx.x, threadIdx.x, epoch);
        __trap();
      }
    }
  }
}

constexpr int kDpThreads = 512;

// "ready" exchange: returns the epoch of this launch once every rank has entered its launch of the same call
__device__ __forceinline__ uint32_t dp_enter(uint32_t* const* flags, int rank, int world) {
  uint32_t* my = flags[rank];
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(my + kDpEpoch) + 1u;
  if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(flags[threadIdx.x] + kDpReady + rank, epoch);
  if (threadIdx.x < world) dp_wait_flags(my + kDpReady + threadIdx.x, epoch, "ready");
  __syncthreads();
  return epoch;
}
// "done" exchange: fence this block's multicast stores, count it in; the last block tells every peer and waits for them
__device__ __forceinline__ void dp_exit(uint32_t* const* flags, int rank, int world, uint32_t epoch) {
  uint32_t* my = flags[rank];
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(my + kDpBlocks, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < world) {
    st_release_sys(flags[threadIdx.x] + kDpDone + rank, epoch);
    dp_wait_flags(my + kDpDone + threadIdx.x, epoch, "done");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    my[kDpBlocks] = 0;
    my[kDpEpoch] = epoch;
  }
}

template <int kDpUnroll>
__global__ void __launch_bounds__(kDpThreads) dp_reduce_adam_kernel(const __grid_constant__ DpArgs a) {
  const uint32_t epoch = dp_enter(a.flags, a.rank, a.world);

  float step_size = a.step_size, inv_sqrt_bc2 = a.inv_sqrt_bc2;
  if (a.hyper) {      // CUDA-graph replay: the step-dependent scalars live in device memory
    step_size = a.hyper[0];
    inv_sqrt_bc2 = a.hyper[1];
  }
  const int64_t T = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int sp = 0; sp < a.n_spans; ++sp) {
    const int64_t n4 = (a.e[sp] - a.s[sp]) >> 2, per = (n4 + a.world - 1) / a.world;
    const int64_t lo = a.rank * per, hi = min(n4, lo + per);          // the vec4 slice this rank owns
    for (int64_t q0 = lo + tid; q0 < hi; q0 += T * kDpUnroll) {
      float4 g[kDpUnroll];
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {                            // all switch reductions of this round in flight
        const int64_t q = q0 + u * T;
        if (q < hi) g[u] = multimem_ld_reduce_add_f32x4(a.grad_mc + a.s[sp] + 4 * q);
      }
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {
        const int64_t q = q0 + u * T;
        if (q >= hi) continue;
        const int64_t i = a.s[sp] + 4 * q;
        if (a.mode == 0) {          // gradient all-reduce only: every rank receives the sum
          multimem_st_f32x4(a.grad_mc + i, g[u]);
          continue;
        }
        const float4 pv = *reinterpret_cast<const float4*>(a.param + i);
        const float4 mv = *reinterpret_cast<const float4*>(a.m + i), vv = *reinterpret_cast<const float4*>(a.v + i);
        float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
        float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vv4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // torch.optim.Adam (same arithmetic as adam_flat_kernel)
          mm[k] = a.beta1 * mm[k] + (1.f - a.beta1) * gg[k];
          vv4[k] = a.beta2 * vv4[k] + (1.f - a.beta2) * gg[k] * gg[k];
          pp[k] -= step_size * mm[k] / (sqrtf(vv4[k]) * inv_sqrt_bc2 + a.eps);
        }
        *reinterpret_cast<float4*>(a.m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(a.v + i) = make_float4(vv4[0], vv4[1], vv4[2], vv4[3]);
        multimem_st_f32x4(a.param_mc + i, make_float4(pp[0], pp[1], pp[2], pp[3]));
        multimem_st_bf16x4(a.shadow_mc + i, pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
      }
    }
  }

  dp_exit(a.flags, a.rank, a.world, epoch);
}

// All-gather by multicast store: every rank writes its items into ITS slot of the gathered buffers of all ranks (the
// switch replicates the stores); same ready / done exchange, so a rank's previous readers of the gathered buffers have
// finished (their launch of this call is stream-ordered behind them) before anyone overwrites them.
constexpr int kDpMaxItems = 8;
struct DpGatherArgs {
  uint32_t* flags[kDpMaxRanks];
  const uint4* src[kDpMaxItems];     // this rank's contribution
  uint4* dst_mc[kDpMaxItems];        // multicast address of this rank's slot in the gathered buffer
  int64_t n16[kDpMaxItems];          // 16-byte units
  int n_items, rank, world;
};

__global__ void __launch_bounds__(kDpThreads) dp_allgather_kernel(const __grid_constant__ DpGatherArgs a) {
  const uint32_t epoch = dp_enter(a.flags, a.rank, a.world);
  const int64_t T = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < a.n_items; ++it) {
    for (int64_t q0 = tid; q0 < a.n16[it]; q0 += 4 * T) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q0 + u * T < a.n16[it]) v[u] = __ldg(a.src[it] + q0 + u * T);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q0 + u * T < a.n16[it])
          multimem_st_f32x4(reinterpret_cast<float*>(a.dst_mc[it] + q0 + u * T),
                            make_float4(__uint_as_float(v[u].x), __uint_as_float(v[u].y), __uint_as_float(v[u].z),
                                        __uint_as_float(v[u].w)));
    }
  }
  dp_exit(a.flags, a.rank, a.world, epoch);
}

}  // namespace ark

using namespace ark;

extern "C" int ark_dp_reduce_adam(float* grad_mc, float* param_mc, uint16_t* shadow_mc, const float* param, float* m,
                                  float* v, uint32_t* const* peer_flags, int rank, int world, const int64_t* span_begin,
                                  const int64_t* span_end, int n_spans, int mode, float lr, float beta1, float beta2,
                                  float eps, int64_t step, const float* hyper, int ctas, void* stream) {
  ARK_REQUIRE(grad_mc && peer_flags && span_begin && span_end, ARK_E_BADARG, "dp_reduce_adam: null pointer");
  ARK_REQUIRE(world >= 2 && world <= kDpMaxRanks && rank >= 0 && rank < world, ARK_E_BADARG,
              "dp_reduce_adam: world %d / rank %d out of range (2..%d ranks)", world, rank, kDpMaxRanks);
  ARK_REQUIRE(n_spans >= 1 && n_spans <= kDpMaxSpans, ARK_E_BADARG, "dp_reduce_adam: 1..%d spans per launch", kDpMaxSpans);
  ARK_REQUIRE(mode == 0 || (param_mc && shadow_mc && param && m && v), ARK_E_BADARG, "dp_reduce_adam: null pointer");
  ARK_REQUIRE(mode == 0 || hyper || step >= 1, ARK_E_BADARG, "dp_reduce_adam: step is 1-based");
  ARK_REQUIRE(aligned16(grad_mc) && aligned16(param_mc) && aligned16(shadow_mc) && aligned16(param) && aligned16(m) &&
                  aligned16(v), ARK_E_ALIGN, "dp_reduce_adam: 16-byte alignment");
  DpArgs a{};
  int64_t work = 0;
  for (int i = 0; i < n_spans; ++i) {
    ARK_REQUIRE(span_begin[i] >= 0 && span_end[i] >= span_begin[i] && span_begin[i] % 4 == 0 && span_end[i] % 4 == 0,
                ARK_E_ALIGN, "dp_reduce_adam: span %d = [%lld, %lld) must be a multiple of 4 elements", i,
                (long long)span_begin[i], (long long)span_end[i]);
    a.s[i] = span_begin[i];
    a.e[i] = span_end[i];
    work = std::max<int64_t>(work, ((span_end[i] - span_begin[i]) / 4 + world - 1) / world);
  }
  for (int r = 0; r < world; ++r) {
    ARK_REQUIRE(peer_flags[r], ARK_E_BADARG, "dp_reduce_adam: null flag block of rank %d", r);
    a.flags[r] = peer_flags[r];
  }
  a.n_spans = n_spans; a.rank = rank; a.world = world; a.mode = mode;
  a.grad_mc = grad_mc; a.param_mc = param_mc; a.shadow_mc = shadow_mc; a.param = param; a.m = m; a.v = v;
  a.hyper = hyper; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.step_size = 0.f; a.inv_sqrt_bc2 = 1.f;
  if (mode != 0 && !hyper) {
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    a.step_size = (float)((double)lr / bc1);
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  static const int unroll = [] {
    const char* e = getenv("ARK_DP_MM_UNROLL");
    return e ? atoi(e) : 4;
  }();
  if (ctas <= 0) ctas = 32;
  const int64_t need = (work + (int64_t)kDpThreads * unroll - 1) / ((int64_t)kDpThreads * unroll);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(ctas, kNumSMs), need));
  if (unroll >= 8)
    dp_reduce_adam_kernel<8><<<grid, kDpThreads, 0, (cudaStream_t)stream>>>(a);
  else if (unroll >= 4)
    dp_reduce_adam_kernel<4><<<grid, kDpThreads, 0, (cudaStream_t)stream>>>(a);
  else
    dp_reduce_adam_kernel<2><<<grid, kDpThreads, 0, (cudaStream_t)stream>>>(a);
  return launched("dp_reduce_adam");
}

extern "C" int ark_dp_allgather_mc(const void* const* src, void* const* dst_mc, const int64_t* nbytes, int n_items,
                                   uint32_t* const* peer_flags, int rank, int world, int ctas, void* stream) {
  ARK_REQUIRE(src && dst_mc && nbytes && peer_flags, ARK_E_BADARG, "dp_allgather_mc: null pointer");
  ARK_REQUIRE(world >= 2 && world <= kDpMaxRanks && rank >= 0 && rank < world, ARK_E_BADARG,
              "dp_allgather_mc: world %d / rank %d out of range (2..%d ranks)", world, rank, kDpMaxRanks);
  ARK_REQUIRE(n_items >= 1 && n_items <= kDpMaxItems, ARK_E_BADARG, "dp_allgather_mc: 1..%d items per launch", kDpMaxItems);
  DpGatherArgs a{};
  int64_t work = 0;
  for (int i = 0; i < n_items; ++i) {
    ARK_REQUIRE(src[i] && dst_mc[i] && aligned16(src[i]) && aligned16(dst_mc[i]) && nbytes[i] >= 0 && nbytes[i] % 16 == 0,
                ARK_E_ALIGN, "dp_allgather_mc: item %d must be 16-byte aligned and a multiple of 16 bytes", i);
    a.src[i] = static_cast<const uint4*>(src[i]);
    a.dst_mc[i] = static_cast<uint4*>(dst_mc[i]);
    a.n16[i] = nbytes[i] / 16;
    work = std::max(work, a.n16[i]);
  }
  for (int r = 0; r < world; ++r) {
    ARK_REQUIRE(peer_flags[r], ARK_E_BADARG, "dp_allgather_mc: null flag block of rank %d", r);
    a.flags[r] = peer_flags[r];
  }
  a.n_items = n_items; a.rank = rank; a.world = world;
  if (ctas <= 0) ctas = 16;
  const int64_t need = (work + (int64_t)kDpThreads * 4 - 1) / ((int64_t)kDpThreads * 4);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(ctas, kNumSMs), need));
  dp_allgather_kernel<<<grid, kDpThreads, 0, (cudaStream_t)stream>>>(a);
  return launched("dp_allgather_mc");
}
