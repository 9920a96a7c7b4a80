// K6 (persistent): one GRU layer through ALL time steps in ONE cooperative launch, forward and backward.
//
// Replaces nn.GRU's recurrence (kgvae/model/models.py:121-127,141) — in the reference a cuDNN/ATen loop of
// L dependent steps per layer; in the first version of this library 2 launches per step (launch-bound:
// 5.6-14 us/step measured).  Here the L steps run inside one kernel:
//
//   grid = (d/DJ hidden slices) x (ceil(B/128) batch tiles), one CTA per SM, all co-resident.
//   CTA (ji, bi) keeps ITS slice of W_hh resident in shared memory for the whole sequence
//     forward : rows {g*d + j0 .. j0+DJ} (g = r,z,n) of W_hh[3d,d]    -> UMMA B operand [3*DJ x d], K = d
//     backward: rows {j0 .. j0+DJ} of W_hh^T[d,3d]                     -> UMMA B operand [DJ x 3d],  K = 3d
//   per step it streams the A operand (h_{t-1} rows, resp. dgh_{t+1} rows: bf16, written to HBM/L2 by all
//   CTAs of the batch tile in the previous step) through a TMA ring, accumulates in TMEM, and runs the gate
//   math in the epilogue warps straight out of TMEM.  The recurrent state of the CTA's own (row, slice)
//   elements never leaves registers.  CTAs of one batch tile synchronise through ONE release/acquire counter
//   per tile in global memory (different batch tiles never wait for each other).
//
// Rows are the packed, length-sorted layout of ark_b200/layout.py: step t owns rows [off[t], off[t]+bt[t]),
// bt non-increasing, so a thread always serves the same graph and inactive tiles simply leave the loop.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "gru_math.cuh"
#include "gru_dev.cuh"
#include "philox.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
#include <vector>

namespace ark {

constexpr int GP_BM = 128;
constexpr int GP_BK = 64;
constexpr int GP_A_BYTES = GP_BM * GP_BK * 2;  // 16 KB per ring stage
// 8 epilogue warps: warps 2..5 drain the 4 TMEM lane quadrants into shared memory, then ALL 256 threads share the
// (row, 4-unit group) work items of the gate math (measured at d = 1024: the math + stores of 4 items per thread took
// 3 400 of the 14 400 cycles of a forward step with 128 threads)
constexpr int GP_EPI = 256;
constexpr int GP_THREADS = 64 + GP_EPI;
__device__ __forceinline__ void gp_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct GruPersistFwdParams {
  const int32_t* bt;
  const int32_t* off;
  int L, d;
  int32_t* sync;       // [n batch tiles], zeroed before launch
  const float* gi;     // [N, 3d] = W_ih x + b_ih
  const float* b_hh;   // [3d]
  const float* h0;     // [bt[0], d] fp32 initial state
  uint16_t* hp_b;      // [N, d] bf16 packed h_prev rows (block 0 pre-filled with bf16(h0)); written for t+1
  uint16_t* y_b;       // [N, d] bf16 outputs
 uint16_t *r, *z, *n, *ghn;  // [N, d] bf16 saved gates (may all be null)
  // fused inter-layer dropout of the OUTPUT rows y_b (hp_b keeps the undropped state): same Philox draw as
  // ark_dropout_bf16 over the [N, d] tensor (counter = offset + element/4); mask u8 [N, d] or null
  uint8_t* mask;
  const uint64_t* offset_dev;
  uint64_t seed, offset;
  float p_drop;
  long long* dbg;      // ARK_GRU_PERSIST_DBG: clock64 timeline of CTA (0,0), steps 2..5 ([4][8] words), else NULL
};

struct GruPersistBwdParams {
  const int32_t* bt;
  const int32_t* off;
  int L, d;
  int32_t* sync;
  const float* dy;                          // [N, d] gradient w.r.t. the layer outputs
  const uint16_t *r, *z, *n, *ghn, *hp_b;   // saved by the forward kernel
  uint16_t *dgi_b, *dgh_b;                  // [N, 3d] bf16 (dgh_b is also the A operand of the next step)
 float* dh0;                               // [bt[0], d]
  int dh0_accumulate;
  const uint8_t* dy_mask;                   // keep mask of the dropout applied to this layer's OUTPUT in the forward pass
  float dy_scale;                           // (dy is multiplied by keep / (1 - p) on the fly), or null
  long long* dbg;                           // as in the forward parameters ([4][8] words behind the forward's)
};

// debug timeline: event e of step t (CTA (0,0) only, 4 steps starting at step 2 of the kernel's own iteration order)
__device__ __forceinline__ void gp_dbg(long long* dbg, int it, int e) {
  if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && it >= 2 && it < 6) dbg[(it - 2) * 8 + e] = clock64();
}

template <int DJ, int STAGES>
struct GpSmem {
  static constexpr int W_BYTES(int d) { return 3 * DJ * d * 2; }
  static constexpr int total(int d) { return W_BYTES(d) + STAGES * GP_A_BYTES + (2 * STAGES + 2) * 8 + 16 + 1024; }
};

// =====================================================================================================
// forward
// =====================================================================================================
template <int DJ, int STAGES, int CS>
__global__ void __launch_bounds__(GP_THREADS, 1) gru_persist_fwd_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmW,
                                                                 const GruPersistFwdParams p) {
  constexpr int NROWS = 3 * DJ;  // UMMA N
  constexpr uint32_t TMEM_COLS = NROWS <= 32 ? 32 : (NROWS <= 64 ? 64 : (NROWS <= 128 ? 128 : 256));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L;
  const int nkc = d / GP_BK;
  uint8_t* w_sm = smem;
  uint8_t* a_sm = smem + 3 * DJ * d * 2;
  float* acc_sm = reinterpret_cast<float*>(a_sm + STAGES * GP_A_BYTES);   // [128][ACC_LD] TMEM -> smem staging
  constexpr int ACC_LD = NROWS + 1;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(acc_sm + GP_BM * ACC_LD + 1);
  full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(full_bar) + 7) & ~(uintptr_t)7);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* w_bar = empty_bar + STAGES;
  uint64_t* tmem_full_bar = w_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ji = blockIdx.x, bi = blockIdx.y, ns = gridDim.x;
  const int j0 = ji * DJ, m0 = bi * GP_BM;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], CS);     // every CTA of the cluster releases the stage (multicast A tile)
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();       // peers' mbarriers exist before any multicast / remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr int MC_ROWS = GP_BM / CS;        // rows of the A tile this CTA fetches (and multicasts to its cluster)
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CS) - 1u);
  const uint32_t crank = CS > 1 ? ptx::cluster_ctarank() : 0u;

  if (warp == 0) {
    if (ptx::elect_one()) {
      // resident weights: nkc chunks of [3*DJ rows x 64 k] (gate-major rows), 128B-swizzled
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(3 * DJ * d * 2));
      for (int kc = 0; kc < nkc; ++kc)
        for (int g = 0; g < 3; ++g)
          ptx::tma_load_2d(w_sm + kc * (NROWS * 128) + g * (DJ * 128), &tmW, w_bar, kc * GP_BK, g * d + j0);
      int it = 0;
      for (int t = 0; t < L; ++t) {
        if (m0 >= p.bt[t]) break;
        if (t > 0) wait_counter(p.sync + bi, t * ns);   // every slice of h_{t-1} is in global memory
        gp_dbg(p.dbg, t, 0);
        asm volatile("fence.proxy.async;" ::: "memory");
        const int row0 = p.off[t] + m0;
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % STAGES;
          ptx::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[s], GP_A_BYTES);
          if (CS == 1) ptx::tma_load_2d(a_sm + s * GP_A_BYTES, &tmA, &full_bar[s], kc * GP_BK, row0);
          else ptx::tma_load_2d_mc(a_sm + s * GP_A_BYTES + crank * (MC_ROWS * 128), &tmA, &full_bar[s], kc * GP_BK,
                                   row0 + (int)crank * MC_ROWS, MC_MASK);
        }
        gp_dbg(p.dbg, t, 1);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(GP_BM, NROWS, 0, 0);
      ptx::mbar_wait(w_bar, 0);
      const uint32_t w_addr = ptx::smem_u32(w_sm), a_addr0 = ptx::smem_u32(a_sm);
      int it = 0;
      for (int t = 0; t < L; ++t) {
        if (m0 >= p.bt[t]) break;
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % STAGES;
          ptx::mbar_wait(&full_bar[s], (it / STAGES) & 1);
          if (kc == 0) gp_dbg(p.dbg, t, 2);
          ptx::tc_fence_after();
{   // descriptors built once per chunk, a k-step is an ADD on the (address >> 4) field (see gemm_tc.cu)
            const uint64_t ad = ptx::make_smem_desc_sw128(a_addr0 + s * GP_A_BYTES, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc_sw128(w_addr + kc * (NROWS * 128), 16, 1024);
            ptx::umma_f16(tmem_base, ad, bd, idesc, kc != 0 ? 1u : 0u);
#pragma unroll
            for (int kk = 1; kk < GP_BK / 16; ++kk) ptx::umma_f16_acc(tmem_base, ad + kk * 2, bd + kk * 2, idesc);
          }
          if (CS == 1) ptx::umma_commit(&empty_bar[s]);
          else ptx::umma_commit_mc(&empty_bar[s], MC_MASK);
        }
        ptx::umma_commit(tmem_full_bar);
        gp_dbg(p.dbg, t, 3);
      }
    }
  } else {
    // ===================== epilogue =====================
    // (1) each warp drains its 32 TMEM lanes (batch rows) into smem; (2) the 128 threads then share the
    // (row, 4-unit group) work items evenly — with 16 live rows that is 64 busy threads instead of 16 —
    // with 16-byte gi loads and 8-byte bf16 stores.  Work item e = tid + 128*i is FIXED over time, so the
    // recurrent state of its 4 units stays in registers.  gi for step t is fetched BEFORE waiting for the
    // MMA of step t (it does not depend on h).
    constexpr int G = DJ / 4;                    // groups per row
    constexpr int IT = G * GP_BM / GP_EPI;       // work items per thread
    const int q = warp & 3;
    const bool drainer = warp < 6;               // warps 2..5 own the TMEM lane quadrants 2,3,0,1
    const int tid = threadIdx.x - 64;            // 0..255
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t d3 = 3 * (int64_t)d;
    float hreg[IT][4];
    const int bt0 = p.bt[0];
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const int e = tid + GP_EPI * i, bl = e / G, j = j0 + (e % G) * 4;
      const int b = m0 + bl;
      const float4 hv = (b < bt0) ? *reinterpret_cast<const float4*>(p.h0 + (int64_t)b * d + j) : make_float4(0, 0, 0, 0);
      hreg[i][0] = hv.x; hreg[i][1] = hv.y; hreg[i][2] = hv.z; hreg[i][3] = hv.w;
    }
    const bool drop = p.p_drop > 0.f;
    const float drop_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const uint64_t ctr0 = p.offset + (p.offset_dev ? *p.offset_dev : 0ull);
    float4 gpre[IT][3];
    auto prefetch_gi = [&](int t) {
      const int Bt = p.bt[t];
      const int64_t base = (int64_t)p.off[t] + m0;
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int e = tid + GP_EPI * i, bl = e / G, j = j0 + (e % G) * 4;
        if (m0 + bl < Bt) {
          const float* gp = p.gi + (base + bl) * d3 + j;
          gpre[i][0] = *reinterpret_cast<const float4*>(gp);
          gpre[i][1] = *reinterpret_cast<const float4*>(gp + d);
          gpre[i][2] = *reinterpret_cast<const float4*>(gp + 2 * d);
        }
      }
    };
    if (m0 < bt0) prefetch_gi(0);
    for (int t = 0; t < L; ++t) {
      const int Bt = p.bt[t];
      if (m0 >= Bt) break;
      const int Bn = (t + 1 < L) ? p.bt[t + 1] : 0;
      const int64_t base = (int64_t)p.off[t] + m0;
      const int64_t base_n = (t + 1 < L) ? (int64_t)p.off[t + 1] + m0 : 0;
      ptx::mbar_wait(tmem_full_bar, t & 1);
      if (tid == 0) gp_dbg(p.dbg, t, 4);
      ptx::tc_fence_after();
      if (drainer && m0 + q * 32 < Bt) {          // warp-uniform: skip lane quadrants without live rows
        float* dst = acc_sm + (q * 32 + lane) * ACC_LD;
#pragma unroll
        for (int c = 0; c < NROWS; c += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)c, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) dst[c + k] = __uint_as_float(v[k]);
        }
      }
      ptx::tc_fence_before();
      gp_bar_sync();
      if (tid == 0) gp_dbg(p.dbg, t, 5);
      float o_r[IT][4], o_z[IT][4], o_n[IT][4], o_g[IT][4];
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int e = tid + GP_EPI * i, bl = e / G, jl = (e % G) * 4;
        const int b = m0 + bl;
        if (b < Bt) {
          const float* ap = acc_sm + bl * ACC_LD + jl;
          const float gr[4] = {gpre[i][0].x, gpre[i][0].y, gpre[i][0].z, gpre[i][0].w};
          const float gz[4] = {gpre[i][1].x, gpre[i][1].y, gpre[i][1].z, gpre[i][1].w};
          const float gn[4] = {gpre[i][2].x, gpre[i][2].y, gpre[i][2].z, gpre[i][2].w};
          const float4 b_r = __ldg(reinterpret_cast<const float4*>(p.b_hh + j0 + jl));
          const float4 b_z = __ldg(reinterpret_cast<const float4*>(p.b_hh + d + j0 + jl));
          const float4 b_n = __ldg(reinterpret_cast<const float4*>(p.b_hh + 2 * d + j0 + jl));
          const float br[4] = {b_r.x, b_r.y, b_r.z, b_r.w}, bz[4] = {b_z.x, b_z.y, b_z.z, b_z.w};
          const float bn[4] = {b_n.x, b_n.y, b_n.z, b_n.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const GruFwd o = gru_fwd_math_fast(gr[k], gz[k], gn[k], ap[k] + br[k], ap[DJ + k] + bz[k],
                                               ap[2 * DJ + k] + bn[k], hreg[i][k]);
            hreg[i][k] = o.h;
            o_r[i][k] = o.r; o_z[i][k] = o.z; o_n[i][k] = o.n; o_g[i][k] = o.ghn;
          }
          // ONLY the h_prev rows of step t+1 sit on the recurrent chain: they go out before the release; the layer
          // output and the saved gates (read by later kernels) are stored after it, off the chain
          if (b < Bn) st4_bf16(p.hp_b + (base_n + bl) * d + j0 + jl, hreg[i]);
        }
      }
      gp_bar_sync();                             // the chain's stores are issued (and acc_sm is free again)
      if (tid == 0) gp_dbg(p.dbg, t, 6);
      if (tid == 0) red_release_add(p.sync + bi, 1);   // release: cumulative over the barrier above
      if (tid == 0) gp_dbg(p.dbg, t, 7);
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int e = tid + GP_EPI * i, bl = e / G, jl = (e % G) * 4;
        if (m0 + bl < Bt) {
          const int64_t o = (base + bl) * d + j0 + jl;
          if (drop) {   // same draw as dropout_bf16_kernel over the [N, d] output of this layer
            const uint64_t c = ctr0 + (uint64_t)(o >> 2);
            const uint4 rn = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
            const uint32_t rr[4] = {rn.x, rn.y, rn.z, rn.w};
            float yo[4];
            uint32_t mk = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const bool keep = (float)(rr[k] >> 8) * (1.f / 16777216.f) >= p.p_drop;
              mk |= (keep ? 1u : 0u) << (8 * k);
              yo[k] = keep ? bf16_bits_to_f32(f32_to_bf16_bits(hreg[i][k])) * drop_scale : 0.f;
            }
            if (p.mask) *reinterpret_cast<uint32_t*>(p.mask + o) = mk;
            st4_bf16(p.y_b + o, yo);
          } else {
            st4_bf16(p.y_b + o, hreg[i]);
          }
          if (p.r) {
            st4_bf16(p.r + o, o_r[i]);
            st4_bf16(p.z + o, o_z[i]);
            st4_bf16(p.n + o, o_n[i]);
            st4_bf16(p.ghn + o, o_g[i]);
          }
        }
      }
      if (t + 1 < L && m0 < Bn) prefetch_gi(t + 1);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();       // no CTA leaves while a peer may still multicast into it
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// =====================================================================================================
// backward through time
// =====================================================================================================
template <int DJ, int STAGES, int CS>
__global__ void __launch_bounds__(GP_THREADS, 1) gru_persist_bwd_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmW,
                                                                 const GruPersistBwdParams p) {
  constexpr int NROWS = DJ;  // UMMA N
  constexpr uint32_t TMEM_COLS = NROWS <= 32 ? 32 : (NROWS <= 64 ? 64 : 128);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L;
  const int nkc = 3 * d / GP_BK;
  uint8_t* w_sm = smem;
  uint8_t* a_sm = smem + 3 * DJ * d * 2;
  float* acc_sm = reinterpret_cast<float*>(a_sm + STAGES * GP_A_BYTES);   // [128][ACC_LD]
  constexpr int ACC_LD = NROWS + 1;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(acc_sm + GP_BM * ACC_LD + 1);
  full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(full_bar) + 7) & ~(uintptr_t)7);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* w_bar = empty_bar + STAGES;
  uint64_t* tmem_full_bar = w_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ji = blockIdx.x, bi = blockIdx.y, ns = gridDim.x;
  const int j0 = ji * DJ, m0 = bi * GP_BM;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], CS);     // every CTA of the cluster releases the stage (multicast A tile)
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();       // peers' mbarriers exist before any multicast / remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr int MC_ROWS = GP_BM / CS;        // rows of the A tile this CTA fetches (and multicasts to its cluster)
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CS) - 1u);
  const uint32_t crank = CS > 1 ? ptx::cluster_ctarank() : 0u;

  // Iterations t = L-1 .. 0 (cell backward of step t) and t = -1 (gradient of the initial state).
  // Iteration t consumes dgh_{t+1} (if step t+1 had rows in this tile) through the tensor cores.
  auto tile_active = [&](int t) { return m0 < p.bt[t < 0 ? 0 : t]; };
  auto has_mma = [&](int t) { return (t + 1 <= L - 1) && (m0 < p.bt[t + 1]); };

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(3 * DJ * d * 2));
      for (int kc = 0; kc < nkc; ++kc)
        ptx::tma_load_2d(w_sm + kc * (NROWS * 128), &tmW, w_bar, kc * GP_BK, j0);
      int it = 0, done = 0;  // done = iterations this tile has completed
      for (int t = L - 1; t >= -1; --t) {
        if (!tile_active(t)) continue;
        if (has_mma(t)) {
          wait_counter(p.sync + bi, done * ns);   // every slice of dgh_{t+1} is in global memory
          gp_dbg(p.dbg, done, 0);
          asm volatile("fence.proxy.async;" ::: "memory");
          const int row0 = p.off[t + 1] + m0;
          for (int kc = 0; kc < nkc; ++kc, ++it) {
            const int s = it % STAGES;
            ptx::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[s], GP_A_BYTES);
            if (CS == 1) ptx::tma_load_2d(a_sm + s * GP_A_BYTES, &tmA, &full_bar[s], kc * GP_BK, row0);
            else ptx::tma_load_2d_mc(a_sm + s * GP_A_BYTES + crank * (MC_ROWS * 128), &tmA, &full_bar[s], kc * GP_BK,
                                     row0 + (int)crank * MC_ROWS, MC_MASK);
          }
          gp_dbg(p.dbg, done, 1);
        }
        ++done;
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(GP_BM, NROWS, 0, 0);
      ptx::mbar_wait(w_bar, 0);
      const uint32_t w_addr = ptx::smem_u32(w_sm), a_addr0 = ptx::smem_u32(a_sm);
      int it = 0, nm = 0;
      for (int t = L - 1; t >= -1; --t) {
        if (!tile_active(t) || !has_mma(t)) continue;
        ++nm;
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % STAGES;
          ptx::mbar_wait(&full_bar[s], (it / STAGES) & 1);
          if (kc == 0) gp_dbg(p.dbg, nm, 2);
          ptx::tc_fence_after();
{   // descriptors built once per chunk, a k-step is an ADD on the (address >> 4) field (see gemm_tc.cu)
            const uint64_t ad = ptx::make_smem_desc_sw128(a_addr0 + s * GP_A_BYTES, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc_sw128(w_addr + kc * (NROWS * 128), 16, 1024);
            ptx::umma_f16(tmem_base, ad, bd, idesc, kc != 0 ? 1u : 0u);
#pragma unroll
            for (int kk = 1; kk < GP_BK / 16; ++kk) ptx::umma_f16_acc(tmem_base, ad + kk * 2, bd + kk * 2, idesc);
          }
          if (CS == 1) ptx::umma_commit(&empty_bar[s]);
          else ptx::umma_commit_mc(&empty_bar[s], MC_MASK);
        }
        ptx::umma_commit(tmem_full_bar);
        gp_dbg(p.dbg, nm, 3);
      }
    }
  } else {
    // same two-phase epilogue as the forward kernel: TMEM -> smem, then (row, 4-unit group) work items
    // spread over the 128 threads; dy and the saved gates of step t are fetched BEFORE the MMA wait.
    constexpr int G = DJ / 4;
    constexpr int IT = G * GP_BM / GP_EPI;       // work items per thread
    const int q = warp & 3;
    const bool drainer = warp < 6;               // warps 2..5 own the TMEM lane quadrants 2,3,0,1
    const int tid = threadIdx.x - 64;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t d3 = 3 * (int64_t)d;
    float carry[IT][4];   // dh_{t+1} * z_{t+1}: the direct path into h_t (valid for rows of step t+1)
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) carry[i][k] = 0.f;
    float4 dyp[IT];
    uint2 sp[IT][5];
    auto prefetch = [&](int t) {
      const int Bt = p.bt[t];
      const int64_t base = (int64_t)p.off[t] + m0;
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int e = tid + GP_EPI * i, bl = e / G, jl = (e % G) * 4;
        if (m0 + bl < Bt) {
          const int64_t o = (base + bl) * d + j0 + jl;
          dyp[i] = *reinterpret_cast<const float4*>(p.dy + o);
          if (p.dy_mask) {     // backward of the inter-layer dropout applied to this layer's output
            const uint32_t mk = *reinterpret_cast<const uint32_t*>(p.dy_mask + o);
            dyp[i].x = (mk & 0xFFu) ? dyp[i].x * p.dy_scale : 0.f;
            dyp[i].y = (mk & 0xFF00u) ? dyp[i].y * p.dy_scale : 0.f;
            dyp[i].z = (mk & 0xFF0000u) ? dyp[i].z * p.dy_scale : 0.f;
            dyp[i].w = (mk & 0xFF000000u) ? dyp[i].w * p.dy_scale : 0.f;
          }
          sp[i][0] = *reinterpret_cast<const uint2*>(p.r + o);
          sp[i][1] = *reinterpret_cast<const uint2*>(p.z + o);
          sp[i][2] = *reinterpret_cast<const uint2*>(p.n + o);
          sp[i][3] = *reinterpret_cast<const uint2*>(p.ghn + o);
          sp[i][4] = *reinterpret_cast<const uint2*>(p.hp_b + o);
        }
      }
    };
    int t_first = L - 1;
    while (t_first >= 0 && !tile_active(t_first)) --t_first;
    if (t_first >= 0) prefetch(t_first);
    int n_mma = 0;
    for (int t = t_first; t >= -1; --t) {
      const bool mma = has_mma(t);
      const int B_next = (t + 1 <= L - 1) ? p.bt[t + 1] : 0;
      const int Bt = p.bt[t < 0 ? 0 : t];
      if (mma) {
        ptx::mbar_wait(tmem_full_bar, n_mma & 1);
        ptx::tc_fence_after();
        ++n_mma;
        if (tid == 0) gp_dbg(p.dbg, n_mma, 4);
        if (drainer && m0 + q * 32 < B_next) {
          float* dst = acc_sm + (q * 32 + lane) * ACC_LD;
#pragma unroll
          for (int c = 0; c < NROWS; c += 16) {
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)c, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[c + k] = __uint_as_float(v[k]);
          }
        }
        ptx::tc_fence_before();
      }
      gp_bar_sync();
      if (tid == 0) gp_dbg(p.dbg, n_mma, 5);
      const int64_t base = (t >= 0) ? (int64_t)p.off[t] + m0 : 0;
      float dar[IT][4], daz[IT][4], dan[IT][4];
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int e = tid + GP_EPI * i, bl = e / G, jl = (e % G) * 4;
        const int b = m0 + bl;
        if (b >= Bt) continue;
        const bool from_next = b < B_next;
        float dh[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          dh[k] = from_next ? carry[i][k] + (mma ? acc_sm[bl * ACC_LD + jl + k] : 0.f) : 0.f;
        if (t < 0) {
          float* o = p.dh0 + (int64_t)b * d + j0 + jl;
          float4 v = make_float4(dh[0], dh[1], dh[2], dh[3]);
          if (p.dh0_accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(o);
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          *reinterpret_cast<float4*>(o) = v;
          continue;
        }
        const float dyv[4] = {dyp[i].x, dyp[i].y, dyp[i].z, dyp[i].w};
        float r[4], z[4], n[4], g[4], hp[4];
        {
          float2 a, c2;
          a = unpack_bf16x2(sp[i][0].x); c2 = unpack_bf16x2(sp[i][0].y); r[0] = a.x; r[1] = a.y; r[2] = c2.x; r[3] = c2.y;
          a = unpack_bf16x2(sp[i][1].x); c2 = unpack_bf16x2(sp[i][1].y); z[0] = a.x; z[1] = a.y; z[2] = c2.x; z[3] = c2.y;
          a = unpack_bf16x2(sp[i][2].x); c2 = unpack_bf16x2(sp[i][2].y); n[0] = a.x; n[1] = a.y; n[2] = c2.x; n[3] = c2.y;
          a = unpack_bf16x2(sp[i][3].x); c2 = unpack_bf16x2(sp[i][3].y); g[0] = a.x; g[1] = a.y; g[2] = c2.x; g[3] = c2.y;
          a = unpack_bf16x2(sp[i][4].x); c2 = unpack_bf16x2(sp[i][4].y); hp[0] = a.x; hp[1] = a.y; hp[2] = c2.x; hp[3] = c2.y;
        }
        float danr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const GruBwd w = gru_bwd_math(dh[k] + dyv[k], r[k], z[k], n[k], g[k], hp[k]);
          dar[i][k] = w.dar; daz[i][k] = w.daz; dan[i][k] = w.dan; danr[k] = w.dan_r;
          carry[i][k] = w.dh_prev;
        }
        // dgh_t is the next iteration's A operand: on the chain, stored before the release; dgi_t (read by the
        // weight-gradient GEMMs after the kernel) is stored after it
        const int64_t o3 = (base + bl) * d3 + j0 + jl;
        st4_bf16(p.dgh_b + o3, dar[i]);
        st4_bf16(p.dgh_b + o3 + d, daz[i]);
        st4_bf16(p.dgh_b + o3 + 2 * d, danr);
      }
      gp_bar_sync();
      if (tid == 0) gp_dbg(p.dbg, n_mma, 6);
      if (tid == 0) red_release_add(p.sync + bi, 1);
      if (tid == 0) gp_dbg(p.dbg, n_mma, 7);
      if (t >= 0) {
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          const int e = tid + GP_EPI * i, bl = e / G, jl = (e % G) * 4;
          if (m0 + bl < Bt) {
            const int64_t o3 = (base + bl) * d3 + j0 + jl;
            st4_bf16(p.dgi_b + o3, dar[i]);
            st4_bf16(p.dgi_b + o3 + d, daz[i]);
            st4_bf16(p.dgi_b + o3 + 2 * d, dan[i]);
          }
        }
      }
      if (t - 1 >= 0) prefetch(t - 1);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();       // no CTA leaves while a peer may still multicast into it
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}


// =====================================================================================================
// backward through time, K-SPLIT variant (d = 1024-class layers: syn-types / syn-tipr)
// =====================================================================================================
// Measured on the kernel above at d = 1024, B = 256 (tools/gru_persist_bench.py, clock64 timeline): of the 25 300
// cycles of a step, 18 700 are the A-operand fetch — every CTA pulls the WHOLE dgh_{t+1} tile ([128 x 3d] bf16 =
// 768 KB) through TMA at ~43 B/clk per SM, whatever the ring depth or multicast width.  The bytes have to go: here a
// CLUSTER of KS = 4 CTAs owns 64 output units together; CTA r multiplies only ITS QUARTER of the contraction
// (gate columns [r*3d/4, (r+1)*3d/4): 192 KB of dgh per step instead of 768 KB, N = 64 MMAs instead of N = 16) and
// the four [128 x 64] fp32 partial sums are reduce-scattered through distributed shared memory: the [128 x 16] block
// that belongs to peer q goes there with ONE 8 KB bulk copy that completes on the peer's mbarrier (24 KB out, 24 KB
// in per CTA per step).  Each CTA then runs the unchanged gate math on its own 16 units.
//   resident weights: rows [64 units of the cluster] x K quarter of W_hh^T  = 64 * (3d/4) * 2 B  (96 KB at d = 1024)
constexpr int KS = 4;            // K splits = cluster width
constexpr int KS_NC = 64;        // output units per cluster = UMMA N
constexpr int KS_DJ = 16;        // output units per CTA
constexpr int KS_BLK = GP_BM * KS_DJ * 4;   // one [128 x 16] fp32 partial block: 8 KB

// 16-byte chunk c (0..3) of row `row` inside a partial block: XOR-swizzled so that a warp's float4 row writes and the
// (row, 4-unit) item reads are both bank-conflict free
__device__ __forceinline__ int ks_chunk_off(int row, int c) { return row * KS_DJ + ((c ^ ((row >> 1) & 3)) << 2); }

template <int STAGES>
__global__ void __launch_bounds__(GP_THREADS, 1) gru_persist_bwd_ks_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmW,
                                                                    const GruPersistBwdParams p) {
  constexpr int DJ = KS_DJ;
  constexpr uint32_t TMEM_COLS = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L;
  const int kq = 3 * d / KS;                 // contraction elements of this CTA
  const int nkc = kq / GP_BK;
  uint8_t* w_sm = smem;                                            // [nkc][64 rows x 128 B]
  uint8_t* a_sm = w_sm + nkc * (KS_NC * 128);                      // ring
  float* send_sm = reinterpret_cast<float*>(a_sm + STAGES * GP_A_BYTES);   // [KS dst][128 x 16] (own block stays here)
  float* part_sm = send_sm + KS * (KS_BLK / 4);                    // [KS src][128 x 16] (slot of the own rank unused)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(part_sm + KS * (KS_BLK / 4));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* w_bar = empty_bar + STAGES;
  uint64_t* tmem_full_bar = w_bar + 1;
  uint64_t* part_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(part_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bi = blockIdx.y, ns = gridDim.x;
  const uint32_t crank = ptx::cluster_ctarank();               // == blockIdx.x % KS (cluster = KS consecutive x)
  const int jc0 = ((int)blockIdx.x / KS) * KS_NC;              // first output unit of the cluster
  const int j0 = jc0 + (int)crank * DJ;                        // first output unit of this CTA
  const int k0 = (int)crank * kq;                              // first contraction element of this CTA
  const int m0 = bi * GP_BM;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::mbar_init(part_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                   // the peers' part_bar exists before anybody copies into them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto tile_active = [&](int t) { return m0 < p.bt[t < 0 ? 0 : t]; };
  auto has_mma = [&](int t) { return (t + 1 <= L - 1) && (m0 < p.bt[t + 1]); };

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(nkc * KS_NC * 128));
      // W_hh itself ([3d gate rows, d units], units contiguous) is the B operand in MN-major form: chunk kc = 64 gate
      // rows x 64 units (128 B rows, 128B swizzle) — no transposed copy of the weights is ever made
      for (int kc = 0; kc < nkc; ++kc)
        ptx::tma_load_2d(w_sm + kc * (KS_NC * 128), &tmW, w_bar, jc0, k0 + kc * GP_BK);
      int it = 0, done = 0;
      for (int t = L - 1; t >= -1; --t) {
        if (!tile_active(t)) continue;
        if (has_mma(t)) {
          wait_counter(p.sync + bi, done * ns);   // every slice of dgh_{t+1} is in global memory
          gp_dbg(p.dbg, done, 0);
          asm volatile("fence.proxy.async;" ::: "memory");
          const int row0 = p.off[t + 1] + m0;
          for (int kc = 0; kc < nkc; ++kc, ++it) {
            const int s = it % STAGES;
            ptx::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[s], GP_A_BYTES);
            ptx::tma_load_2d(a_sm + s * GP_A_BYTES, &tmA, &full_bar[s], k0 + kc * GP_BK, row0);
          }
          gp_dbg(p.dbg, done, 1);
        }
        ++done;
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(GP_BM, KS_NC, 0, 1);     // A K-major, B MN-major
      ptx::mbar_wait(w_bar, 0);
      const uint32_t w_addr = ptx::smem_u32(w_sm), a_addr0 = ptx::smem_u32(a_sm);
      int it = 0, nm = 0;
      for (int t = L - 1; t >= -1; --t) {
        if (!tile_active(t) || !has_mma(t)) continue;
        ++nm;
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % STAGES;
          ptx::mbar_wait(&full_bar[s], (it / STAGES) & 1);
          if (kc == 0) gp_dbg(p.dbg, nm, 2);
          ptx::tc_fence_after();
{   // descriptors built once per chunk, a k-step is an ADD on the (address >> 4) field (see gemm_tc.cu)
            const uint64_t ad = ptx::make_smem_desc_sw128(a_addr0 + s * GP_A_BYTES, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc_sw128(w_addr + kc * (KS_NC * 128), GP_BK * 128, 1024);
            ptx::umma_f16(tmem_base, ad, bd, idesc, kc != 0 ? 1u : 0u);
#pragma unroll
            for (int kk = 1; kk < GP_BK / 16; ++kk) ptx::umma_f16_acc(tmem_base, ad + kk * 2, bd + kk * 128, idesc);
          }
          ptx::umma_commit(&empty_bar[s]);
        }
        ptx::umma_commit(tmem_full_bar);
        gp_dbg(p.dbg, nm, 3);
      }
    }
  } else {
    constexpr int G = DJ / 4;                    // 4-unit groups per row
    constexpr int IT = G * GP_BM / GP_EPI;       // work items per thread
    const int q = warp & 3;
    const bool drainer = warp < 6;               // warps 2..5 own the TMEM lane quadrants 2,3,0,1
    const int tid = threadIdx.x - 64;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t d3 = 3 * (int64_t)d;
    float carry[IT][4];
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) carry[i][k] = 0.f;
    float4 dyp[IT];
    uint2 sp[IT][5];
    auto prefetch = [&](int t) {
      const int Bt = p.bt[t];
      const int64_t base = (int64_t)p.off[t] + m0;
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int e = tid + GP_EPI * i, bl = e / G, jl = (e % G) * 4;
        if (m0 + bl < Bt) {
          const int64_t o = (base + bl) * d + j0 + jl;
          dyp[i] = *reinterpret_cast<const float4*>(p.dy + o);
          if (p.dy_mask) {     // backward of the inter-layer dropout applied to this layer's output
            const uint32_t mk = *reinterpret_cast<const uint32_t*>(p.dy_mask + o);
            dyp[i].x = (mk & 0xFFu) ? dyp[i].x * p.dy_scale : 0.f;
            dyp[i].y = (mk & 0xFF00u) ? dyp[i].y * p.dy_scale : 0.f;
            dyp[i].z = (mk & 0xFF0000u) ? dyp[i].z * p.dy_scale : 0.f;
            dyp[i].w = (mk & 0xFF000000u) ? dyp[i].w * p.dy_scale : 0.f;
          }
          sp[i][0] = *reinterpret_cast<const uint2*>(p.r + o);
          sp[i][1] = *reinterpret_cast<const uint2*>(p.z + o);
          sp[i][2] = *reinterpret_cast<const uint2*>(p.n + o);
          sp[i][3] = *reinterpret_cast<const uint2*>(p.ghn + o);
          sp[i][4] = *reinterpret_cast<const uint2*>(p.hp_b + o);
        }
      }
    };
    int t_first = L - 1;
    while (t_first >= 0 && !tile_active(t_first)) --t_first;
    if (t_first >= 0) prefetch(t_first);
    int n_mma = 0;
    const uint32_t part_bar_a = ptx::smem_u32(part_bar);
    for (int t = t_first; t >= -1; --t) {
      const bool mma = has_mma(t);
      const int B_next = (t + 1 <= L - 1) ? p.bt[t + 1] : 0;
      const int Bt = p.bt[t < 0 ? 0 : t];
      if (mma) {
        ptx::mbar_wait(tmem_full_bar, n_mma & 1);
        ptx::tc_fence_after();
        ++n_mma;
        if (tid == 0) gp_dbg(p.dbg, n_mma, 4);
        if (drainer && m0 + q * 32 < B_next) {   // warp-uniform: quadrants without live rows have nothing to send
          const int row = q * 32 + lane;
#pragma unroll
          for (int dst = 0; dst < KS; ++dst) {   // columns [16 dst, 16 dst + 16) of the cluster belong to peer `dst`
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(dst * DJ), v);
            ptx::tmem_ld_wait();
            float* blk = send_sm + dst * (KS_BLK / 4);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(blk + ks_chunk_off(row, c)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
        }
        ptx::tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the bulk copies below read what we just wrote
        gp_bar_sync();
        if (tid == 0) {
          // incoming: the three peers' blocks for MY units (8 KB each) complete on part_bar
          ptx::mbar_arrive_expect_tx(part_bar, (uint32_t)((KS - 1) * KS_BLK));
#pragma unroll
          for (int o = 1; o < KS; ++o) {
            const uint32_t dst = (crank + (uint32_t)o) % KS;
            ptx::bulk_copy_s2c(ptx::mapa_u32(ptx::smem_u32(part_sm + crank * (KS_BLK / 4)), dst),
                               ptx::smem_u32(send_sm + dst * (KS_BLK / 4)), (uint32_t)KS_BLK, ptx::mapa_u32(part_bar_a, dst));
          }
        }
        ptx::mbar_wait_cluster(part_bar, (n_mma - 1) & 1);
        if (tid == 0) gp_dbg(p.dbg, n_mma, 5);
      } else {
        gp_bar_sync();
      }
      const int64_t base = (t >= 0) ? (int64_t)p.off[t] + m0 : 0;
      float dar[IT][4], daz[IT][4], dan[IT][4];
#pragma unroll
      for (int i = 0; i < IT; ++i) {
        const int e = tid + GP_EPI * i, bl = e / G, g4 = e % G, jl = g4 * 4;
        const int b = m0 + bl;
        if (b >= Bt) continue;
        const bool from_next = b < B_next;
        float dh[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dh[k] = from_next ? carry[i][k] : 0.f;
        if (from_next && mma) {
          const int o4 = ks_chunk_off(bl, g4);
#pragma unroll
          for (int src = 0; src < KS; ++src) {
            const float* blk = (src == (int)crank) ? send_sm + crank * (KS_BLK / 4) : part_sm + src * (KS_BLK / 4);
            const float4 v = *reinterpret_cast<const float4*>(blk + o4);
            dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
          }
        }
        if (t < 0) {
          float* o = p.dh0 + (int64_t)b * d + j0 + jl;
          float4 v = make_float4(dh[0], dh[1], dh[2], dh[3]);
          if (p.dh0_accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(o);
            v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
          }
          *reinterpret_cast<float4*>(o) = v;
          continue;
        }
        const float dyv[4] = {dyp[i].x, dyp[i].y, dyp[i].z, dyp[i].w};
        float r[4], z[4], n[4], g[4], hp[4];
        {
          float2 a, c2;
          a = unpack_bf16x2(sp[i][0].x); c2 = unpack_bf16x2(sp[i][0].y); r[0] = a.x; r[1] = a.y; r[2] = c2.x; r[3] = c2.y;
          a = unpack_bf16x2(sp[i][1].x); c2 = unpack_bf16x2(sp[i][1].y); z[0] = a.x; z[1] = a.y; z[2] = c2.x; z[3] = c2.y;
          a = unpack_bf16x2(sp[i][2].x); c2 = unpack_bf16x2(sp[i][2].y); n[0] = a.x; n[1] = a.y; n[2] = c2.x; n[3] = c2.y;
          a = unpack_bf16x2(sp[i][3].x); c2 = unpack_bf16x2(sp[i][3].y); g[0] = a.x; g[1] = a.y; g[2] = c2.x; g[3] = c2.y;
          a = unpack_bf16x2(sp[i][4].x); c2 = unpack_bf16x2(sp[i][4].y); hp[0] = a.x; hp[1] = a.y; hp[2] = c2.x; hp[3] = c2.y;
        }
        float danr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const GruBwd w = gru_bwd_math(dh[k] + dyv[k], r[k], z[k], n[k], g[k], hp[k]);
          dar[i][k] = w.dar; daz[i][k] = w.daz; dan[i][k] = w.dan; danr[k] = w.dan_r;
          carry[i][k] = w.dh_prev;
        }
        // dgh_t is the next iteration's A operand: on the chain, stored before the release; dgi_t (read by the
        // weight-gradient GEMMs after the kernel) is stored after it
        const int64_t o3 = (base + bl) * d3 + j0 + jl;
        st4_bf16(p.dgh_b + o3, dar[i]);
        st4_bf16(p.dgh_b + o3 + d, daz[i]);
        st4_bf16(p.dgh_b + o3 + 2 * d, danr);
      }
      gp_bar_sync();
      if (tid == 0) gp_dbg(p.dbg, n_mma, 6);
      if (tid == 0) red_release_add(p.sync + bi, 1);
      if (tid == 0) gp_dbg(p.dbg, n_mma, 7);
      if (t >= 0) {
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          const int e = tid + GP_EPI * i, bl = e / G, jl = (e % G) * 4;
          if (m0 + bl < Bt) {
            const int64_t o3 = (base + bl) * d3 + j0 + jl;
            st4_bf16(p.dgi_b + o3, dar[i]);
            st4_bf16(p.dgi_b + o3 + d, daz[i]);
            st4_bf16(p.dgi_b + o3 + 2 * d, dan[i]);
          }
        }
      }
      if (t - 1 >= 0) prefetch(t - 1);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                   // no CTA leaves while a peer may still copy into it
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}


// =====================================================================================================
// HALF-TILE variants (d = 1024-class layers): two independent recurrence chains per CTA
// =====================================================================================================
// Timeline of the kernels above at d = 1024, B = 256 (tools/gru_persist_bench.py): a forward step is 13 300 cycles of
// which only ~4 400 stream the A operand; the rest is a chain of latencies — counter observed -> first TMA byte
// (2 000), MMA tail (1 300), TMEM drain + gate math (2 100), release fence (1 350), counter propagation to the
// slowest CTA (1 600+) — during which the TMA and tensor pipes of the SM idle.  The 128 rows of a batch tile are
// independent graphs, so each CTA now runs TWO chains, the 64-row halves of its tile, with their own counters and
// TMEM accumulators: while half 0 sits in its release/propagation/latency window the CTA loads, multiplies and
// finishes half 1, and vice versa.  Bytes per CTA are unchanged; a step of both halves costs about one chain latency
// (~9 000 cycles) instead of latency + stream.
// The MMAs are M = 64 instructions (half the shared-memory operand traffic of an M = 128 one: measured, the MMA phase
// of a step is bound by the ~64 B/clk at which the tensor core reads its smem operands, 92 cycles per 128 x 48 x 16
// MMA).  Their accumulator layout (tools/probe_m64.cu): row i of the 64 lives in TMEM lane (i % 16) + 32 * (i / 16),
// i.e. each of the four lane quadrants holds 16 rows in its lanes 0..15.
constexpr int H2_ROWS = 64;
constexpr int H2_A_BYTES = H2_ROWS * GP_BK * 2;     // 8 KB per ring stage

template <int STAGES>
__global__ void __launch_bounds__(GP_THREADS, 1) gru_persist_fwd_h2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                           const __grid_constant__ CUtensorMap tmW,
                                                                           const GruPersistFwdParams p) {
  constexpr int DJ = 16, NROWS = 3 * DJ, ACC_LD = NROWS + 1;
  constexpr uint32_t TMEM_COLS = 128;          // two [128 x 48] accumulators at columns 0 and 64
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L;
  const int nkc = d / GP_BK;
  uint8_t* w_sm = smem;
  uint8_t* a_sm = smem + 3 * DJ * d * 2;
  float* acc_sm = reinterpret_cast<float*>(a_sm + (STAGES + 1) * H2_A_BYTES);   // [64][ACC_LD]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(acc_sm + H2_ROWS * ACC_LD + 1);
  full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(full_bar) + 7) & ~(uintptr_t)7);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* w_bar = empty_bar + STAGES;
  uint64_t* tmem_full_bar = w_bar + 1;         // [2]: one per half
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ji = blockIdx.x, bi = blockIdx.y, ns = gridDim.x;
  const int j0 = ji * DJ, m0 = bi * GP_BM;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(&tmem_full_bar[0], 1);
    ptx::mbar_init(&tmem_full_bar[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(3 * DJ * d * 2));
      for (int kc = 0; kc < nkc; ++kc)
        for (int g = 0; g < 3; ++g)
          ptx::tma_load_2d(w_sm + kc * (NROWS * 128) + g * (DJ * 128), &tmW, w_bar, kc * GP_BK, g * d + j0);
      int it = 0;
      for (int t = 0; t < L; ++t) {
        if (m0 >= p.bt[t]) break;
        for (int h = 0; h < 2; ++h) {
          const int mh = m0 + h * H2_ROWS;
          if (mh >= p.bt[t]) continue;
          if (t > 0) wait_counter(p.sync + 2 * bi + h, t * ns);   // every slice of this half's h_{t-1} is in global memory
          if (h == 0) gp_dbg(p.dbg, t, 0);
          asm volatile("fence.proxy.async;" ::: "memory");
          const int row0 = p.off[t] + mh;
          for (int kc = 0; kc < nkc; ++kc, ++it) {
            const int s = it % STAGES;
            ptx::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[s], H2_A_BYTES);
            ptx::tma_load_2d(a_sm + s * H2_A_BYTES, &tmA, &full_bar[s], kc * GP_BK, row0);
          }
          if (h == 0) gp_dbg(p.dbg, t, 1);
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(H2_ROWS, NROWS, 0, 0);
      ptx::mbar_wait(w_bar, 0);
      const uint32_t w_addr = ptx::smem_u32(w_sm), a_addr0 = ptx::smem_u32(a_sm);
      int it = 0;
      for (int t = 0; t < L; ++t) {
        if (m0 >= p.bt[t]) break;
        for (int h = 0; h < 2; ++h) {
          if (m0 + h * H2_ROWS >= p.bt[t]) continue;
          const uint32_t d_tmem = tmem_base + (uint32_t)(h * 64);
          for (int kc = 0; kc < nkc; ++kc, ++it) {
            const int s = it % STAGES;
            ptx::mbar_wait(&full_bar[s], (it / STAGES) & 1);
            if (kc == 0 && h == 0) gp_dbg(p.dbg, t, 2);
            ptx::tc_fence_after();
{   // descriptors built once per chunk, a k-step is an ADD on the (address >> 4) field (see gemm_tc.cu)
              const uint64_t ad = ptx::make_smem_desc_sw128(a_addr0 + s * H2_A_BYTES, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc_sw128(w_addr + kc * (NROWS * 128), 16, 1024);
              ptx::umma_f16(d_tmem, ad, bd, idesc, kc != 0 ? 1u : 0u);
#pragma unroll
              for (int kk = 1; kk < GP_BK / 16; ++kk) ptx::umma_f16_acc(d_tmem, ad + kk * 2, bd + kk * 2, idesc);
            }
            ptx::umma_commit(&empty_bar[s]);
          }
          ptx::umma_commit(&tmem_full_bar[h]);
          if (h == 0) gp_dbg(p.dbg, t, 3);
        }
      }
    }
  } else {
    // epilogue: warps 2..5 drain the 16 rows held by lanes 0..15 of their TMEM lane quadrant (M = 64 layout); then the
    // 256 threads take one (row, 4-unit group) item each
    constexpr int G = DJ / 4;
    const int q = warp & 3;
    const int tid = threadIdx.x - 64;            // 0..255
    const bool drainer = warp < 6;
    const int bl = tid / G, jl = (tid % G) * 4;  // this thread's row (of a half) and unit group
    const int64_t d3 = 3 * (int64_t)d;
    const int bt0 = p.bt[0];
    float hreg[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int b = m0 + h * H2_ROWS + bl;
      const float4 hv = (b < bt0) ? *reinterpret_cast<const float4*>(p.h0 + (int64_t)b * d + j0 + jl) : make_float4(0, 0, 0, 0);
      hreg[h][0] = hv.x; hreg[h][1] = hv.y; hreg[h][2] = hv.z; hreg[h][3] = hv.w;
    }
    const bool drop = p.p_drop > 0.f;
    const float drop_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const uint64_t ctr0 = p.offset + (p.offset_dev ? *p.offset_dev : 0ull);
    const float4 b_r = __ldg(reinterpret_cast<const float4*>(p.b_hh + j0 + jl));
    const float4 b_z = __ldg(reinterpret_cast<const float4*>(p.b_hh + d + j0 + jl));
    const float4 b_n = __ldg(reinterpret_cast<const float4*>(p.b_hh + 2 * d + j0 + jl));
    const float br[4] = {b_r.x, b_r.y, b_r.z, b_r.w}, bz[4] = {b_z.x, b_z.y, b_z.z, b_z.w};
    const float bn[4] = {b_n.x, b_n.y, b_n.z, b_n.w};
    float4 gpre[2][3];
    auto prefetch_gi = [&](int t, int h) {
      const int mh = m0 + h * H2_ROWS;
      if (mh + bl < p.bt[t]) {
        const float* gp = p.gi + ((int64_t)p.off[t] + mh + bl) * d3 + j0 + jl;
        gpre[h][0] = *reinterpret_cast<const float4*>(gp);
        gpre[h][1] = *reinterpret_cast<const float4*>(gp + d);
        gpre[h][2] = *reinterpret_cast<const float4*>(gp + 2 * d);
      }
    };
    if (m0 < bt0) prefetch_gi(0, 0);
    if (m0 + H2_ROWS < bt0) prefetch_gi(0, 1);
    for (int t = 0; t < L; ++t) {
      const int Bt = p.bt[t];
      if (m0 >= Bt) break;
      const int Bn = (t + 1 < L) ? p.bt[t + 1] : 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int mh = m0 + h * H2_ROWS;
        if (mh >= Bt) continue;
        const int64_t base = (int64_t)p.off[t] + mh;
        const int64_t base_n = (t + 1 < L) ? (int64_t)p.off[t + 1] + mh : 0;
        ptx::mbar_wait(&tmem_full_bar[h], t & 1);
        if (tid == 0 && h == 0) gp_dbg(p.dbg, t, 4);
        ptx::tc_fence_after();
        if (drainer && mh + q * 16 < Bt) {
          const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64);
          float* dst = acc_sm + (q * 16 + (lane & 15)) * ACC_LD;
#pragma unroll
          for (int c = 0; c < NROWS; c += 16) {
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)c, v);
            ptx::tmem_ld_wait();
            if (lane < 16) {
#pragma unroll
              for (int k = 0; k < 16; ++k) dst[c + k] = __uint_as_float(v[k]);
            }
          }
        }
        ptx::tc_fence_before();
        gp_bar_sync();
        if (tid == 0 && h == 0) gp_dbg(p.dbg, t, 5);
        const int b = mh + bl;
        float o_r[4], o_z[4], o_n[4], o_g[4];
        if (b < Bt) {
          const float* ap = acc_sm + bl * ACC_LD + jl;
          const float gr[4] = {gpre[h][0].x, gpre[h][0].y, gpre[h][0].z, gpre[h][0].w};
          const float gz[4] = {gpre[h][1].x, gpre[h][1].y, gpre[h][1].z, gpre[h][1].w};
          const float gn[4] = {gpre[h][2].x, gpre[h][2].y, gpre[h][2].z, gpre[h][2].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const GruFwd o = gru_fwd_math_fast(gr[k], gz[k], gn[k], ap[k] + br[k], ap[DJ + k] + bz[k],
                                               ap[2 * DJ + k] + bn[k], hreg[h][k]);
            hreg[h][k] = o.h;
            o_r[k] = o.r; o_z[k] = o.z; o_n[k] = o.n; o_g[k] = o.ghn;
          }
          if (b < Bn) st4_bf16(p.hp_b + (base_n + bl) * d + j0 + jl, hreg[h]);   // the only store on the chain
        }
        gp_bar_sync();                             // the chain's stores are issued (and acc_sm is free again)
        if (tid == 0 && h == 0) gp_dbg(p.dbg, t, 6);
        if (tid == 0) red_release_add(p.sync + 2 * bi + h, 1);
        if (tid == 0 && h == 0) gp_dbg(p.dbg, t, 7);
        if (b < Bt) {
          const int64_t o = (base + bl) * d + j0 + jl;
          if (drop) {   // same draw as dropout_bf16_kernel over the [N, d] output of this layer
            const uint64_t c = ctr0 + (uint64_t)(o >> 2);
            const uint4 rn = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
            const uint32_t rr[4] = {rn.x, rn.y, rn.z, rn.w};
            float yo[4];
            uint32_t mk = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const bool keep = (float)(rr[k] >> 8) * (1.f / 16777216.f) >= p.p_drop;
              mk |= (keep ? 1u : 0u) << (8 * k);
              yo[k] = keep ? bf16_bits_to_f32(f32_to_bf16_bits(hreg[h][k])) * drop_scale : 0.f;
            }
            if (p.mask) *reinterpret_cast<uint32_t*>(p.mask + o) = mk;
            st4_bf16(p.y_b + o, yo);
          } else {
            st4_bf16(p.y_b + o, hreg[h]);
          }
          if (p.r) {
            st4_bf16(p.r + o, o_r);
            st4_bf16(p.z + o, o_z);
            st4_bf16(p.n + o, o_n);
            st4_bf16(p.ghn + o, o_g);
          }
        }
        if (t + 1 < L && mh < Bn) prefetch_gi(t + 1, h);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// K-split backward with half tiles: as gru_persist_bwd_ks_kernel, two chains per CTA.  The [128 x 16] partial blocks
// are exchanged per half (rows 64h..64h+63 of a block are contiguous: one 4 KB bulk copy per peer), each half has its
// own part_bar.
template <int STAGES>
__global__ void __launch_bounds__(GP_THREADS, 1) gru_persist_bwd_ks_h2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                              const __grid_constant__ CUtensorMap tmW,
                                                                              const GruPersistBwdParams p) {
  constexpr int DJ = KS_DJ;
  constexpr uint32_t TMEM_COLS = 128;          // two [128 x 64] accumulators
  constexpr int HBLK = KS_BLK / 2;             // bytes of half a partial block (64 rows x 16 fp32)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L;
  const int kq = 3 * d / KS;
  const int nkc = kq / GP_BK;
  uint8_t* w_sm = smem;
  uint8_t* a_sm = w_sm + nkc * (KS_NC * 128);
  float* send_sm = reinterpret_cast<float*>(a_sm + (STAGES + 1) * H2_A_BYTES);
  float* part_sm = send_sm + KS * (KS_BLK / 4);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(part_sm + KS * (KS_BLK / 4));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* w_bar = empty_bar + STAGES;
  uint64_t* tmem_full_bar = w_bar + 1;         // [2]
  uint64_t* part_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(part_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bi = blockIdx.y, ns = gridDim.x;
  const uint32_t crank = ptx::cluster_ctarank();
  const int jc0 = ((int)blockIdx.x / KS) * KS_NC;
  const int j0 = jc0 + (int)crank * DJ;
  const int k0 = (int)crank * kq;
  const int m0 = bi * GP_BM;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(w_bar, 1);
    for (int h = 0; h < 2; ++h) {
      ptx::mbar_init(&tmem_full_bar[h], 1);
      ptx::mbar_init(&part_bar[h], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // iterations t = L-1 .. -1 of half h (rows m0 + 64h ..): active while the half has rows at step max(t, 0)
  auto half_active = [&](int t, int h) { return m0 + h * H2_ROWS < p.bt[t < 0 ? 0 : t]; };
  auto has_mma = [&](int t, int h) { return (t + 1 <= L - 1) && (m0 + h * H2_ROWS < p.bt[t + 1]); };

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(nkc * KS_NC * 128));
      for (int kc = 0; kc < nkc; ++kc)
        ptx::tma_load_2d(w_sm + kc * (KS_NC * 128), &tmW, w_bar, jc0, k0 + kc * GP_BK);
      int it = 0, done[2] = {0, 0};
      for (int t = L - 1; t >= -1; --t) {
        for (int h = 0; h < 2; ++h) {
          if (!half_active(t, h)) continue;
          if (has_mma(t, h)) {
            wait_counter(p.sync + 2 * bi + h, done[h] * ns);
            if (h == 0) gp_dbg(p.dbg, done[0], 0);
            asm volatile("fence.proxy.async;" ::: "memory");
            const int row0 = p.off[t + 1] + m0 + h * H2_ROWS;
            for (int kc = 0; kc < nkc; ++kc, ++it) {
              const int s = it % STAGES;
              ptx::mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
              ptx::mbar_arrive_expect_tx(&full_bar[s], H2_A_BYTES);
              ptx::tma_load_2d(a_sm + s * H2_A_BYTES, &tmA, &full_bar[s], k0 + kc * GP_BK, row0);
            }
            if (h == 0) gp_dbg(p.dbg, done[0], 1);
          }
          ++done[h];
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(H2_ROWS, KS_NC, 0, 1);
      ptx::mbar_wait(w_bar, 0);
      const uint32_t w_addr = ptx::smem_u32(w_sm), a_addr0 = ptx::smem_u32(a_sm);
      int it = 0, nm0 = 0;
      for (int t = L - 1; t >= -1; --t) {
        for (int h = 0; h < 2; ++h) {
          if (!half_active(t, h) || !has_mma(t, h)) continue;
          if (h == 0) ++nm0;
          const uint32_t d_tmem = tmem_base + (uint32_t)(h * 64);
          for (int kc = 0; kc < nkc; ++kc, ++it) {
            const int s = it % STAGES;
            ptx::mbar_wait(&full_bar[s], (it / STAGES) & 1);
            if (kc == 0 && h == 0) gp_dbg(p.dbg, nm0, 2);
            ptx::tc_fence_after();
{   // descriptors built once per chunk, a k-step is an ADD on the (address >> 4) field (see gemm_tc.cu)
              const uint64_t ad = ptx::make_smem_desc_sw128(a_addr0 + s * H2_A_BYTES, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc_sw128(w_addr + kc * (KS_NC * 128), GP_BK * 128, 1024);
              ptx::umma_f16(d_tmem, ad, bd, idesc, kc != 0 ? 1u : 0u);
#pragma unroll
              for (int kk = 1; kk < GP_BK / 16; ++kk) ptx::umma_f16_acc(d_tmem, ad + kk * 2, bd + kk * 128, idesc);
            }
            ptx::umma_commit(&empty_bar[s]);
          }
          ptx::umma_commit(&tmem_full_bar[h]);
          if (h == 0) gp_dbg(p.dbg, nm0, 3);
        }
      }
    }
  } else {
    constexpr int G = DJ / 4;
    const int q = warp & 3;
    const int tid = threadIdx.x - 64;            // 0..255: one (row, 4-unit group) item per half
    const bool drainer = warp < 6;               // M = 64 accumulator: 16 rows in lanes 0..15 of every lane quadrant
    const int bl = tid / G, g4 = tid % G, jl = g4 * 4;
    const int64_t d3 = 3 * (int64_t)d;
    float carry[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int k = 0; k < 4; ++k) carry[h][k] = 0.f;
    float4 dyp[2];
    uint2 sp[2][5];
    auto prefetch = [&](int t, int h) {
      const int mh = m0 + h * H2_ROWS;
      if (mh + bl < p.bt[t]) {
        const int64_t o = ((int64_t)p.off[t] + mh + bl) * d + j0 + jl;
        dyp[h] = *reinterpret_cast<const float4*>(p.dy + o);
        if (p.dy_mask) {
          const uint32_t mk = *reinterpret_cast<const uint32_t*>(p.dy_mask + o);
          dyp[h].x = (mk & 0xFFu) ? dyp[h].x * p.dy_scale : 0.f;
          dyp[h].y = (mk & 0xFF00u) ? dyp[h].y * p.dy_scale : 0.f;
          dyp[h].z = (mk & 0xFF0000u) ? dyp[h].z * p.dy_scale : 0.f;
          dyp[h].w = (mk & 0xFF000000u) ? dyp[h].w * p.dy_scale : 0.f;
        }
        sp[h][0] = *reinterpret_cast<const uint2*>(p.r + o);
        sp[h][1] = *reinterpret_cast<const uint2*>(p.z + o);
        sp[h][2] = *reinterpret_cast<const uint2*>(p.n + o);
        sp[h][3] = *reinterpret_cast<const uint2*>(p.ghn + o);
        sp[h][4] = *reinterpret_cast<const uint2*>(p.hp_b + o);
      }
    };
    int t_first[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      t_first[h] = L - 1;
      while (t_first[h] >= 0 && !half_active(t_first[h], h)) --t_first[h];
      if (t_first[h] >= 0) prefetch(t_first[h], h);
    }
    int n_mma[2] = {0, 0};
    const uint32_t part_bar_a[2] = {ptx::smem_u32(&part_bar[0]), ptx::smem_u32(&part_bar[1])};
    for (int t = L - 1; t >= -1; --t) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (t > t_first[h] || !half_active(t, h)) continue;
        const int mh = m0 + h * H2_ROWS;
        const bool mma = has_mma(t, h);
        const int B_next = (t + 1 <= L - 1) ? p.bt[t + 1] : 0;
        const int Bt = p.bt[t < 0 ? 0 : t];
        if (mma) {
          ptx::mbar_wait(&tmem_full_bar[h], n_mma[h] & 1);
          ptx::tc_fence_after();
          ++n_mma[h];
          if (tid == 0 && h == 0) gp_dbg(p.dbg, n_mma[0], 4);
          if (drainer) {
            const int row = h * H2_ROWS + q * 16 + (lane & 15);    // row of the full [128 x 16] block
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64);
#pragma unroll
            for (int dst = 0; dst < KS; ++dst) {
              uint32_t v[16];
              ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(dst * DJ), v);
              ptx::tmem_ld_wait();
              float* blk = send_sm + dst * (KS_BLK / 4);
              if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  *reinterpret_cast<uint4*>(blk + ks_chunk_off(row, c)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
              }
            }
          }
          ptx::tc_fence_before();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          gp_bar_sync();
          if (tid == 0) {
            ptx::mbar_arrive_expect_tx(&part_bar[h], (uint32_t)((KS - 1) * HBLK));
#pragma unroll
            for (int o = 1; o < KS; ++o) {
              const uint32_t dst = (crank + (uint32_t)o) % KS;
              ptx::bulk_copy_s2c(ptx::mapa_u32(ptx::smem_u32(part_sm + crank * (KS_BLK / 4) + h * (HBLK / 4)), dst),
                                 ptx::smem_u32(send_sm + dst * (KS_BLK / 4) + h * (HBLK / 4)), (uint32_t)HBLK,
                                 ptx::mapa_u32(part_bar_a[h], dst));
            }
          }
          ptx::mbar_wait_cluster(&part_bar[h], (n_mma[h] - 1) & 1);
          if (tid == 0 && h == 0) gp_dbg(p.dbg, n_mma[0], 5);
        }
        const int64_t base = (t >= 0) ? (int64_t)p.off[t] + mh : 0;
        const int b = mh + bl;
        const bool live = b < Bt;
        float dar[4], daz[4], dan[4];
        if (live) {
          const bool from_next = b < B_next;
          float dh[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) dh[k] = from_next ? carry[h][k] : 0.f;
          if (from_next && mma) {
            const int o4 = ks_chunk_off(h * H2_ROWS + bl, g4);
#pragma unroll
            for (int src = 0; src < KS; ++src) {
              const float* blk = (src == (int)crank) ? send_sm + crank * (KS_BLK / 4) : part_sm + src * (KS_BLK / 4);
              const float4 v = *reinterpret_cast<const float4*>(blk + o4);
              dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
            }
          }
          if (t < 0) {
            float* o = p.dh0 + (int64_t)b * d + j0 + jl;
            float4 v = make_float4(dh[0], dh[1], dh[2], dh[3]);
            if (p.dh0_accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(o);
              v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
            }
            *reinterpret_cast<float4*>(o) = v;
          } else {
            const float dyv[4] = {dyp[h].x, dyp[h].y, dyp[h].z, dyp[h].w};
            float r[4], z[4], n[4], g[4], hp[4];
            {
              float2 a, c2;
              a = unpack_bf16x2(sp[h][0].x); c2 = unpack_bf16x2(sp[h][0].y); r[0] = a.x; r[1] = a.y; r[2] = c2.x; r[3] = c2.y;
              a = unpack_bf16x2(sp[h][1].x); c2 = unpack_bf16x2(sp[h][1].y); z[0] = a.x; z[1] = a.y; z[2] = c2.x; z[3] = c2.y;
              a = unpack_bf16x2(sp[h][2].x); c2 = unpack_bf16x2(sp[h][2].y); n[0] = a.x; n[1] = a.y; n[2] = c2.x; n[3] = c2.y;
              a = unpack_bf16x2(sp[h][3].x); c2 = unpack_bf16x2(sp[h][3].y); g[0] = a.x; g[1] = a.y; g[2] = c2.x; g[3] = c2.y;
              a = unpack_bf16x2(sp[h][4].x); c2 = unpack_bf16x2(sp[h][4].y); hp[0] = a.x; hp[1] = a.y; hp[2] = c2.x; hp[3] = c2.y;
            }
            float danr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const GruBwd w = gru_bwd_math(dh[k] + dyv[k], r[k], z[k], n[k], g[k], hp[k]);
              dar[k] = w.dar; daz[k] = w.daz; dan[k] = w.dan; danr[k] = w.dan_r;
              carry[h][k] = w.dh_prev;
            }
            const int64_t o3 = (base + bl) * d3 + j0 + jl;
            st4_bf16(p.dgh_b + o3, dar);                    // dgh_t: the next iteration's A operand, on the chain
            st4_bf16(p.dgh_b + o3 + d, daz);
            st4_bf16(p.dgh_b + o3 + 2 * d, danr);
          }
        }
        gp_bar_sync();
        if (tid == 0 && h == 0) gp_dbg(p.dbg, n_mma[0], 6);
        if (tid == 0) red_release_add(p.sync + 2 * bi + h, 1);
        if (tid == 0 && h == 0) gp_dbg(p.dbg, n_mma[0], 7);
        if (live && t >= 0) {
          const int64_t o3 = (base + bl) * d3 + j0 + jl;
          st4_bf16(p.dgi_b + o3, dar);
          st4_bf16(p.dgi_b + o3 + d, daz);
          st4_bf16(p.dgi_b + o3 + 2 * d, dan);
        }
        if (t - 1 >= 0) prefetch(t - 1, h);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// [R, C] bf16 -> [C, R] bf16 (W_hh^T for the backward kernel; 32x32 tiles through padded smem)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const uint16_t* __restrict__ in, int R, int C,
                                                             uint16_t* __restrict__ out) {
  __shared__ uint16_t tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < R && c0 + tx < C) tile[i][tx] = in[(int64_t)(r0 + i) * C + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && r0 + tx < R) out[(int64_t)(c0 + i) * R + r0 + tx] = tile[tx][i];
}

// 64x64 tiles, 4-byte accesses on both sides (R, C even; 4-byte aligned bases): a warp touches 128 contiguous bytes
// per row instead of 64, a quarter of the CTAs
__global__ void __launch_bounds__(256) transpose_bf16_x2_kernel(const uint16_t* __restrict__ in, int R, int C,
                                                                uint16_t* __restrict__ out) {
  __shared__ uint16_t tile[64][66];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 64; i += 8) {
    uint32_t v = 0;
    if (r0 + i < R && c0 + 2 * tx < C) v = *reinterpret_cast<const uint32_t*>(in + (int64_t)(r0 + i) * C + c0 + 2 * tx);
    *reinterpret_cast<uint32_t*>(&tile[i][2 * tx]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 64; i += 8) {
    if (c0 + i < C && r0 + 2 * tx < R) {
      const uint32_t v = (uint32_t)tile[2 * tx][i] | ((uint32_t)tile[2 * tx + 1][i] << 16);
      *reinterpret_cast<uint32_t*>(out + (int64_t)(c0 + i) * R + r0 + 2 * tx) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------- host
static int max_stages() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ARK_GRU_STAGES"); v = e ? atoi(e) : 4; if (v < 2) v = 2; if (v > 6) v = 6; }
  return v;
}
static long long* g_pdbg = nullptr;      // ARK_GRU_PERSIST_DBG=1: [fwd 4x8 | bwd 4x8] clock64 words
static long long* pdbg_buffer() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ARK_GRU_PERSIST_DBG"); on = e ? atoi(e) : 0; }
  if (!on) return nullptr;
  if (!g_pdbg) {
    if (cudaMalloc(&g_pdbg, sizeof(long long) * 64) != cudaSuccess) { (void)cudaGetLastError(); g_pdbg = nullptr; return nullptr; }
    cudaMemset(g_pdbg, 0, sizeof(long long) * 64);
  }
  return g_pdbg;
}
static int pick_dj(int64_t d, int64_t bt0, int* stages_out) {
  if (d % 64 != 0 || d < 64 || bt0 <= 0) return 0;
  const int64_t nbt = (bt0 + GP_BM - 1) / GP_BM;
  const int cand[3] = {16, 32, 64};
  for (int i = 0; i < 3; ++i) {
    const int dj = cand[i];
    if (d % dj) continue;
    if (nbt * (d / dj) > kNumSMs) continue;
    for (int st = (dj == 16 ? max_stages() : (max_stages() < 4 ? max_stages() : 4)); st >= 2; --st) {
      const int64_t smem = 3LL * dj * d * 2 + (int64_t)st * GP_A_BYTES + 128LL * (3 * dj + 1) * 4 + 2048;
      if (smem <= 227 * 1024) {
        *stages_out = st;
        return dj;
      }
    }
  }
  return 0;
}

template <typename Params, typename Kern>
static int launch_coop(Kern kern, const CUtensorMap& tmA, const CUtensorMap& tmW, const Params& prm, dim3 grid,
                       int smem, cudaStream_t s, const char* who, int cluster) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail((int)e, "%s: smem attribute (%d B): %s", who, smem, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GP_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeCooperative;
  attrs[0].val.cooperative = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = (unsigned)cluster;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = cluster > 1 ? 2 : 1;
  if (cluster > 1 && under_profiler()) {      // cluster launch without the cooperative attribute (see common.cuh)
    attrs[0] = attrs[1];
    cfg.numAttrs = 1;
  }
  e = cudaLaunchKernelEx(&cfg, kern, tmA, tmW, prm);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail((int)e, "%s: cooperative launch grid=(%u,%u) cluster=%d: %s", who, grid.x, grid.y, cluster, cudaGetErrorString(e));
  }
  count_launch();
  return 0;
}

// Largest cluster width (4, 2, 1) along the hidden-slice dimension for which ALL clusters of the grid are co-resident
// (the kernel spins on flags written by other CTAs, so co-residency is a correctness requirement).
template <typename Kern>
static int pick_cluster(Kern k8, Kern k4, Kern k2, dim3 grid, int smem) {
  static int forced = -1;
  if (forced < 0) {
    // Default 1 (no clusters): measured on B200 (syn-types, d=1024) the 4-CTA multicast cuts the L2 reads of the A
    // tile 4x but not the step time — the chain is bound by the latency of the 4-stage ring, not by L2 bandwidth.
    const char* env = getenv("ARK_GRU_CLUSTER");
    forced = env ? atoi(env) : 1;
  }
  // the answer depends only on (kernel, grid, smem): remember it (the occupancy query costs host time every step)
  static thread_local struct { const void* k; unsigned gx, gy; int smem, cs; } memo[16];
  static thread_local int n_memo = 0;
  for (int i = 0; i < n_memo; ++i)
    if (memo[i].k == (const void*)k8 && memo[i].gx == grid.x && memo[i].gy == grid.y && memo[i].smem == smem) return memo[i].cs;
  auto remember = [&](int cs) {
    if (n_memo < 16) { memo[n_memo].k = (const void*)k8; memo[n_memo].gx = grid.x; memo[n_memo].gy = grid.y; memo[n_memo].smem = smem; memo[n_memo].cs = cs; ++n_memo; }
    return cs;
  };
  const int cand[3] = {8, 4, 2};
  for (int i = 0; i < 3; ++i) {
    const int cs = cand[i];
    if (cs > forced) continue;
    if (grid.x % cs) continue;
    Kern k = cs == 8 ? k8 : (cs == 4 ? k4 : k2);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) continue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(GP_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cs > 8) continue;
    const cudaError_t qe = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    if (qe != cudaSuccess) {
      if (getenv("ARK_GRU_DEBUG")) fprintf(stderr, "[arkb200] cudaOccupancyMaxActiveClusters(cluster=%d): %s\n", cs, cudaGetErrorString(qe));
      (void)cudaGetLastError();
      continue;
    }
    if (getenv("ARK_GRU_DEBUG"))
      fprintf(stderr, "[arkb200] gru_persist grid=(%u,%u) smem=%d cluster=%d: max active clusters %d (need %u)\n", grid.x,
              grid.y, smem, cs, n, grid.x * grid.y / cs);
    if ((long long)n * cs >= (long long)grid.x * grid.y) return remember(cs);
  }
  return remember(1);
}


// all CTAs (clusters) of a cooperative launch must be co-resident: occupancy query, memoised per (kernel, grid, smem)
static bool gp_coresident(const void* kern, dim3 grid, int smem, int cluster) {
  struct Memo { const void* k; unsigned gx, gy; int smem, cluster, ok; };
  static thread_local std::vector<Memo> memo;
  for (const Memo& m : memo)
    if (m.k == kern && m.gx == grid.x && m.gy == grid.y && m.smem == smem && m.cluster == cluster) return m.ok != 0;
  int ok = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(GP_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess) ok = ((long long)n * cluster >= (long long)grid.x * grid.y) ? 1 : 0;
    else (void)cudaGetLastError();
    if (getenv("ARK_GRU_DEBUG"))
      fprintf(stderr, "[arkb200] gru_persist variant grid=(%u,%u) cluster=%d smem=%d: max active clusters %d -> %s\n", grid.x,
              grid.y, cluster, smem, n, ok ? "ok" : "does not fit");
  } else {
    (void)cudaGetLastError();
  }
  memo.push_back(Memo{kern, grid.x, grid.y, smem, cluster, ok});
  return ok != 0;
}
static bool h2_wanted() {
  static int want = -1;
  if (want < 0) { const char* e = getenv("ARK_GRU_H2"); want = e ? atoi(e) : 1; }
  return want != 0;
}
// half-tile forward kernel: ring stages (12 / 8 / 4) and shared-memory bytes, or 0 when it does not apply (it pays for
// d >= 512 with more than one half tile of rows; 16-unit slices must fill <= 148 SMs)
static int h2_fwd_plan(int64_t d, int64_t bt0, int* stages_out) {
  if (!h2_wanted() || d < 512 || d % 64 != 0 || bt0 <= H2_ROWS) return 0;
  const int64_t nbt = (bt0 + GP_BM - 1) / GP_BM;
  if (nbt * (d / 16) > kNumSMs) return 0;
  const int cand[3] = {12, 8, 4};
  for (int i = 0; i < 3; ++i) {
    const int st = cand[i];
    const int64_t smem = 3LL * 16 * d * 2 + (int64_t)(st + 1) * H2_A_BYTES + (int64_t)H2_ROWS * 49 * 4 + (2 * st + 4) * 8 + 64 + 1024;
    if (smem > 227 * 1024) continue;
    const void* k = st == 12 ? (const void*)gru_persist_fwd_h2_kernel<12> : (st == 8 ? (const void*)gru_persist_fwd_h2_kernel<8> : (const void*)gru_persist_fwd_h2_kernel<4>);
    if (!gp_coresident(k, dim3((unsigned)(d / 16), (unsigned)nbt), (int)smem, 1)) continue;
    *stages_out = st;
    return (int)smem;
  }
  return 0;
}
static int h2_bwd_plan(int64_t d, int64_t bt0, int* stages_out) {
  if (!h2_wanted() || d < 512 || d % 256 != 0 || bt0 <= H2_ROWS) return 0;
  const int64_t nbt = (bt0 + GP_BM - 1) / GP_BM;
  if (nbt * (d / KS_DJ) > kNumSMs) return 0;
  const int cand[2] = {7, 4};
  for (int i = 0; i < 2; ++i) {
    const int st = cand[i];
    const int64_t smem = (3 * d / KS) * KS_NC * 2 + (int64_t)(st + 1) * H2_A_BYTES + 2LL * KS * KS_BLK + (2 * st + 6) * 8 + 64 + 1024;
    if (smem > 227 * 1024) continue;
    const void* k = st == 7 ? (const void*)gru_persist_bwd_ks_h2_kernel<7> : (const void*)gru_persist_bwd_ks_h2_kernel<4>;
    if (!gp_coresident(k, dim3((unsigned)(d / KS_DJ), (unsigned)nbt), (int)smem, KS)) continue;
    *stages_out = st;
    return (int)smem;
  }
  return 0;
}

// K-split backward (gru_persist_bwd_ks_kernel): shared-memory bytes and ring depth, or 0 when it does not apply.
// It pays where the dgh tile is large (d >= 512) and needs 3d/4 % 64 == 0, 16-unit slices filling <= 148 SMs and all
// d/64 * nbt clusters of 4 co-resident (ARK_GRU_KSPLIT=0 switches it off).
static int ks_plan(int64_t d, int64_t bt0, int* stages_out) {
  static int want = -1;
  if (want < 0) { const char* e = getenv("ARK_GRU_KSPLIT"); want = e ? atoi(e) : 1; }
  if (!want || d < 512 || d % 256 != 0 || bt0 <= 0) return 0;
  const int64_t nbt = (bt0 + GP_BM - 1) / GP_BM;
  if (nbt * (d / KS_DJ) > kNumSMs) return 0;
  for (int st = 4; st >= 2; --st) {
    const int64_t smem = (3 * d / KS) * KS_NC * 2 + (int64_t)st * GP_A_BYTES + 2LL * KS * KS_BLK + (2 * st + 4) * 8 + 16 + 1024;
    if (smem > 227 * 1024) continue;
    // co-residency of every cluster (the kernel spins on counters written by the other clusters)
    static thread_local struct { int64_t d, nbt; int st, ok; } memo[8];
    static thread_local int n_memo = 0;
    int ok = -1;
    for (int i = 0; i < n_memo; ++i)
      if (memo[i].d == d && memo[i].nbt == nbt && memo[i].st == st) ok = memo[i].ok;
    if (ok < 0) {
      auto kern = st == 4 ? gru_persist_bwd_ks_kernel<4> : (st == 3 ? gru_persist_bwd_ks_kernel<3> : gru_persist_bwd_ks_kernel<2>);
      ok = 0;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(d / KS_DJ), (unsigned)nbt);
        cfg.blockDim = dim3(GP_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = KS;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess) ok = ((int64_t)n * KS >= nbt * (d / KS_DJ)) ? 1 : 0;
        else (void)cudaGetLastError();
        if (getenv("ARK_GRU_DEBUG")) fprintf(stderr, "[arkb200] gru_persist_bwd_ks d=%lld nbt=%lld stages=%d smem=%lld: max active clusters %d\n", (long long)d, (long long)nbt, st, (long long)smem, n);
      } else {
        (void)cudaGetLastError();
      }
      if (n_memo < 8) { memo[n_memo].d = d; memo[n_memo].nbt = nbt; memo[n_memo].st = st; memo[n_memo].ok = ok; ++n_memo; }
    }
    if (ok) { *stages_out = st; return (int)smem; }
  }
  return 0;
}

}  // namespace ark

using namespace ark;

extern "C" int ark_gru_persist_debug_dump(int64_t* out_host, int64_t n_words) {
  if (!g_pdbg) return fail(ARK_E_BADARG, "gru_persist_debug_dump: ARK_GRU_PERSIST_DBG was not set");
  if (n_words > 64) n_words = 64;
  cudaError_t e = cudaMemcpy(out_host, g_pdbg, sizeof(long long) * n_words, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? 0 : fail((int)e, "gru_persist_debug_dump: %s", cudaGetErrorString(e));
}

extern "C" int ark_gru_persist_bwd_ksplit(int64_t d, int64_t bt0) {
  int st;
  return ks_plan(d, bt0, &st) > 0 ? 1 : 0;
}

extern "C" int ark_gru_persist_supported(int64_t d, int64_t bt0) {
  int st;
  return pick_dj(d, bt0, &st);
}

extern "C" int ark_transpose_bf16(const uint16_t* in, int64_t R, int64_t C, uint16_t* out, void* stream) {
  ARK_REQUIRE(in && out && R > 0 && C > 0, ARK_E_BADARG, "transpose_bf16: bad arguments");
  if (R % 2 == 0 && C % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    dim3 grid((unsigned)((C + 63) / 64), (unsigned)((R + 63) / 64));
    transpose_bf16_x2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (int)R, (int)C, out);
  } else {
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((R + 31) / 32));
    transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (int)R, (int)C, out);
  }
  return launched("transpose_bf16");
}

extern "C" int ark_gru_persist_fwd(uint16_t* hp_b, const float* h0, const uint16_t* Whh_b, const float* gi,
                                   const float* b_hh, const int32_t* bt_dev, const int32_t* off_dev, int64_t L,
                                   int64_t bt0, int64_t N, int64_t d, uint16_t* y_b, uint16_t* r, uint16_t* z,
                                   uint16_t* n, uint16_t* ghn, uint8_t* mask, float p_drop, uint64_t seed,
                                   uint64_t offset, const uint64_t* offset_dev, int32_t* sync_ws, void* stream) {
  ARK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, ARK_E_BADARG, "gru_persist_fwd: dropout probability must be in [0,1)");
  ARK_REQUIRE(hp_b && h0 && Whh_b && gi && b_hh && bt_dev && off_dev && y_b && sync_ws, ARK_E_BADARG,
              "gru_persist_fwd: null pointer");
  ARK_REQUIRE((r && z && n && ghn) || (!r && !z && !n && !ghn), ARK_E_BADARG,
              "gru_persist_fwd: gate outputs must be all set or all NULL");
  ARK_REQUIRE(L > 0 && N > 0 && bt0 > 0, ARK_E_BADARG, "gru_persist_fwd: bad sizes");
  int stages = 0;
  const int dj = pick_dj(d, bt0, &stages);
  ARK_REQUIRE(dj > 0, ARK_E_SHAPE, "gru_persist_fwd: unsupported shape d=%lld bt0=%lld (need d %% 64 == 0 and the "
              "grid to fit 148 SMs)", (long long)d, (long long)bt0);
  ARK_REQUIRE(aligned16(hp_b) && aligned16(Whh_b) && aligned16(gi) && aligned16(h0) && aligned16(y_b), ARK_E_ALIGN,
              "gru_persist_fwd: 16-byte alignment");
  cudaStream_t s = (cudaStream_t)stream;
  const int nbt = (int)((bt0 + GP_BM - 1) / GP_BM);
  cudaError_t e = cudaMemsetAsync(sync_ws, 0, sizeof(int32_t) * 2 * nbt, s);     // (two counters per tile: half-tile kernels)
  if (e != cudaSuccess) return fail((int)e, "gru_persist_fwd: memset: %s", cudaGetErrorString(e));
  CUtensorMap tmA, tmW;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmW, Whh_b, (uint64_t)d, (uint64_t)(3 * d), (uint64_t)d, GP_BK, dj))) return rc;
  GruPersistFwdParams prm;
  prm.bt = bt_dev; prm.off = off_dev; prm.L = (int)L; prm.d = (int)d; prm.sync = sync_ws; prm.gi = gi; prm.b_hh = b_hh;
  prm.h0 = h0; prm.hp_b = hp_b; prm.y_b = y_b; prm.r = r; prm.z = z; prm.n = n; prm.ghn = ghn;
  prm.mask = p_drop > 0.f ? mask : nullptr; prm.p_drop = p_drop; prm.seed = seed; prm.offset = offset;
  prm.offset_dev = p_drop > 0.f ? offset_dev : nullptr;
  prm.dbg = pdbg_buffer();
  dim3 grid((unsigned)(d / dj), (unsigned)nbt);
  {
    int h2_st = 0;
    const int h2_smem = h2_fwd_plan(d, bt0, &h2_st);
    if (h2_smem > 0) {       // two 64-row chains per CTA (16-unit slices)
      CUtensorMap tmAh, tmWh;
      if ((rc = make_tmap_2d_bf16(&tmWh, Whh_b, (uint64_t)d, (uint64_t)(3 * d), (uint64_t)d, GP_BK, 16))) return rc;
      if ((rc = make_tmap_2d_bf16(&tmAh, hp_b, (uint64_t)d, (uint64_t)N, (uint64_t)d, GP_BK, H2_ROWS))) return rc;
      dim3 gh((unsigned)(d / 16), (unsigned)nbt);
      if (h2_st == 12) return launch_coop(gru_persist_fwd_h2_kernel<12>, tmAh, tmWh, prm, gh, h2_smem, s, "gru_persist_fwd_h2", 1);
      if (h2_st == 8) return launch_coop(gru_persist_fwd_h2_kernel<8>, tmAh, tmWh, prm, gh, h2_smem, s, "gru_persist_fwd_h2", 1);
      return launch_coop(gru_persist_fwd_h2_kernel<4>, tmAh, tmWh, prm, gh, h2_smem, s, "gru_persist_fwd_h2", 1);
    }
  }
  const int smem = (int)(3LL * dj * d * 2 + (int64_t)stages * GP_A_BYTES + 128LL * (3 * dj + 1) * 4 + 2048);
#define ARK_GP_FWD(DJ, ST)                                                                                              \
  if (dj == DJ && stages == ST) {                                                                                       \
    const int cs = (bt0 >= GP_BM) ? pick_cluster(gru_persist_fwd_kernel<DJ, ST, 8>, gru_persist_fwd_kernel<DJ, ST, 4>, gru_persist_fwd_kernel<DJ, ST, 2>, grid, smem) : 1; \
    if ((rc = make_tmap_2d_bf16(&tmA, hp_b, (uint64_t)d, (uint64_t)N, (uint64_t)d, GP_BK, GP_BM / cs))) return rc;       \
    if (cs == 8) return launch_coop(gru_persist_fwd_kernel<DJ, ST, 8>, tmA, tmW, prm, grid, smem, s, "gru_persist_fwd", 8); \
    if (cs == 4) return launch_coop(gru_persist_fwd_kernel<DJ, ST, 4>, tmA, tmW, prm, grid, smem, s, "gru_persist_fwd", 4); \
    if (cs == 2) return launch_coop(gru_persist_fwd_kernel<DJ, ST, 2>, tmA, tmW, prm, grid, smem, s, "gru_persist_fwd", 2); \
    return launch_coop(gru_persist_fwd_kernel<DJ, ST, 1>, tmA, tmW, prm, grid, smem, s, "gru_persist_fwd", 1);          \
  }
  ARK_GP_FWD(16, 6); ARK_GP_FWD(16, 5); ARK_GP_FWD(16, 4); ARK_GP_FWD(16, 3); ARK_GP_FWD(16, 2);
  ARK_GP_FWD(32, 4); ARK_GP_FWD(32, 3); ARK_GP_FWD(32, 2);
  ARK_GP_FWD(64, 4); ARK_GP_FWD(64, 3); ARK_GP_FWD(64, 2);
#undef ARK_GP_FWD
  return fail(ARK_E_SHAPE, "gru_persist_fwd: no kernel instance for dj=%d stages=%d", dj, stages);
}

extern "C" int ark_gru_persist_bwd(const float* dy, const uint16_t* r, const uint16_t* z, const uint16_t* n,
                                   const uint16_t* ghn, const uint16_t* hp_b, const uint16_t* WhhT_b,
                                   const int32_t* bt_dev, const int32_t* off_dev, int64_t L, int64_t bt0, int64_t N,
                                   int64_t d, uint16_t* dgi_b, uint16_t* dgh_b, float* dh0, int dh0_accumulate,
                                   const uint16_t* Whh_b, const uint8_t* dy_mask, float p_drop, int32_t* sync_ws,
                                   void* stream) {
  ARK_REQUIRE(dy && r && z && n && ghn && hp_b && (WhhT_b || Whh_b) && bt_dev && off_dev && dgi_b && dgh_b && dh0 && sync_ws,
              ARK_E_BADARG, "gru_persist_bwd: null pointer");
  ARK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, ARK_E_BADARG, "gru_persist_bwd: dropout probability must be in [0,1)");
  ARK_REQUIRE(L > 0 && N > 0 && bt0 > 0, ARK_E_BADARG, "gru_persist_bwd: bad sizes");
  int stages = 0;
  const int dj = pick_dj(d, bt0, &stages);
  ARK_REQUIRE(dj > 0, ARK_E_SHAPE, "gru_persist_bwd: unsupported shape d=%lld bt0=%lld", (long long)d, (long long)bt0);
  ARK_REQUIRE(aligned16(dy) && aligned16(WhhT_b) && aligned16(Whh_b) && aligned16(dgi_b) && aligned16(dgh_b) && aligned16(dh0),
              ARK_E_ALIGN, "gru_persist_bwd: 16-byte alignment");
  cudaStream_t s = (cudaStream_t)stream;
  const int nbt = (int)((bt0 + GP_BM - 1) / GP_BM);
  cudaError_t e = cudaMemsetAsync(sync_ws, 0, sizeof(int32_t) * 2 * nbt, s);
  if (e != cudaSuccess) return fail((int)e, "gru_persist_bwd: memset: %s", cudaGetErrorString(e));
  CUtensorMap tmA, tmW;
  int rc;
  GruPersistBwdParams prm;
  prm.bt = bt_dev; prm.off = off_dev; prm.L = (int)L; prm.d = (int)d; prm.sync = sync_ws; prm.dy = dy; prm.r = r;
  prm.z = z; prm.n = n; prm.ghn = ghn; prm.hp_b = hp_b; prm.dgi_b = dgi_b; prm.dgh_b = dgh_b; prm.dh0 = dh0;
  prm.dh0_accumulate = dh0_accumulate;
  prm.dy_mask = (p_drop > 0.f) ? dy_mask : nullptr;
  prm.dy_scale = 1.f / (1.f - p_drop);
  prm.dbg = pdbg_buffer() ? pdbg_buffer() + 32 : nullptr;
  dim3 grid((unsigned)(d / dj), (unsigned)nbt);
  {
    int ks_st = 0;
    const int ks_smem = Whh_b ? ks_plan(d, bt0, &ks_st) : 0;
    int h2_st = 0;
    const int h2_smem = ks_smem > 0 ? h2_bwd_plan(d, bt0, &h2_st) : 0;
    if (h2_smem > 0) {       // K-split clusters AND two 64-row chains per CTA
      CUtensorMap tmAk, tmWk;
      if ((rc = make_tmap_2d_bf16(&tmWk, Whh_b, (uint64_t)d, (uint64_t)(3 * d), (uint64_t)d, 64, GP_BK))) return rc;
      if ((rc = make_tmap_2d_bf16(&tmAk, dgh_b, (uint64_t)(3 * d), (uint64_t)N, (uint64_t)(3 * d), GP_BK, H2_ROWS))) return rc;
      dim3 gk((unsigned)(d / KS_DJ), (unsigned)nbt);
      if (h2_st == 7) return launch_coop(gru_persist_bwd_ks_h2_kernel<7>, tmAk, tmWk, prm, gk, h2_smem, s, "gru_persist_bwd_ks_h2", KS);
      return launch_coop(gru_persist_bwd_ks_h2_kernel<4>, tmAk, tmWk, prm, gk, h2_smem, s, "gru_persist_bwd_ks_h2", KS);
    }
    if (ks_smem > 0) {
      CUtensorMap tmAk, tmWk;
      // W_hh [3d, d] untransposed: inner = units, box = 64 units x 64 gate rows
      if ((rc = make_tmap_2d_bf16(&tmWk, Whh_b, (uint64_t)d, (uint64_t)(3 * d), (uint64_t)d, 64, GP_BK))) return rc;
      if ((rc = make_tmap_2d_bf16(&tmAk, dgh_b, (uint64_t)(3 * d), (uint64_t)N, (uint64_t)(3 * d), GP_BK, GP_BM))) return rc;
      dim3 gk((unsigned)(d / KS_DJ), (unsigned)nbt);
      if (ks_st == 4) return launch_coop(gru_persist_bwd_ks_kernel<4>, tmAk, tmWk, prm, gk, ks_smem, s, "gru_persist_bwd_ks", KS);
      if (ks_st == 3) return launch_coop(gru_persist_bwd_ks_kernel<3>, tmAk, tmWk, prm, gk, ks_smem, s, "gru_persist_bwd_ks", KS);
      return launch_coop(gru_persist_bwd_ks_kernel<2>, tmAk, tmWk, prm, gk, ks_smem, s, "gru_persist_bwd_ks", KS);
    }
  }
  ARK_REQUIRE(WhhT_b, ARK_E_BADARG, "gru_persist_bwd: this shape runs the N-sliced kernel, which needs W_hh^T (WhhT_b)");
  if ((rc = make_tmap_2d_bf16(&tmW, WhhT_b, (uint64_t)(3 * d), (uint64_t)d, (uint64_t)(3 * d), GP_BK, dj))) return rc;
  const int smem = (int)(3LL * dj * d * 2 + (int64_t)stages * GP_A_BYTES + 128LL * (3 * dj + 1) * 4 + 2048);
#define ARK_GP_BWD(DJ, ST)                                                                                              \
  if (dj == DJ && stages == ST) {                                                                                       \
    const int cs = (bt0 >= GP_BM) ? pick_cluster(gru_persist_bwd_kernel<DJ, ST, 8>, gru_persist_bwd_kernel<DJ, ST, 4>, gru_persist_bwd_kernel<DJ, ST, 2>, grid, smem) : 1; \
    if ((rc = make_tmap_2d_bf16(&tmA, dgh_b, (uint64_t)(3 * d), (uint64_t)N, (uint64_t)(3 * d), GP_BK, GP_BM / cs))) return rc; \
    if (cs == 8) return launch_coop(gru_persist_bwd_kernel<DJ, ST, 8>, tmA, tmW, prm, grid, smem, s, "gru_persist_bwd", 8); \
    if (cs == 4) return launch_coop(gru_persist_bwd_kernel<DJ, ST, 4>, tmA, tmW, prm, grid, smem, s, "gru_persist_bwd", 4); \
    if (cs == 2) return launch_coop(gru_persist_bwd_kernel<DJ, ST, 2>, tmA, tmW, prm, grid, smem, s, "gru_persist_bwd", 2); \
    return launch_coop(gru_persist_bwd_kernel<DJ, ST, 1>, tmA, tmW, prm, grid, smem, s, "gru_persist_bwd", 1);          \
  }
  ARK_GP_BWD(16, 6); ARK_GP_BWD(16, 5); ARK_GP_BWD(16, 4); ARK_GP_BWD(16, 3); ARK_GP_BWD(16, 2);
  ARK_GP_BWD(32, 4); ARK_GP_BWD(32, 3); ARK_GP_BWD(32, 2);
  ARK_GP_BWD(64, 4); ARK_GP_BWD(64, 3); ARK_GP_BWD(64, 2);
#undef ARK_GP_BWD
  return fail(ARK_E_SHAPE, "gru_persist_bwd: no kernel instance for dj=%d stages=%d", dj, stages);
}
