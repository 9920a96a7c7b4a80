// K3: bf16 tensor-core GEMM for sm_100a —  C[M,N] = epi(A[M,K] . B[N,K]^T + bias[N])
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (UMMA 128 x BN x 16,
//   fp32 accumulator in TMEM) -> tcgen05.ld -> fused epilogue (bias, GELU/tanh, aux pre-activation,
//   f32|bf16 store, optional += for weight-gradient accumulation).
// Replaces the reference's nn.Linear / F.linear calls and their autograd GEMMs
// (kgvae/model/models.py:36,43-44,61-62,120,128,139,142): every dense contraction of the ELBO step.
// Operand majors: K-major (contraction contiguous) and MN-major (stored [K, M|N]) are both consumed
// directly, so dX = dY.W and dW = dY^T.X need no transposed copies in HBM.
//
// Warp roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM owner + MMA issuer
// (one elected lane), warps 2..9 = epilogue (TMEM lane quadrant = warp % 4, column half = (warp-2)/4).  One output tile per CTA;
// 96 KB of smem per CTA lets two CTAs share an SM so one tile's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "gemm_tc.cuh"
#include <string.h>
#include <stdlib.h>

namespace ark {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int TC_EPI_THREADS = 256;              // 8 epilogue warps: 2 per TMEM lane quadrant (column halves)
constexpr int TC_THREADS = 64 + TC_EPI_THREADS;  // + warp 0 (TMA producer) + warp 1 (TMEM owner, MMA issuer)

template <int BN, int STAGES, bool PERSIST>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_LD = BN + 4;  // +4 floats: conflict-free float4 row-per-lane writes
  static constexpr int STAGING_BYTES = TC_BM * STAGING_LD * 4;
  // one-tile kernel: the epilogue staging tile aliases the (by then idle) TMA ring;
  // persistent kernel: the next tile's main loop runs during the epilogue, so staging has its own memory
  static constexpr int STAGING_OFFSET = PERSIST ? RING_BYTES : 0;
  static constexpr int BAR_OFFSET = PERSIST ? RING_BYTES + STAGING_BYTES : RING_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + BN * 4 + 1024;  // barriers, tmem ptr, bias tile, slack
  static_assert(PERSIST || STAGING_BYTES <= RING_BYTES, "epilogue staging must fit in the TMA ring");
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ float apply_act(float v, int epilogue) {
  if (epilogue == ARK_EPI_GELU) return gelu_erf(v);
  if (epilogue == ARK_EPI_TANH) return tanhf(v);
  if (epilogue == ARK_EPI_RELU) return fmaxf(v, 0.f);
  return v;
}

// PERSIST = false: one output tile per CTA (grid = #tiles), two CTAs per SM overlap each other's epilogue.
// PERSIST = true : grid = #SMs, each CTA walks tiles blockIdx.x, +gridDim.x, ...; the accumulator is double
//                  buffered in TMEM (2 x BN columns) so the MMA warp starts tile i+1 while the epilogue warps
//                  drain tile i — no per-tile launch / TMEM alloc / barrier init, no idle tensor pipe during stores.
template <int BN, bool A_MN, bool B_MN, int STAGES, bool PERSIST>
__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB,
                                                      const __grid_constant__ CUtensorMap tmC, const EpiParams ep,
                                                      const int M, const int N, const int K,
                                                      const int a_row0, const int b_row0) {
  using L = TcSmem<BN, STAGES, PERSIST>;
  constexpr int NACC = PERSIST ? 2 : 1;
  constexpr uint32_t TMEM_COLS = (NACC * BN) < 32 ? 32 : NACC * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (M + TC_BM - 1) / TC_BM, nt = (N + BN - 1) / BN;
  const int num_tiles = mt * nt;
  const int num_kb_all = (K + TC_BK - 1) / TC_BK;
  // split-K (one-tile kernel only): blockIdx.y owns k-blocks [kb0, kb0 + num_kb)
  const int kb0 = PERSIST ? 0 : (int)blockIdx.y * ep.kb_per_split;
  const int num_kb = PERSIST ? num_kb_all : min(ep.kb_per_split, num_kb_all - kb0);
  const int tile_step = PERSIST ? (int)gridDim.x : num_tiles;   // one-tile kernel: a single trip
  // swap_raster: consecutive tile indices walk the M tiles (A is the smaller operand), so a B tile is fetched
  // once while A stays L2-resident; otherwise they walk the N tiles.
  auto tile_origin = [&](int tile, int& m0, int& n0) {
    const int mi = ep.swap_raster ? tile % mt : tile / nt;
    const int ni = ep.swap_raster ? tile / mt : tile % nt;
    m0 = mi * TC_BM;
    n0 = ni * BN;
  };

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], TC_EPI_THREADS / 32);   // one arrival per epilogue warp
    }
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
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t ph = 1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step) {
        int m0, n0;
        tile_origin(tile, m0, n0);
        for (int kb = 0; kb < num_kb; ++kb, s = (s + 1 == STAGES ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
          ptx::mbar_wait(&empty_bar[s], ph);
          uint8_t* a_s = smem + s * L::STAGE_BYTES;
          uint8_t* b_s = a_s + L::A_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < TC_BM / 64; ++j)
              ptx::tma_load_2d(a_s + j * (TC_BK * 128), &tmA, &full_bar[s], a_row0 + m0 + 64 * j, (kb0 + kb) * TC_BK);
          } else {
            ptx::tma_load_2d(a_s, &tmA, &full_bar[s], (kb0 + kb) * TC_BK, a_row0 + m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(b_s + j * (TC_BK * 128), &tmB, &full_bar[s], b_row0 + n0 + 64 * j, (kb0 + kb) * TC_BK);
          } else {
            ptx::tma_load_2d(b_s, &tmB, &full_bar[s], (kb0 + kb) * TC_BK, b_row0 + n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // The issue loop of this ONE thread is the tensor pipe's feeder: measured (tools/probe_mma_rate.cu) a
      // 128 x 128 x 16 MMA retires every 64 cycles, but with the descriptors rebuilt from the address for every
      // instruction the loop issued one only every ~100.  So: descriptors of stage 0 are built once; a stage / k-step
      // is an ADD on the 14-bit (address >> 4) field (shared-memory addresses < 256 KB never carry out of it).
      // K-major: a k-step is 32 B along the swizzled 128 B row; MN-major: 16 k-rows of 128 B = 2048 B.
      const uint32_t a_addr0 = ptx::smem_u32(smem);
      const uint64_t adesc0 = A_MN ? ptx::make_smem_desc_sw128(a_addr0, TC_BK * 128, 1024) : ptx::make_smem_desc_sw128(a_addr0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? ptx::make_smem_desc_sw128(a_addr0 + L::A_BYTES, TC_BK * 128, 1024)
                                   : ptx::make_smem_desc_sw128(a_addr0 + L::A_BYTES, 16, 1024);
      constexpr uint64_t A_STEP = (A_MN ? 2048 : 32) >> 4, B_STEP = (B_MN ? 2048 : 32) >> 4, S_STEP = L::STAGE_BYTES >> 4;
      int lt = 0, s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step, ++lt) {
        const int acc = PERSIST ? (lt & 1) : 0;
        if (PERSIST) {   // the epilogue must have drained this accumulator (2 tiles ago)
          ptx::mbar_wait(&tmem_empty_bar[acc], ((lt >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)s * S_STEP, bd = bdesc0 + (uint64_t)s * S_STEP;
          ptx::umma_f16(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 1; kk < TC_BK / 16; ++kk) ptx::umma_f16_acc(d_tmem, ad + kk * A_STEP, bd + kk * B_STEP, idesc);
          ptx::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        ptx::umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // Phase 1: each warp drains its 32 TMEM lanes (tile rows), applies bias / activation in registers and parks
    // the finished values in shared memory.  Phase 2: the 128 epilogue threads write the tile out
    // ROW-CONTIGUOUSLY — a warp stores 512 contiguous bytes per instruction.
    const int q = warp & 3;                 // TMEM lanes [32q, 32q+32)
    const int half = (warp - 2) >> 2;       // which half of the tile's columns this warp drains
    constexpr int LD = L::STAGING_LD;
    constexpr int HALF_N = BN / 2;
    float* stage = reinterpret_cast<float*>(smem + L::STAGING_OFFSET);
    float* bias_s = reinterpret_cast<float*>(tmem_ptr_smem + 4);   // [BN] this tile's bias, staged once per tile
    const int r_loc = q * 32 + lane;
    const int tid = threadIdx.x - 64;
    const bool vec_ok = (ep.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15) == 0) &&
                        (!ep.aux || (reinterpret_cast<uintptr_t>(ep.aux) & 15) == 0);
    const bool plain = vec_ok && ep.epilogue == ARK_EPI_NONE && !ep.aux && (ep.ldc % 8 == 0) && !ep.atomic;
    constexpr int F4_PER_ROW = BN / 4;
    constexpr int G8_PER_ROW = BN / 8;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += tile_step, ++lt) {
      int m0, n0;
      tile_origin(tile, m0, n0);
      const int acc = PERSIST ? (lt & 1) : 0;
      const bool full_tile = (m0 + TC_BM <= M) && (n0 + BN <= N);
      if constexpr (PERSIST && BN == 128) {
        if (ep.tma_store) {
          // the staging tile may still be the source of the previous tile's TMA store: its issuers wait, then all meet
          if (lane == 0 && q == 0) ptx::bulk_wait_group_read0();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      // bias for the tile's columns: ONE coalesced global load per tile, issued before the MMA wait
      if (tid < BN) bias_s[tid] = (ep.bias && n0 + tid < N && kb0 == 0) ? __ldg(ep.bias + n0 + tid) : 0.f;
      ptx::mbar_wait(&tmem_full_bar[acc], PERSIST ? ((lt >> 1) & 1) : 0);
      ptx::tc_fence_after();
      if constexpr (PERSIST && BN == 128) {
        if (ep.tma_store && plain && full_tile && ep.c_bf16) {
          // ---- full bf16 tile: TMEM -> registers (+ bias) -> bf16 -> 128B-swizzled staging -> ONE TMA store per
          // 64-column half.  A thread owns one row: its 64 columns are exactly one 128-byte swizzle row.
          asm volatile("bar.sync 1, 256;" ::: "memory");      // bias_s visible
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * HALF_N);
          uint8_t* half_sm = reinterpret_cast<uint8_t*>(stage) + half * (TC_BM * 128);
          uint8_t* row_sm = half_sm + r_loc * 128;
          const float* bs = bias_s + half * HALF_N;
          uint32_t r[HALF_N / 16][16];
#pragma unroll
          for (int c = 0; c < HALF_N / 16; ++c) ptx::tmem_ld_32x32b_x16(t_addr + (uint32_t)(c * 16), r[c]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < HALF_N / 16; ++c) {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              uint4 pk;
              const int b0 = c * 16 + h8 * 8;
              const float4 ba = *reinterpret_cast<const float4*>(bs + b0), bb = *reinterpret_cast<const float4*>(bs + b0 + 4);
              pk.x = pack_bf16x2(__uint_as_float(r[c][h8 * 8 + 0]) + ba.x, __uint_as_float(r[c][h8 * 8 + 1]) + ba.y);
              pk.y = pack_bf16x2(__uint_as_float(r[c][h8 * 8 + 2]) + ba.z, __uint_as_float(r[c][h8 * 8 + 3]) + ba.w);
              pk.z = pack_bf16x2(__uint_as_float(r[c][h8 * 8 + 4]) + bb.x, __uint_as_float(r[c][h8 * 8 + 5]) + bb.y);
              pk.w = pack_bf16x2(__uint_as_float(r[c][h8 * 8 + 6]) + bb.z, __uint_as_float(r[c][h8 * 8 + 7]) + bb.w);
              const int chunk = (c * 2 + h8) ^ (r_loc & 7);           // 128B swizzle: 16-byte chunk index ^ (row % 8)
              *reinterpret_cast<uint4*>(row_sm + chunk * 16) = pk;
            }
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);   // accumulator back to the MMA warp
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
          else asm volatile("bar.sync 3, 128;" ::: "memory");
          if (lane == 0 && q == 0) {
            ptx::tma_store_2d(&tmC, half_sm, n0 + half * HALF_N, m0);
            ptx::bulk_commit_group();
          }
          continue;
        }
      }
      // Phase 1: raw accumulators TMEM -> smem; all of this warp's loads are in flight before the single wait
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * HALF_N);
      float* sp = stage + r_loc * LD + half * HALF_N;
      {
        uint32_t r[HALF_N / 16][16];
#pragma unroll
        for (int c = 0; c < HALF_N / 16; ++c) ptx::tmem_ld_32x32b_x16(t_addr + (uint32_t)(c * 16), r[c]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < HALF_N / 16; ++c)
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<uint4*>(sp + c * 16 + i) = make_uint4(r[c][i], r[c][i + 1], r[c][i + 2], r[c][i + 3]);
      }
      if (PERSIST) {   // this warp's TMEM reads are complete: hand the accumulator back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // Phase 2: bias (+ activation, aux) + stores, row-contiguous.
      if (plain && full_tile) {
        // fast path (every large GEMM of the step): 8 columns per thread, no bounds checks, no activation
        if (ep.c_bf16) {
          uint16_t* cbase = reinterpret_cast<uint16_t*>(ep.C) + (int64_t)m0 * ep.ldc + n0;
#pragma unroll 4
          for (int item = tid; item < TC_BM * G8_PER_ROW; item += TC_EPI_THREADS) {
            const int rl = item / G8_PER_ROW, c8 = (item % G8_PER_ROW) * 8;
            const float4 a0 = *reinterpret_cast<const float4*>(stage + rl * LD + c8);
            const float4 a1 = *reinterpret_cast<const float4*>(stage + rl * LD + c8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c8 + 4);
            uint4 pk;
            pk.x = pack_bf16x2(a0.x + b0.x, a0.y + b0.y);
            pk.y = pack_bf16x2(a0.z + b0.z, a0.w + b0.w);
            pk.z = pack_bf16x2(a1.x + b1.x, a1.y + b1.y);
            pk.w = pack_bf16x2(a1.z + b1.z, a1.w + b1.w);
            *reinterpret_cast<uint4*>(cbase + (int64_t)rl * ep.ldc + c8) = pk;
          }
        } else {
          float* cbase = reinterpret_cast<float*>(ep.C) + (int64_t)m0 * ep.ldc + n0;
          const bool accum = ep.accumulate != 0;
#pragma unroll 4
          for (int item = tid; item < TC_BM * F4_PER_ROW; item += TC_EPI_THREADS) {
            const int rl = item / F4_PER_ROW, c4 = (item % F4_PER_ROW) * 4;
            float4 w = *reinterpret_cast<const float4*>(stage + rl * LD + c4);
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + c4);
            float* cp = cbase + (int64_t)rl * ep.ldc + c4;
            w.x += bv.x; w.y += bv.y; w.z += bv.z; w.w += bv.w;
            if (accum) {
              const float4 old = *reinterpret_cast<const float4*>(cp);
              w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
            }
            *reinterpret_cast<float4*>(cp) = w;
          }
        }
      } else {
        // general path: edge tiles, activations, pre-activation output
#pragma unroll 1
        for (int item = tid; item < TC_BM * F4_PER_ROW; item += TC_EPI_THREADS) {
          const int rl = item / F4_PER_ROW, c4 = item % F4_PER_ROW;
          const int64_t grow = (int64_t)m0 + rl;
          const int col = n0 + c4 * 4;
          if (grow >= M || col >= N) continue;
          float4 w = *reinterpret_cast<const float4*>(stage + rl * LD + c4 * 4);
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + c4 * 4);
          w.x += bv.x; w.y += bv.y; w.z += bv.z; w.w += bv.w;
          const int64_t o = grow * ep.ldc + col;
          const bool v4 = vec_ok && col + 4 <= N;
          if (ep.aux) {   // pre-activation, needed by the backward pass
            if (v4) {
              *reinterpret_cast<float4*>(ep.aux + o) = w;
            } else {
              const float e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (col + i < N) ep.aux[o + i] = e[i];
            }
          }
          if (ep.epilogue != ARK_EPI_NONE) {
            w.x = apply_act(w.x, ep.epilogue); w.y = apply_act(w.y, ep.epilogue);
            w.z = apply_act(w.z, ep.epilogue); w.w = apply_act(w.w, ep.epilogue);
          }
          if (ep.c_bf16) {
            uint16_t* cp = reinterpret_cast<uint16_t*>(ep.C) + o;
            if (v4) {
              uint2 pk;
              pk.x = pack_bf16x2(w.x, w.y);
              pk.y = pack_bf16x2(w.z, w.w);
              *reinterpret_cast<uint2*>(cp) = pk;
            } else {
              const float e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (col + i < N) cp[i] = f32_to_bf16_bits(e[i]);
            }
          } else {
            float* cp = reinterpret_cast<float*>(ep.C) + o;
            if (ep.atomic) {          // split-K partial sum
              if (v4) {
                red_add_v4(cp, w);
              } else {
                const float e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (col + i < N) atomicAdd(cp + i, e[i]);
              }
            } else if (v4) {
              float4 x = w;
              if (ep.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(cp);
                x.x += old.x; x.y += old.y; x.z += old.z; x.w += old.w;
              }
              *reinterpret_cast<float4*>(cp) = x;
            } else {
              const float e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (col + i < N) cp[i] = ep.accumulate ? cp[i] + e[i] : e[i];
            }
          }
        }
      }
      if (PERSIST) asm volatile("bar.sync 1, 256;" ::: "memory");   // staging tile free for the next drain
    }
    if constexpr (PERSIST && BN == 128) {
      if (ep.tma_store && lane == 0 && q == 0) ptx::bulk_wait_group0();   // the last tiles' TMA stores are complete
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN, int STAGES, bool PERSIST>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& ep, int M, int N, int K,
                     int a_row0, int b_row0, cudaStream_t s) {
  using L = TcSmem<BN, STAGES, PERSIST>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, STAGES, PERSIST>;
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return fail((int)e, "gemm_bf16_tc: smem attribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int64_t nt = (N + BN - 1) / BN, mt = (M + TC_BM - 1) / TC_BM;
  EpiParams ep2 = ep;
  ep2.swap_raster = (M < N) ? 1 : 0;
  const int64_t tiles = nt * mt;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  int splits = 1;
  // split-K: few output tiles, long K (dW = dY^T X of the recurrent / attention weights: M,N = O(d), K = N_tok).
  // Only for a plain f32 output; C is zeroed (unless accumulating) and the partial sums meet through red.add.
  if (!PERSIST && !ep.c_bf16 && ep.epilogue == ARK_EPI_NONE && !ep.aux && num_kb >= 8 &&
      (tiles * 2 <= kNumSMs || (tiles <= kNumSMs && num_kb >= 24))) {
    // up to one CTA per SM; a long K with 75..148 tiles (dX of the encoder MLP: M = batch, K = 3d) goes to two
    // co-resident CTAs per SM so that the weight panel streams from all SMs at once
    splits = (int)((tiles * 2 <= kNumSMs ? kNumSMs : 2 * kNumSMs) / tiles);
    if (splits > num_kb / 4) splits = num_kb / 4;
    if (splits < 1) splits = 1;
  } else if (!PERSIST && !ep.c_bf16 && ep.epilogue == ARK_EPI_NONE && !ep.aux && num_kb >= 128 && L::TOTAL * 2 <= 227 * 1024) {
    // WAVE QUANTISATION: a grid a little larger than the GPU (vocabulary dY on wd-articles: 156 tiles, K = 60 943) runs a
    // nearly empty second round — 53 % of the SMs' time.  Split K so that the work units fill whole rounds of the 2 x 148
    // co-resident CTA slots: pick the split count with the best units / (rounds x slots), at least 8 k-blocks per unit
    // (measured 417 -> 315 us, 743 -> 983 TFLOP/s).  Only for a long K (>= 8192): on the 25 us GRU dX / dW GEMMs of
    // syn-types (K = 2560..3072) the zero-fill + red.add epilogue cost more than the idle half round (24 -> 31 us).
    const int64_t slots = 2 * kNumSMs;
    double best = (double)tiles / (double)(((tiles + slots - 1) / slots) * slots);
    for (int sp = 2; sp <= 16 && num_kb / sp >= 8; ++sp) {
      const int64_t units = tiles * sp;
      const double eff = (double)units / (double)(((units + slots - 1) / slots) * slots);
      if (eff > best + 0.04) { best = eff; splits = sp; }
    }
  }
  ep2.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + ep2.kb_per_split - 1) / ep2.kb_per_split;
  ep2.atomic = splits > 1 ? 1 : 0;
  if (splits > 1 && !ep.accumulate) {
    cudaError_t e = cudaMemset2DAsync(ep.C, (size_t)ep.ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, s);
    if (e != cudaSuccess) return fail((int)e, "gemm_bf16_tc: split-K memset: %s", cudaGetErrorString(e));
  }
  // persistent kernel, plain bf16 output: full tiles leave through TMA stores (the thread-written epilogue took
  // ~2.6 us per 128x128 tile, longer than the main loop of a K = 512 product)
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof(tmC));
  ep2.tma_store = 0;
  if (PERSIST && BN == 128 && ep.c_bf16 && ep.epilogue == ARK_EPI_NONE && !ep.aux && ep.ldc % 8 == 0 && aligned16(ep.C)) {
    static int want = -1;
    if (want < 0) { const char* ev = getenv("ARK_GEMM_TMA_STORE"); want = ev ? atoi(ev) : 1; }
    if (want && make_tmap_2d_bf16(&tmC, ep.C, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ldc, 64, TC_BM) == 0) ep2.tma_store = 1;
  }
  dim3 grid((unsigned)(PERSIST ? (tiles < kNumSMs ? tiles : kNumSMs) : tiles), (unsigned)splits);
  kern<<<grid, TC_THREADS, L::TOTAL, s>>>(tmA, tmB, tmC, ep2, M, N, K, a_row0, b_row0);
  return launched("gemm_bf16_tc");
}

template <int BN, int STAGES, bool PERSIST>
static int dispatch_major(int a_major, int b_major, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const EpiParams& ep, int M, int N, int K, int a_row0, int b_row0, cudaStream_t s) {
  if (a_major == ARK_MAJOR_K && b_major == ARK_MAJOR_K)
    return launch_tc<BN, false, false, STAGES, PERSIST>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  if (a_major == ARK_MAJOR_K) return launch_tc<BN, false, true, STAGES, PERSIST>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  if (b_major == ARK_MAJOR_K) return launch_tc<BN, true, false, STAGES, PERSIST>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  return launch_tc<BN, true, true, STAGES, PERSIST>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
}

int tc_pick_bn(int64_t M, int64_t N) {
  // N tile: 128 by default; 64 when the grid would otherwise leave most of the 148 SMs idle
  const int64_t mt = (M + TC_BM - 1) / TC_BM;
  return (mt * ((N + 127) / 128) >= kNumSMs) ? 128 : 64;
}

int tc_make_operand_map(CUtensorMap* tm, const uint16_t* P, int major, int64_t rows, int64_t K, int64_t ld,
                        int tile_rows) {
  if (major == ARK_MAJOR_K) return make_tmap_2d_bf16(tm, P, (uint64_t)K, (uint64_t)rows, (uint64_t)ld, TC_BK, tile_rows);
  return make_tmap_2d_bf16(tm, P, (uint64_t)rows, (uint64_t)K, (uint64_t)ld, 64, TC_BK);
}

int tc_enqueue(const CUtensorMap& tmA, const CUtensorMap& tmB, int a_major, int b_major, int BN, const EpiParams& ep,
               int M, int N, int K, int a_row0, int b_row0, cudaStream_t s) {
  if (BN == 128) {
    // enough tiles to keep every SM busy for >= 2 rounds: persistent CTAs with a double-buffered accumulator
    const int64_t tiles = (int64_t)((M + TC_BM - 1) / TC_BM) * ((N + 127) / 128);
    if (tiles >= 2 * kNumSMs) return dispatch_major<128, 4, true>(a_major, b_major, tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
    return dispatch_major<128, 3, false>(a_major, b_major, tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  }
  // BN = 64 means the grid is small (< 148 tiles of 128x128).  With at most one CTA per SM and a long K the main loop
  // is bound by bytes in flight x HBM latency (measured: the encoder MLP, M = 256, K = 3072, 96 CTAs, ran 26 us with a
  // 4 x 24 KB ring = 25 B/clk per SM): such shapes get an 8-stage ring (192 KB in flight per SM)
  const int64_t tiles64 = (int64_t)((M + TC_BM - 1) / TC_BM) * ((N + 63) / 64);
  const bool splitk_ok = !ep.c_bf16 && ep.epilogue == ARK_EPI_NONE && !ep.aux;   // those keep 2 CTAs/SM + split-K
  if (tiles64 <= kNumSMs && K >= 16 * TC_BK && !splitk_ok)
    return dispatch_major<64, 8, false>(a_major, b_major, tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  return dispatch_major<64, 4, false>(a_major, b_major, tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
}

int tc_check_operands(const char* who, const void* A, int a_major, int64_t lda, const void* B, int b_major, int64_t ldb,
                      int64_t M, int64_t N, int64_t K) {
  ARK_REQUIRE(A && B, ARK_E_BADARG, "%s: null pointer", who);
  ARK_REQUIRE(M >= 0 && N >= 0 && K > 0, ARK_E_BADARG, "%s: need M,N >= 0 and K > 0", who);
  ARK_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), ARK_E_SHAPE, "%s: dims must fit int32", who);
  ARK_REQUIRE((a_major == ARK_MAJOR_K || a_major == ARK_MAJOR_MN) && (b_major == ARK_MAJOR_K || b_major == ARK_MAJOR_MN),
              ARK_E_BADARG, "%s: bad major", who);
  ARK_REQUIRE(lda >= (a_major == ARK_MAJOR_K ? K : M) && ldb >= (b_major == ARK_MAJOR_K ? K : N), ARK_E_BADARG,
              "%s: leading dimension too small", who);
  ARK_REQUIRE(aligned16(A) && aligned16(B) && lda % 8 == 0 && ldb % 8 == 0, ARK_E_ALIGN,
              "%s: TMA needs 16-byte aligned bases and lda/ldb multiples of 8 (got lda=%lld ldb=%lld)", who,
              (long long)lda, (long long)ldb);
  return 0;
}

}  // namespace ark

using namespace ark;

extern "C" int ark_gemm_bf16_tc(const uint16_t* A, int a_major, int64_t lda, const uint16_t* B, int b_major,
                                int64_t ldb, void* C, int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K,
                                const float* bias, int epilogue, int accumulate, float* aux, void* stream) {
  int rc = tc_check_operands("gemm_bf16_tc", A, a_major, lda, B, b_major, ldb, M, N, K);
  if (rc) return rc;
  ARK_REQUIRE(C && ldc >= N, ARK_E_BADARG, "gemm_bf16_tc: bad C / ldc");
  ARK_REQUIRE(c_dtype == ARK_F32 || c_dtype == ARK_BF16, ARK_E_BADARG, "gemm_bf16_tc: bad c_dtype");
  ARK_REQUIRE(!(accumulate && c_dtype != ARK_F32), ARK_E_BADARG, "gemm_bf16_tc: accumulate needs f32 C");
  ARK_REQUIRE(epilogue >= ARK_EPI_NONE && epilogue <= ARK_EPI_RELU, ARK_E_BADARG, "gemm_bf16_tc: bad epilogue");
  if (M == 0 || N == 0) return 0;
  const int BN = tc_pick_bn(M, N);
  CUtensorMap tmA, tmB;
  if ((rc = tc_make_operand_map(&tmA, A, a_major, M, K, lda, TC_BM))) return rc;
  if ((rc = tc_make_operand_map(&tmB, B, b_major, N, K, ldb, BN))) return rc;
  EpiParams ep;
  ep.C = C; ep.aux = aux; ep.bias = bias; ep.ldc = ldc;
  ep.c_bf16 = (c_dtype == ARK_BF16); ep.epilogue = epilogue; ep.accumulate = accumulate;
  return tc_enqueue(tmA, tmB, a_major, b_major, BN, ep, (int)M, (int)N, (int)K, 0, 0, (cudaStream_t)stream);
}
