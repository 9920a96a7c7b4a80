// K6 (cluster): the multi-layer GRU stack (kgvae/model/models.py:121-127,141; decoder-only :329-343) for SHORT
// batch tiles and LONG chains (wd-movies, wd-articles), where gru_wave.cu is bound by the latency of exchanging the
// recurrent state through global memory (store -> red.release -> ld.acquire poll -> TMA: ~2.3 us per step measured
// with tools/ubench_dsmem.cu, vs ~0.7-1.0 us for an all-to-all through distributed shared memory).
//
// One thread-block CLUSTER of CS = d/32 CTAs owns (layer, batch tile); CTA c owns hidden units j0 = 32c .. 32c+31.
// Operands are swapped w.r.t. gru_wave: the resident WEIGHT slice is the UMMA A operand (M = gate rows), the batch
// rows are the B operand (N = NB in {16,32,64} rows), so TMEM lanes = gate rows, TMEM columns = batch rows.  The
// weight slice is copied ONCE into TENSOR MEMORY (tcgen05.st) and the MMAs take A from TMEM (tcgen05.mma [d],[a],bdesc):
// with A in shared memory every step re-read the whole slice (4 KB per 128x16x16 MMA, ~37 cycles each); from TMEM the
// 32 MMAs of a step take ~440 cycles.  (The backward projection keeps its [32 x 3d] slice in shared memory: it does
// not fit the 512 TMEM columns.)
//
//   forward, recurrence CTA (layer k): A = rows {g*d + j0..j0+32} (g = r,z,n) of W_hh^k, resident.  Step t:
//     D[96 x NB] = W_hh_slice . h_{t-1}^T  (h_{t-1} [NB x d] bf16 lives in THIS CTA's shared memory),
//     gates = D + gi_t (input projection, from the projection CTAs below), h_t slice [NB x 32] -> bf16 ->
//     cp.async.bulk (shared::cta -> shared::cluster) into the h buffer of every CTA of the cluster, completing on
//     the peer's mbarrier.  No global memory on the recurrent chain.
//   forward, projection CTA (layer k): gi_t = W_ih_slice . u_t^T + b for every step as soon as u_t exists (token
//     embeddings for k = 0, the layer below's output rows otherwise): a pipelined stage with no recurrence; its
//     result goes through a global scratch buffer + release/acquire counter to the recurrence CTA, which prefetches
//     it into a shared-memory ring.
//   backward, recurrence CTA: dh_{t-1} needs dgh_t W_hh (K = 3d).  Instead of gathering dgh_t [NB x 3d] (3x the
//     forward exchange) every CTA multiplies ITS 96 gate columns: P_c = W_hh[rows of c, :]^T . dgh_t[:, rows of c]^T
//     ([d x NB], K = 96) and the partial sums are reduce-scattered: the [32 x NB] block of P_c that belongs to CTA
//     c' goes (bf16) into c' shared memory; c' adds the CS blocks in fp32.  Same exchange volume as forward.
//   backward, projection CTA: dx_t = dgi^{k+1}_t W_ih^{k+1} for the layer below (pipelined, K = 3d).
//
// Layers run as a wavefront exactly as in gru_wave.cu (L + nl - 1 dependent steps); the clock two layers
// synchronise on is "iterations of this batch tile done", one release/acquire counter PER CTA of a (stage, batch
// tile) (see wait_all_counters), bumped by a signaller warp so that the release fence is never on the chain.
#include <vector>
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "gru_math.cuh"
#include "gru_dev.cuh"
#include "philox.cuh"
#include <string.h>
#include <stdlib.h>

namespace ark {

constexpr int GC_MAXL = 4;
constexpr int GC_DJ = 32;      // hidden units per CTA
constexpr int GC_GS = 4;       // slots of the gi / dx prefetch ring in a recurrence CTA

struct GruClFwdParams {
  CUtensorMap tmU[GC_MAXL];    // layer input rows u^k [N,d] bf16 (u^0 = token embeddings), k-chunked box {64, NB, d/64}
  const uint16_t* wih[GC_MAXL];   // W_ih^k / W_hh^k [3d,d] bf16: every CTA copies its 96 rows into TENSOR MEMORY once
  const uint16_t* whh[GC_MAXL];   // (the UMMA A operand), so a step only streams the [NB x d] B operand from smem
  const float* b_ih[GC_MAXL];
  const float* b_hh[GC_MAXL];
  const int32_t* bt;
  const int32_t* off;
  int32_t* sync;               // [2][nl][nbt][16]: iterations finished by every recurrence / projection CTA; zeroed
  const float* h0;             // [bt0, d] fp32 or null
  float* git;                  // scratch [nl][L][nbt][CS][96][NB] fp32: gi^T slices (b_ih + b_hh(r,z) included)
  uint16_t* hp_b;              // [nl, N, d] bf16 h_prev rows (block 0 pre-filled with bf16(h0) by the caller)
  uint16_t* out_b;             // [nl, N, d] bf16 layer outputs (after dropout for k < nl-1)
  uint16_t *r, *z, *n, *ghn;   // [nl, N, d] bf16 saved gates (all null in eval mode)
  uint8_t* mask;               // [nl-1, N, d] dropout keep mask or null
  const uint64_t* offset_dev;
  uint64_t seed, offset, drop_stride;
  int64_t layer_stride;        // N * d
  float p_drop;
  int L, d, nl, nbt, S;        // S = ring slots of a projection CTA
  int swap_lbo;                // debug: swap LBO/SBO of the no-swizzle descriptors
  long long* dbg;              // optional clock64 timeline (ARK_GRU_CLUSTER_DBG=<first iteration>), else null
  int dbg_it0;
};

struct GruClBwdParams {
  CUtensorMap tmDgi[GC_MAXL];     // dgi^k [N,3d] bf16, k-chunked box {64, NB, 3d/64}   (read by projection k-1)
  const uint16_t* whhT[GC_MAXL];  // W_hh^k^T [d,3d] bf16: every recurrence CTA copies its [d x 96] slice into TMEM
  CUtensorMap tmWihT[GC_MAXL];    // W_ih^k^T [d,3d] bf16, box {64, 32}                   (read by projection k-1)
  const int32_t* bt;
  const int32_t* off;
  int32_t* sync;                  // [2][nl][nbt][16]
  const float* dy_top;            // [N, d] fp32
  const uint16_t *r, *z, *n, *ghn, *hp_b;   // [nl, N, d]
  const uint8_t* mask;            // [nl-1, N, d] or null
  uint16_t *dgi_b, *dgh_b;        // [nl, N, 3d] bf16
  float* dxt;                     // scratch [nl-1][L][nbt][CS][32][NB] fp32: dx^T slices
  float* dh0;                     // [bt0, d] fp32 zeroed, or null
  int64_t layer_stride;
  float p_drop;
  int L, d, nl, nbt, S;
  int swap_lbo;
  long long* dbg;
  int dbg_it0;
};

__device__ __forceinline__ uint64_t nosw_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int swap) {
  return swap ? ptx::make_smem_desc_nosw(addr, sbo, lbo) : ptx::make_smem_desc_nosw(addr, lbo, sbo);
}
// timeline of CTA (c = 0, bi = 0) of every (layer, stage): dbg[dir][blockIdx.z][thread role][iteration - it0][point]
constexpr int GC_DBG_ITERS = 4, GC_DBG_PTS = 16, GC_DBG_WORDS = 2 * 8 * 3 * GC_DBG_ITERS * GC_DBG_PTS;
#define GC_DBG(dir, role, it, pt)                                                                              \
  do {                                                                                                         \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (it) >= p.dbg_it0 && (it) < p.dbg_it0 + GC_DBG_ITERS)   \
      p.dbg[((((dir) * 8 + blockIdx.z) * 3 + (role)) * GC_DBG_ITERS + ((it) - p.dbg_it0)) * GC_DBG_PTS + (pt)] = clock64(); \
  } while (0)

// advance the start-address field of a shared-memory descriptor (addresses < 256 KB: no carry out of the field)
__device__ __forceinline__ uint64_t desc_adv(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int NB>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
#pragma unroll
  for (int c = 0; c < NB; c += 16) ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c, v + c);
  ptx::tmem_ld_wait();
}

// =====================================================================================================
// Thread roles of a CTA (224 threads): warp 0 = loader (TMA / bulk copies / flag waits), warp 1 = TMEM owner +
// single-thread UMMA issuer, warps 2-5 = epilogue (TMEM lane quadrant = warp & 3), warp 6 = signaller: it turns
// "all 128 epilogue threads finished the global stores of an iteration" (an mbarrier) into the gpu-scope release
// of the stage counter, so the ~1000-cycle release fence never sits on the recurrent chain.
// Epilogue work items: thread (quad = tid & 7, row = tid >> 3) owns hidden units 4*quad .. 4*quad+3 of batch rows
// row + 16*i, i < NB/16: every global access is an 8-byte vector and one Philox call covers an item.
// =====================================================================================================
constexpr int GC_THREADS = 224;

__device__ __forceinline__ void st_mask4c(uint8_t* p, const bool* keep) {
  *reinterpret_cast<uint32_t*>(p) = (keep[0] ? 1u : 0u) | (keep[1] ? 0x100u : 0u) | (keep[2] ? 0x10000u : 0u) |
                                    (keep[3] ? 0x1000000u : 0u);
}
// number of leading steps at which batch row m0 is alive (bt is non-increasing)
__device__ __forceinline__ int active_steps(const int32_t* bt, int L, int m0) {
  int lo = 0, hi = L;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bt[mid] > m0) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// Signaller protocol: every epilogue thread bumps a shared-memory counter after its global stores of an iteration;
// one thread of warp 6 turns each completed group of `per_iter` arrivals into a gpu-scope release of the stage
// counter (catching up when it falls behind), so the release fence never sits on the recurrent chain.
__device__ __forceinline__ void sig_arrive(uint32_t* ctr) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(ptx::smem_u32(ctr)) : "memory");
}
__device__ __forceinline__ void signaller_loop(uint32_t* ctr, uint32_t per_iter, uint32_t n_iters, int32_t* flag) {
  uint32_t done = 0;
  for (ptx::SpinGuard g; done < n_iters;) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(ctr)) : "memory");
    v /= per_iter;
    if (v > done) {
      red_release_add(flag, (int)(v - done));
      done = v;
      g = ptx::SpinGuard();
    } else {
      __nanosleep(64);
      if (g.expired()) { printf("arkb200: gru_cluster signaller timed out (block %d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z); __trap(); }
    }
  }
}
// Stage counters are PER CTA (one int32 per slice of a (stage, batch tile): [.][GC_MAXCS]).  A consumer needs all CS
// slices of an iteration; one summed counter would let a CTA that runs an iteration ahead (the DSMEM exchange does not
// wait for the peers' global stores) complete the count of a slower peer's iteration.
constexpr int GC_MAXCS = 16;
__device__ __forceinline__ void wait_all_counters(const int32_t* p, int n, int target) {
  for (ptx::SpinGuard g;;) {
    int v[GC_MAXCS];
#pragma unroll
    for (int c = 0; c < GC_MAXCS; ++c) {      // all loads in flight together (one 64-byte line)
      v[c] = target;
      if (c < n) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v[c]) : "l"(p + c) : "memory");
    }
    int m = v[0];
#pragma unroll
    for (int c = 1; c < GC_MAXCS; ++c) m = min(m, v[c]);
    if (m >= target) break;
    if (g.expired()) {
      printf("arkb200: gru_cluster stage counters timed out (block %d,%d,%d want %d have %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, target, m);
      __trap();
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Copy rows {g*d + j0 + lane} (g = TMEM lane quadrant q < 3) of a [3d, d] bf16 weight matrix into TMEM columns
// [0, d/2) of lanes 32q..32q+31 (A operand of kind::f16 in tensor memory: column c = K elements 2c, 2c+1).
// Called by the four epilogue warps.
__device__ __forceinline__ void load_gate_rows_to_tmem(uint32_t tmem_base, const uint16_t* W, int d, int j0, int q, int lane) {
  const uint4* src = reinterpret_cast<const uint4*>(W + ((int64_t)q * d + j0 + lane) * d);
  for (int c0 = 0; c0 < d / 2; c0 += 16) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (q < 3) v = src[c0 / 4 + i];
      r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
    ptx::tmem_st_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
  }
  ptx::tmem_st_wait();
}

// =====================================================================================================
// forward
// =====================================================================================================
template <int NB>
__global__ void __launch_bounds__(GC_THREADS, 1) gru_cluster_fwd_kernel(const __grid_constant__ GruClFwdParams p) {
  constexpr int NI = NB / 16;                 // work items per epilogue thread
  constexpr uint32_t TMEM_COLS = 512;        // weight slice (d/2 columns) + two accumulators of NB columns
  constexpr int XROW = 100;                   // staged recurrent pre-activations [NB][r 32 | z 32 | n_h 32 | pad]
  constexpr uint32_t SB = NB * 64;            // bytes of one CTA's slice of an h buffer: [4 k-groups][NB/8][8 x 16 B]
  constexpr uint32_t GB = 96 * NB * 4;        // bytes of one gi slice [NB][96] fp32
  constexpr uint32_t H_LBO = NB * 16, H_SBO = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L, nl = p.nl, nbt = p.nbt;
  // the step table (rows alive per step, first packed row of a step) is read on the recurrent chain: keep it in
  // shared memory (visible after the __syncthreads of the set-up below)
  int32_t* bt_s = reinterpret_cast<int32_t*>(smem);
  int32_t* off_s = bt_s + L;
  for (int i = threadIdx.x; i < L; i += GC_THREADS) { bt_s[i] = p.bt[i]; off_s[i] = p.off[i]; }
  smem += (2 * L * 4 + 1023) / 1024 * 1024;
  const int CS = d / GC_DJ, nkc = d / 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, bi = blockIdx.y;
  const bool is_proj = (int)blockIdx.z >= nl;
  const int k = is_proj ? (int)blockIdx.z - nl : (int)blockIdx.z;
  const int j0 = c * GC_DJ, m0 = bi * NB;
  int32_t* rec_done = p.sync + (k * nbt + bi) * GC_MAXCS;            // [CS] iterations finished by each recurrence CTA
  int32_t* proj_done = p.sync + ((nl + k) * nbt + bi) * GC_MAXCS;    // ... by each projection CTA
  const uint16_t* W = is_proj ? p.wih[k] : p.whh[k];
  const uint32_t A_COLS = (uint32_t)d / 2;    // TMEM columns of the resident weight slice [128 lanes x d bf16]
  const int q = warp & 3;                     // TMEM lane quadrant of an epilogue warp
  const int tid = threadIdx.x - 64;           // epilogue thread index (warps 2..5)
  const int n_steps = active_steps(p.bt, L, m0);

  if (is_proj) {
    // ------------------------------------------------------------------------------ projection CTA
    const int S = p.S;
    const uint32_t slot_bytes = (uint32_t)nkc * NB * 128;
    uint8_t* ring = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)S * slot_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = full_bar + S;
    uint64_t* tmem_full = empty_bar + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* sig_ctr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint32_t* tmem_ptr_smem = sig_ctr + 1;
    if (threadIdx.x == 0) {
      ptx::prefetch_tmap(&p.tmU[k]);
      for (int s = 0; s < S; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 128); }
      *sig_ctr = 0;
      ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp >= 2 && warp < 6) load_gate_rows_to_tmem(tmem_base, W, d, j0, q, lane);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    ptx::cluster_sync_all();
    if (warp == 0) {
      if (ptx::elect_one()) {
        const int32_t* below = p.sync + ((k - 1) * nbt + bi) * GC_MAXCS;
        for (int t = 0; t < n_steps; ++t) {
          GC_DBG(0, 0, t, 0);
          if (k > 0) { wait_all_counters(below, CS, t + 1); fence_proxy_async_all(); }
          GC_DBG(0, 0, t, 1);
          const int s = t % S;
          ptx::mbar_wait(&empty_bar[s], ((t / S) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[s], slot_bytes);
          ptx::tma_load_3d(ring + (size_t)s * slot_bytes, &p.tmU[k], &full_bar[s], 0, off_s[t] + m0, 0);
          GC_DBG(0, 0, t, 2);
        }
      }
    } else if (warp == 1) {
      if (ptx::elect_one()) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, NB, 0, 0);
        const uint32_t r_addr = ptx::smem_u32(ring);
        for (int t = 0; t < n_steps; ++t) {
          const int s = t % S, par = t & 1;
          if (t >= 2) ptx::mbar_wait(&tmem_empty[par], ((t >> 1) - 1) & 1);
          ptx::mbar_wait(&full_bar[s], (t / S) & 1);
          ptx::tc_fence_after();
          GC_DBG(0, 1, t, 0);
          const uint32_t acc = tmem_base + A_COLS + (uint32_t)(par * NB);
          uint32_t a_tm = tmem_base;
          uint64_t bdesc = ptx::make_smem_desc_sw128(r_addr + s * slot_bytes, 16, 1024);
          for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16_ts(acc, a_tm + kk * 8, desc_adv(bdesc, kk * 32), idesc, (kc | kk) != 0 ? 1u : 0u);
            a_tm += 32;
            bdesc = desc_adv(bdesc, NB * 128);
          }
          ptx::umma_commit(&empty_bar[s]);
          ptx::umma_commit(&tmem_full[par]);
          GC_DBG(0, 1, t, 1);
        }
      }
    } else if (warp < 6) {
      float bias = 0.f;
      if (q < 3) {
        const int jr = q * d + j0 + lane;
        bias = p.b_ih[k][jr] + (q < 2 ? p.b_hh[k][jr] : 0.f);
      }
      for (int t = 0; t < n_steps; ++t) {
        const int par = t & 1;
        if (tid == 0) GC_DBG(0, 2, t, 0);
        ptx::mbar_wait(&tmem_full[par], (t >> 1) & 1);
        ptx::tc_fence_after();
        if (tid == 0) GC_DBG(0, 2, t, 1);
        if (q < 3) {
          uint32_t v[NB];
          tmem_ld_cols<NB>(tmem_base + ((uint32_t)(q * 32) << 16) + A_COLS + (uint32_t)(par * NB), v);
          // gi slice [NB][96]: lanes of a warp write 128 contiguous bytes per batch row
          float* dst = p.git + ((((int64_t)k * L + t) * nbt + bi) * CS + c) * 96 * NB + q * 32 + lane;
#pragma unroll
          for (int b = 0; b < NB; ++b) dst[b * 96] = __uint_as_float(v[b]) + bias;
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tmem_empty[par]);     // all 128 epilogue threads arrive
        // The signaller counts arrivals in aggregate (128 per step).  Without this barrier the four warps could drift a
        // step apart (the accumulator is double buffered), three fast warps' arrivals for step t+1 would complete the
        // count of step t and the stage counter would announce a gi slice whose last 32 rows are not stored yet.  With
        // it, an arrival for step t+1 implies that every thread's rows of step t are stored.
        epi_bar_sync();
        sig_arrive(sig_ctr);
        if (tid == 0) GC_DBG(0, 2, t, 2);
      }
    } else if (warp == 6) {
      if (ptx::elect_one()) signaller_loop(sig_ctr, 128, (uint32_t)n_steps, proj_done + c);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    ptx::cluster_sync_all();
    return;
  }

  // -------------------------------------------------------------------------------- recurrence CTA
  const uint32_t HB = (uint32_t)NB * d * 2;    // one h buffer
  uint8_t* hbuf = smem;                        // [2][HB]
  uint8_t* send = hbuf + 2 * HB;               // [2][SB]
  uint8_t* gi_sm = send + 2 * SB;              // [GC_GS][GB]
  float* xs = reinterpret_cast<float*>(gi_sm + GC_GS * GB);   // [NB][XROW]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + NB * XROW);
  uint64_t* hbar = bars;                       // [2]
  uint64_t* gi_full = hbar + 2;                // [GC_GS]
  uint64_t* gi_empty = gi_full + GC_GS;        // [GC_GS]
  uint64_t* tmem_full = gi_empty + GC_GS;      // [2]
  uint32_t* sig_ctr = reinterpret_cast<uint32_t*>(tmem_full + 2);
  uint32_t* tmem_ptr_smem = sig_ctr + 1;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&hbar[0], 1);
    ptx::mbar_init(&hbar[1], 1);
    for (int s = 0; s < GC_GS; ++s) { ptx::mbar_init(&gi_full[s], 1); ptx::mbar_init(&gi_empty[s], 1); }
    ptx::mbar_init(&tmem_full[0], 1);
    ptx::mbar_init(&tmem_full[1], 1);
    *sig_ctr = 0;
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp >= 2 && warp < 6) load_gate_rows_to_tmem(tmem_base, W, d, j0, q, lane);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::cluster_sync_all();                     // every peer's mbarriers exist before the first remote complete_tx

  if (warp == 0) {
    if (ptx::elect_one()) {
      for (int t = 0; t < n_steps; ++t) {
        const int s = t % GC_GS;
        GC_DBG(0, 0, t, 0);
        ptx::mbar_wait(&gi_empty[s], ((t / GC_GS) & 1) ^ 1);
        GC_DBG(0, 0, t, 1);
        wait_all_counters(proj_done, CS, t + 1);
        fence_proxy_async_all();
        GC_DBG(0, 0, t, 2);
        ptx::mbar_arrive_expect_tx(&gi_full[s], GB);
        ptx::bulk_copy_g2s(gi_sm + s * GB, p.git + ((((int64_t)k * L + t) * nbt + bi) * CS + c) * 96 * NB, GB, &gi_full[s]);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, NB, 0, 0);
      const uint32_t h_addr = ptx::smem_u32(hbuf);
      for (int t = 0; t < n_steps; ++t) {
        const int par = t & 1;
        GC_DBG(0, 1, t, 0);
        ptx::mbar_wait_cluster(&hbar[par], (t >> 1) & 1);     // all CS slices of h_{t-1} landed in hbuf[par]
        ptx::tc_fence_after();
        GC_DBG(0, 1, t, 1);
        const uint32_t acc = tmem_base + A_COLS + (uint32_t)(par * NB);
        uint32_t a_tm = tmem_base;
        uint64_t bdesc = nosw_desc(h_addr + par * HB, H_LBO, H_SBO, p.swap_lbo);
        for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::umma_f16_ts(acc, a_tm + kk * 8, desc_adv(bdesc, kk * 2 * H_LBO), idesc, (kc | kk) != 0 ? 1u : 0u);
          a_tm += 32;
          bdesc = desc_adv(bdesc, 8 * H_LBO);
        }
        ptx::umma_commit(&tmem_full[par]);
        GC_DBG(0, 1, t, 2);
      }
    }
  } else if (warp == 6) {
    if (ptx::elect_one()) signaller_loop(sig_ctr, 128, (uint32_t)n_steps, rec_done + c);
  } else {
    const uint32_t me = (uint32_t)c;
    const int64_t LS = p.layer_stride;
    uint16_t* hp_k = p.hp_b + (int64_t)k * LS;
    uint16_t* out_k = p.out_b + (int64_t)k * LS;
    uint16_t* r_k = p.r ? p.r + (int64_t)k * LS : nullptr;
    uint16_t* z_k = p.r ? p.z + (int64_t)k * LS : nullptr;
    uint16_t* n_k = p.r ? p.n + (int64_t)k * LS : nullptr;
    uint16_t* g_k = p.r ? p.ghn + (int64_t)k * LS : nullptr;
    const bool drop = p.p_drop > 0.f && k < nl - 1;
    uint8_t* mask_k = (drop && p.mask) ? p.mask + (int64_t)k * LS : nullptr;
    const float scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
    const uint64_t ctr0 = p.offset + (p.offset_dev ? *p.offset_dev : 0ull) + (uint64_t)k * p.drop_stride;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const float bhn = (q == 2) ? p.b_hh[k][2 * d + j0 + lane] : 0.f;
    const int quad = tid & 7, row0 = tid >> 3;  // item i: batch row row0 + 16 i, hidden units 4 quad .. 4 quad + 3
    // byte offset of (row bl, unit 4*quad) inside a slice [4 k-groups][NB/8][8 rows x 16 B]
    auto slice_off = [&](int bl) { return (uint32_t)((quad >> 1) * (NB * 16) + (bl >> 3) * 128 + (bl & 7) * 16 + (quad & 1) * 8); };
    // loop-invariant remote addresses of the bulk copies this thread issues (tid < CS: to peer (tid + me) % CS)
    // (peer index pidx = 4 * (tid & 31) + (tid >> 5) for the first CS/4 lanes of each warp: the copies are issued by
    // four warps in parallel)
    uint32_t dst_h[2] = {0, 0}, dst_bar[2] = {0, 0};
    const int pidx = 4 * (tid & 31) + (tid >> 5);
    const bool sender = pidx < CS;
    if (sender) {
      const uint32_t peer = ((uint32_t)pidx + me) % (uint32_t)CS;
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        dst_h[par] = ptx::mapa_u32(ptx::smem_u32(hbuf + par * HB + me * SB), peer);
        dst_bar[par] = ptx::mapa_u32(ptx::smem_u32(&hbar[par]), peer);
      }
    }
    auto send_slice = [&](int par) {           // called by all 128 epilogue threads after send[par] is written
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      epi_bar_sync();
      if (tid == 0) ptx::mbar_arrive_expect_tx(&hbar[par], (uint32_t)CS * SB);
      if (sender) ptx::bulk_copy_s2c(dst_h[par], ptx::smem_u32(send + par * SB), SB, dst_bar[par]);
    };
    float hreg[NI][4];
    const int bt0 = bt_s[0];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int bl = row0 + 16 * i, b = m0 + bl;
      float4 hv = make_float4(0, 0, 0, 0);
      if (p.h0 && b < bt0) hv = *reinterpret_cast<const float4*>(p.h0 + (int64_t)b * d + j0 + quad * 4);
      hreg[i][0] = hv.x; hreg[i][1] = hv.y; hreg[i][2] = hv.z; hreg[i][3] = hv.w;
      uint2 w;
      w.x = pack_bf16x2(hv.x, hv.y);
      w.y = pack_bf16x2(hv.z, hv.w);
      *reinterpret_cast<uint2*>(send + slice_off(bl)) = w;
    }
    send_slice(0);
    for (int t = 0; t < n_steps; ++t) {
      const int Bt = bt_s[t];
      const int Bn = (t + 1 < L) ? bt_s[t + 1] : 0;
      const bool next = t + 1 < n_steps;
      const int par = t & 1, s = t % GC_GS;
      const float* gi = reinterpret_cast<const float*>(gi_sm + s * GB);   // [NB][96]
      if (tid == 0) GC_DBG(0, 2, t, 0);
      ptx::mbar_wait(&gi_full[s], (t / GC_GS) & 1);
      ptx::mbar_wait(&tmem_full[par], (t >> 1) & 1);
      ptx::tc_fence_after();
      if (tid == 0) GC_DBG(0, 2, t, 1);
      if (q < 3) {                              // TMEM lane = gate row (q, lane), column = batch row
        uint32_t v[NB];
        tmem_ld_cols<NB>(tmem_base + ((uint32_t)(q * 32) << 16) + A_COLS + (uint32_t)(par * NB), v);
        float* dst = xs + q * 32 + lane;
#pragma unroll
        for (int b = 0; b < NB; ++b) dst[b * XROW] = __uint_as_float(v[b]) + bhn;
      }
      ptx::tc_fence_before();
      epi_bar_sync();
      if (tid == 0) GC_DBG(0, 2, t, 2);
      float o_r[NI][4], o_z[NI][4], o_n[NI][4], o_g[NI][4];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int bl = row0 + 16 * i;
        const float4 hr = *reinterpret_cast<const float4*>(xs + bl * XROW + quad * 4);
        const float4 hz = *reinterpret_cast<const float4*>(xs + bl * XROW + 32 + quad * 4);
        const float4 hn = *reinterpret_cast<const float4*>(xs + bl * XROW + 64 + quad * 4);
        const float4 ir = *reinterpret_cast<const float4*>(gi + bl * 96 + quad * 4);
        const float4 iz = *reinterpret_cast<const float4*>(gi + bl * 96 + 32 + quad * 4);
        const float4 in = *reinterpret_cast<const float4*>(gi + bl * 96 + 64 + quad * 4);
        const float ar[4] = {hr.x + ir.x, hr.y + ir.y, hr.z + ir.z, hr.w + ir.w};
        const float az[4] = {hz.x + iz.x, hz.y + iz.y, hz.z + iz.z, hz.w + iz.w};
        const float an[4] = {in.x, in.y, in.z, in.w};
        const float gn[4] = {hn.x, hn.y, hn.z, hn.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const GruFwd o = gru_fwd_math_fast(ar[e], az[e], an[e], 0.f, 0.f, gn[e], hreg[i][e]);
          hreg[i][e] = o.h;
          o_r[i][e] = o.r; o_z[i][e] = o.z; o_n[i][e] = o.n; o_g[i][e] = o.ghn;
        }
        if (next) {
          uint2 w;
          w.x = pack_bf16x2(hreg[i][0], hreg[i][1]);
          w.y = pack_bf16x2(hreg[i][2], hreg[i][3]);
          *reinterpret_cast<uint2*>(send + ((t + 1) & 1) * SB + slice_off(bl)) = w;
        }
      }
      if (tid == 0) GC_DBG(0, 2, t, 3);
      if (next) {
        send_slice((t + 1) & 1);               // (contains the barrier after which gi[s] / xs are free again)
      } else {
        epi_bar_sync();
      }
      if (tid == 0) ptx::mbar_arrive(&gi_empty[s]);
      if (tid == 0) GC_DBG(0, 2, t, 4);
      // ---- off the recurrent chain: saved tensors, layer output (+ dropout); the signaller releases the counter
      const int64_t base = (int64_t)off_s[t] + m0;
      const int64_t base_n = next ? (int64_t)off_s[t + 1] + m0 : 0;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int bl = row0 + 16 * i, b = m0 + bl;
        if (b >= Bt) continue;
        const int64_t o = (base + bl) * d + j0 + quad * 4;
        float hv[4] = {hreg[i][0], hreg[i][1], hreg[i][2], hreg[i][3]};
        if (b < Bn) st4_bf16(hp_k + (base_n + bl) * d + j0 + quad * 4, hv);
        if (r_k) {
          st4_bf16(r_k + o, o_r[i]);
          st4_bf16(z_k + o, o_z[i]);
          st4_bf16(n_k + o, o_n[i]);
          st4_bf16(g_k + o, o_g[i]);
        }
        if (drop) {   // same draw as dropout_bf16_kernel over the [N, d] output of layer k
          const uint64_t cc = ctr0 + (uint64_t)(o >> 2);
          const uint4 rn = philox4x32_10(make_uint4((uint32_t)cc, (uint32_t)(cc >> 32), 0u, 0u), key);
          const uint32_t rr[4] = {rn.x, rn.y, rn.z, rn.w};
          bool keep[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            keep[e] = (float)(rr[e] >> 8) * (1.f / 16777216.f) >= p.p_drop;
            hv[e] = keep[e] ? bf16_bits_to_f32(f32_to_bf16_bits(hv[e])) * scale : 0.f;
          }
          if (mask_k) st_mask4c(mask_k + o, keep);
        }
        st4_bf16(out_k + o, hv);
      }
      sig_arrive(sig_ctr);
      if (tid == 0) GC_DBG(0, 2, t, 5);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  ptx::cluster_sync_all();                     // no CTA leaves while a peer's bulk copy may still be reading / writing it
}

// =====================================================================================================
// backward through time
// =====================================================================================================
template <int NB>
__global__ void __launch_bounds__(GC_THREADS, 1) gru_cluster_bwd_kernel(const __grid_constant__ GruClBwdParams p) {
  constexpr int NI = NB / 16;
  constexpr int BQ = NB / 4;                  // batch rows per (unit, quarter) cell of a partial-sum block
  constexpr int RT = 36;
  constexpr uint32_t SB = NB * 64;            // one (source, destination) block of partial sums: [4][32 units][BQ] bf16
  constexpr uint32_t DXB = 32 * NB * 4;       // one dx slice [NB][32] fp32
  constexpr uint32_t B_LBO = NB * 16, B_SBO = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L, nl = p.nl, nbt = p.nbt;
  // the step table (rows alive per step, first packed row of a step) is read on the recurrent chain: keep it in
  // shared memory (visible after the __syncthreads of the set-up below)
  int32_t* bt_s = reinterpret_cast<int32_t*>(smem);
  int32_t* off_s = bt_s + L;
  for (int i = threadIdx.x; i < L; i += GC_THREADS) { bt_s[i] = p.bt[i]; off_s[i] = p.off[i]; }
  smem += (2 * L * 4 + 1023) / 1024 * 1024;
  const int CS = d / GC_DJ, MT = d / 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, bi = blockIdx.y;
  const bool is_proj = (int)blockIdx.z >= nl;
  const int k = is_proj ? (int)blockIdx.z - nl : (int)blockIdx.z;    // projection k feeds layer k from layer k+1
  const int j0 = c * GC_DJ, m0 = bi * NB;
  int32_t* rec_done = p.sync + (k * nbt + bi) * GC_MAXCS;            // [CS] iterations finished by each recurrence CTA
  int32_t* proj_done = p.sync + ((nl + k) * nbt + bi) * GC_MAXCS;    // ... by each projection CTA
  const int q = warp & 3;
  const int tid = threadIdx.x - 64;
  const int t_first = active_steps(p.bt, L, m0) - 1;      // last step at which this batch tile is alive

  if (is_proj) {
    // ------------------------------------------------------------------------------ projection CTA
    constexpr uint32_t TMEM_COLS = 2 * NB < 32 ? 32 : 2 * NB;
    const int nkc = 3 * d / 64, S = p.S;
    const uint32_t w_bytes = (uint32_t)nkc * 4096;           // [nkc][32 rows x 128 B], 128B swizzle
    const uint32_t slot_bytes = (uint32_t)nkc * NB * 128;
    uint8_t* w_sm = smem;                                    // the 128-row UMMA window of the last chunks overshoots
    uint8_t* ring = w_sm + w_bytes;                          // up to 12 KB into the ring (in-allocation)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)S * slot_bytes);
    uint64_t* w_bar = bars;
    uint64_t* full_bar = bars + 1;
    uint64_t* empty_bar = full_bar + S;
    uint64_t* tmem_full = empty_bar + S;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* sig_ctr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint32_t* tmem_ptr_smem = sig_ctr + 1;
    const CUtensorMap* tmA = &p.tmDgi[k + 1];
    if (threadIdx.x == 0) {
      ptx::prefetch_tmap(tmA);
      ptx::mbar_init(w_bar, 1);
      for (int s = 0; s < S; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 32); }
      *sig_ctr = 0;
      ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    ptx::cluster_sync_all();
    if (warp == 0) {
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(w_bar, w_bytes);
        for (int kc = 0; kc < nkc; ++kc) ptx::tma_load_2d(w_sm + kc * 4096, &p.tmWihT[k + 1], w_bar, kc * 64, j0);
        const int32_t* above = p.sync + ((k + 1) * nbt + bi) * GC_MAXCS;
        for (int it = 0; it <= t_first; ++it) {
          const int t = t_first - it, s = it % S;
          GC_DBG(1, 0, it, 0);
          wait_all_counters(above, CS, it + 1);             // every slice of dgi^{k+1}_t is in global memory
          fence_proxy_async_all();
          GC_DBG(1, 0, it, 1);
          ptx::mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
          GC_DBG(1, 0, it, 2);
          ptx::mbar_arrive_expect_tx(&full_bar[s], slot_bytes);
          ptx::tma_load_3d(ring + (size_t)s * slot_bytes, tmA, &full_bar[s], 0, off_s[t] + m0, 0);
        }
      }
    } else if (warp == 1) {
      if (ptx::elect_one()) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(128, NB, 0, 0);
        ptx::mbar_wait(w_bar, 0);
        const uint32_t w_addr = ptx::smem_u32(w_sm), r_addr = ptx::smem_u32(ring);
        for (int it = 0; it <= t_first; ++it) {
          const int s = it % S, par = it & 1;
          GC_DBG(1, 1, it, 0);
          if (it >= 2) ptx::mbar_wait(&tmem_empty[par], ((it >> 1) - 1) & 1);
          ptx::mbar_wait(&full_bar[s], (it / S) & 1);
          ptx::tc_fence_after();
          GC_DBG(1, 1, it, 1);
          const uint32_t acc = tmem_base + (uint32_t)(par * NB);
          uint64_t adesc = ptx::make_smem_desc_sw128(w_addr, 16, 1024);
          uint64_t bdesc = ptx::make_smem_desc_sw128(r_addr + s * slot_bytes, 16, 1024);
          for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16(acc, desc_adv(adesc, kk * 32), desc_adv(bdesc, kk * 32), idesc, (kc | kk) != 0 ? 1u : 0u);
            adesc = desc_adv(adesc, 4096);
            bdesc = desc_adv(bdesc, NB * 128);
          }
          ptx::umma_commit(&empty_bar[s]);
          ptx::umma_commit(&tmem_full[par]);
          GC_DBG(1, 1, it, 2);
        }
      }
    } else if (warp < 6) {
      if (q == 0) {                              // TMEM lanes 0..31 = the 32 output units of this CTA
        for (int it = 0; it <= t_first; ++it) {
          const int t = t_first - it, par = it & 1;
          if (lane == 0) GC_DBG(1, 2, it, 0);
          ptx::mbar_wait(&tmem_full[par], (it >> 1) & 1);
          ptx::tc_fence_after();
          if (lane == 0) GC_DBG(1, 2, it, 1);
          uint32_t v[NB];
          tmem_ld_cols<NB>(tmem_base + (uint32_t)(par * NB), v);
          float* dst = p.dxt + ((((int64_t)k * L + t) * nbt + bi) * CS + c) * 32 * NB + lane;   // [NB][32]
#pragma unroll
          for (int b = 0; b < NB; ++b) dst[b * 32] = __uint_as_float(v[b]);
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tmem_empty[par]);
          sig_arrive(sig_ctr);
          if (lane == 0) GC_DBG(1, 2, it, 2);
        }
      }
    } else if (warp == 6) {
      if (ptx::elect_one()) signaller_loop(sig_ctr, 32, (uint32_t)(t_first + 1), proj_done + c);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    ptx::cluster_sync_all();
    return;
  }

  // -------------------------------------------------------------------------------- recurrence CTA
  constexpr uint32_t TMEM_COLS_R = 512;   // MT <= 4 weight tiles of 48 columns + MT accumulators of NB columns
  const bool top = (k == nl - 1);
  const uint32_t ACC0 = (uint32_t)MT * 48;                   // first accumulator column
  uint8_t* bop = smem;                                       // dgh_c^T [NB x 96] bf16, no swizzle: [12][NB/8][128 B]
  uint8_t* stage = bop + NB * 192;                           // [2][CS][SB]
  uint8_t* recv = stage + 2 * CS * SB;                       // [2][CS][SB]
  uint8_t* dx_sm = recv + 2 * CS * SB;                       // [GC_GS][DXB]
  float* recT = reinterpret_cast<float*>(dx_sm + GC_GS * DXB);   // [NB][RT]: summed partials, transposed for the gate math
  uint64_t* bars = reinterpret_cast<uint64_t*>(recT + NB * RT);
  uint64_t* rbar = bars;                                     // [2]
  uint64_t* bop_full = rbar + 2;
  uint64_t* tmem_full = bop_full + 1;
  uint64_t* dx_full = tmem_full + 1;                         // [GC_GS]
  uint64_t* dx_empty = dx_full + GC_GS;                      // [GC_GS]
  uint32_t* sig_ctr = reinterpret_cast<uint32_t*>(dx_empty + GC_GS);
  uint32_t* tmem_ptr_smem = sig_ctr + 1;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&rbar[0], 1);
    ptx::mbar_init(&rbar[1], 1);
    ptx::mbar_init(bop_full, 1);
    ptx::mbar_init(tmem_full, 1);
    for (int s = 0; s < GC_GS; ++s) { ptx::mbar_init(&dx_full[s], 1); ptx::mbar_init(&dx_empty[s], 1); }
    *sig_ctr = 0;
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS_R); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp >= 2 && warp < 6) {
    // A operand in tensor memory: M-tile mt = output units mt*128 .. +128 (lane = unit), K = the 96 gate rows of this
    // CTA (column 16 g + c = rows g*d + j0 + 2c, 2c+1 of W_hh, i.e. columns of W_hh^T)
    for (int mt = 0; mt < MT; ++mt) {
      const uint16_t* row = p.whhT[k] + (int64_t)(mt * 128 + q * 32 + lane) * (3 * d) + j0;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 v = *reinterpret_cast<const uint4*>(row + (int64_t)g * d + i * 8);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        ptx::tmem_st_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 48 + g * 16), r);
      }
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::cluster_sync_all();

  if (warp == 0) {
    if (ptx::elect_one()) {
      if (!top) {
        for (int it = 0; it <= t_first; ++it) {
          const int t = t_first - it, s = it % GC_GS;
          GC_DBG(1, 0, it, 0);
          ptx::mbar_wait(&dx_empty[s], ((it / GC_GS) & 1) ^ 1);
          GC_DBG(1, 0, it, 1);
          wait_all_counters(proj_done, CS, it + 1);
          fence_proxy_async_all();
          GC_DBG(1, 0, it, 2);
          ptx::mbar_arrive_expect_tx(&dx_full[s], DXB);
          ptx::bulk_copy_g2s(dx_sm + s * DXB, p.dxt + ((((int64_t)k * L + t) * nbt + bi) * CS + c) * 32 * NB, DXB, &dx_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, NB, 0, 0);
      const uint32_t b_addr = ptx::smem_u32(bop);
      for (int it = 0; it <= t_first; ++it) {
        GC_DBG(1, 1, it, 0);
        ptx::mbar_wait(bop_full, it & 1);
        ptx::tc_fence_after();
        GC_DBG(1, 1, it, 1);
        const uint64_t bdesc = nosw_desc(b_addr, B_LBO, B_SBO, p.swap_lbo);
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int ks = 0; ks < 6; ++ks)
            ptx::umma_f16_ts(tmem_base + ACC0 + (uint32_t)(mt * NB), tmem_base + (uint32_t)(mt * 48 + ks * 8),
                             desc_adv(bdesc, ks * 2 * B_LBO), idesc, ks != 0 ? 1u : 0u);
        }
        ptx::umma_commit(tmem_full);
        GC_DBG(1, 1, it, 2);
      }
    }
  } else if (warp == 6) {
    if (ptx::elect_one()) signaller_loop(sig_ctr, 128, (uint32_t)(t_first + 1), rec_done + c);
  } else {
    const uint32_t me = (uint32_t)c;
    const int quad = tid & 7, row0 = tid >> 3;
    const int64_t d3 = 3 * (int64_t)d;
    const int64_t LS = p.layer_stride;
    const uint16_t* r_k = p.r + (int64_t)k * LS;
    const uint16_t* z_k = p.z + (int64_t)k * LS;
    const uint16_t* n_k = p.n + (int64_t)k * LS;
    const uint16_t* g_k = p.ghn + (int64_t)k * LS;
    const uint16_t* hp_k = p.hp_b + (int64_t)k * LS;
    const bool drop = !top && p.p_drop > 0.f && p.mask != nullptr;
    const uint8_t* mask_k = drop ? p.mask + (int64_t)k * LS : nullptr;
    const float scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
    uint16_t* dgi_k = p.dgi_b + (int64_t)k * LS * 3;
    uint16_t* dgh_k = p.dgh_b + (int64_t)k * LS * 3;
    uint32_t dst_r[2] = {0, 0}, dst_bar[2] = {0, 0}, src_off = 0;
    const int pidx = 4 * (tid & 31) + (tid >> 5);           // the copies are issued by four warps in parallel
    const bool sender = pidx < CS;
    if (sender) {
      const uint32_t peer = ((uint32_t)pidx + me) % (uint32_t)CS;
      src_off = peer * SB;
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        dst_r[par] = ptx::mapa_u32(ptx::smem_u32(recv + par * CS * SB + me * SB), peer);
        dst_bar[par] = ptx::mapa_u32(ptx::smem_u32(&rbar[par]), peer);
      }
    }
    float carry[NI][4];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) carry[i][e] = 0.f;
    // saved rows of the next iteration, kept RAW (bf16 bits) so that nothing waits on the loads before their use
    float4 dyp[NI];
    uint32_t mk[NI];
    uint2 sp[NI][5];
    auto prefetch = [&](int t) {
      const int Bt = bt_s[t];
      const int64_t base = (int64_t)off_s[t] + m0;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int bl = row0 + 16 * i;
        if (m0 + bl < Bt) {
          const int64_t o = (base + bl) * d + j0 + quad * 4;
          if (top) dyp[i] = *reinterpret_cast<const float4*>(p.dy_top + o);
          if (drop) mk[i] = *reinterpret_cast<const uint32_t*>(mask_k + o);
          sp[i][0] = *reinterpret_cast<const uint2*>(r_k + o);
          sp[i][1] = *reinterpret_cast<const uint2*>(z_k + o);
          sp[i][2] = *reinterpret_cast<const uint2*>(n_k + o);
          sp[i][3] = *reinterpret_cast<const uint2*>(g_k + o);
          sp[i][4] = *reinterpret_cast<const uint2*>(hp_k + o);
        }
      }
    };
    auto prefetch_far = [&](int t) {           // pull the rows of a later iteration into L2 (they come from HBM)
      const int Bt = bt_s[t];
      const int64_t base = (int64_t)off_s[t] + m0;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int bl = row0 + 16 * i;
        if (m0 + bl < Bt && (quad & 3) == 0) {  // one 32-byte sector per (row, half slice)
          const int64_t o = (base + bl) * d + j0 + quad * 4;
          if (top) { prefetch_l2(p.dy_top + o); prefetch_l2(p.dy_top + o + 8); }
          if (drop && quad == 0) prefetch_l2(mask_k + o);
          prefetch_l2(r_k + o); prefetch_l2(z_k + o); prefetch_l2(n_k + o); prefetch_l2(g_k + o); prefetch_l2(hp_k + o);
        }
      }
    };
    // byte offset of (row bl, gate g, unit 4*quad) in the B operand [12 k-groups][NB/8][8 rows x 16 B]
    auto bop_off = [&](int g, int bl) {
      return (uint32_t)((g * 4 + (quad >> 1)) * (NB * 16) + (bl >> 3) * 128 + (bl & 7) * 16 + (quad & 1) * 8);
    };
    prefetch(t_first);
    if (t_first >= 1) prefetch_far(t_first - 1);
    if (t_first >= 2) prefetch_far(t_first - 2);
    for (int it = 0; it <= t_first + 1; ++it) {
      const int t = t_first - it;                           // t = -1: gradient of the initial state
      const int Bt = bt_s[t < 0 ? 0 : t];
      const int B_next = (t + 1 <= L - 1) ? bt_s[t + 1] : 0;
      if (tid == 0) GC_DBG(1, 2, it, 0);
      const int s = it % GC_GS;
      const float* dxs = reinterpret_cast<const float*>(dx_sm + s * DXB);   // [NB][32]
      // (the projection runs ahead: this wait is normally over at once and its latency hides behind the exchange)
      if (t >= 0 && !top) ptx::mbar_wait(&dx_full[s], (it / GC_GS) & 1);
      if (it > 0) {                                         // partial sums of dgh_{t+1} W_hh from every CTA of the cluster
        const int e = it - 1;
        ptx::mbar_wait_cluster(&rbar[e & 1], (e >> 1) & 1);
        if (tid == 0) GC_DBG(1, 2, it, 1);
        // thread (unit = lane, quarter = tid >> 5) adds the CS blocks of its BQ batch rows, then the sums are
        // transposed through shared memory into the (row, 4 units) items of the gate math
        const uint8_t* rb = recv + (e & 1) * CS * SB + tid * (BQ * 2);
        float acc[4][BQ];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < BQ; ++i) acc[u][i] = 0.f;
        for (int cc = 0; cc < CS; cc += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if constexpr (BQ == 4) {
              const uint2 v = *reinterpret_cast<const uint2*>(rb + (cc + u) * SB);
              const float2 a = unpack_bf16x2(v.x), b2 = unpack_bf16x2(v.y);
              acc[u][0] += a.x; acc[u][1] += a.y; acc[u][2] += b2.x; acc[u][3] += b2.y;
            } else {
#pragma unroll
              for (int i = 0; i < BQ; i += 8) {
                const uint4 v = *reinterpret_cast<const uint4*>(rb + (cc + u) * SB + i * 2);
                const float2 a = unpack_bf16x2(v.x), b2 = unpack_bf16x2(v.y), c2 = unpack_bf16x2(v.z), e2 = unpack_bf16x2(v.w);
                acc[u][i] += a.x; acc[u][i + 1] += a.y; acc[u][i + 2] += b2.x; acc[u][i + 3] += b2.y;
                acc[u][i + 4] += c2.x; acc[u][i + 5] += c2.y; acc[u][i + 6] += e2.x; acc[u][i + 7] += e2.y;
              }
            }
          }
        }
        float* rt = recT + (tid >> 5) * BQ * RT + lane;
#pragma unroll
        for (int i = 0; i < BQ; ++i) rt[i * RT] = (acc[0][i] + acc[1][i]) + (acc[2][i] + acc[3][i]);
        epi_bar_sync();
      }
      float rec[NI][4];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        float4 v = make_float4(0, 0, 0, 0);
        if (it > 0) v = *reinterpret_cast<const float4*>(recT + (row0 + 16 * i) * RT + quad * 4);
        rec[i][0] = v.x; rec[i][1] = v.y; rec[i][2] = v.z; rec[i][3] = v.w;
      }
      if (tid == 0) GC_DBG(1, 2, it, 2);
      if (tid == 0) GC_DBG(1, 2, it, 3);
      float dar[NI][4], daz[NI][4], dan[NI][4], danr[NI][4];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int bl = row0 + 16 * i, b = m0 + bl;
#pragma unroll
        for (int e = 0; e < 4; ++e) dar[i][e] = daz[i][e] = dan[i][e] = danr[i][e] = 0.f;
        if (b < Bt) {
          float dh[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) dh[e] = (b < B_next) ? carry[i][e] + rec[i][e] : 0.f;
          if (t < 0) {
            if (p.dh0) red_add_v4(p.dh0 + (int64_t)b * d + j0 + quad * 4, make_float4(dh[0], dh[1], dh[2], dh[3]));
          } else {
            float dyv[4];
            if (top) {
              dyv[0] = dyp[i].x; dyv[1] = dyp[i].y; dyv[2] = dyp[i].z; dyv[3] = dyp[i].w;
            } else {
              const float4 x = *reinterpret_cast<const float4*>(dxs + bl * 32 + quad * 4);
              const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) dyv[e] = (drop ? (((mk[i] >> (8 * e)) & 0xffu) ? scale : 0.f) : 1.f) * xv[e];
            }
            float r[4], z[4], n[4], g[4], hp[4];
            {
              float2 a, c2;
              a = unpack_bf16x2(sp[i][0].x); c2 = unpack_bf16x2(sp[i][0].y); r[0] = a.x; r[1] = a.y; r[2] = c2.x; r[3] = c2.y;
              a = unpack_bf16x2(sp[i][1].x); c2 = unpack_bf16x2(sp[i][1].y); z[0] = a.x; z[1] = a.y; z[2] = c2.x; z[3] = c2.y;
              a = unpack_bf16x2(sp[i][2].x); c2 = unpack_bf16x2(sp[i][2].y); n[0] = a.x; n[1] = a.y; n[2] = c2.x; n[3] = c2.y;
              a = unpack_bf16x2(sp[i][3].x); c2 = unpack_bf16x2(sp[i][3].y); g[0] = a.x; g[1] = a.y; g[2] = c2.x; g[3] = c2.y;
              a = unpack_bf16x2(sp[i][4].x); c2 = unpack_bf16x2(sp[i][4].y); hp[0] = a.x; hp[1] = a.y; hp[2] = c2.x; hp[3] = c2.y;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const GruBwd w = gru_bwd_math(dh[e] + dyv[e], r[e], z[e], n[e], g[e], hp[e]);
              dar[i][e] = w.dar; daz[i][e] = w.daz; dan[i][e] = w.dan; danr[i][e] = w.dan_r;
              carry[i][e] = w.dh_prev;
            }
          }
        }
        if (t >= 0) {
          uint2 w;
          w.x = pack_bf16x2(dar[i][0], dar[i][1]); w.y = pack_bf16x2(dar[i][2], dar[i][3]);
          *reinterpret_cast<uint2*>(bop + bop_off(0, bl)) = w;
          w.x = pack_bf16x2(daz[i][0], daz[i][1]); w.y = pack_bf16x2(daz[i][2], daz[i][3]);
          *reinterpret_cast<uint2*>(bop + bop_off(1, bl)) = w;
          w.x = pack_bf16x2(danr[i][0], danr[i][1]); w.y = pack_bf16x2(danr[i][2], danr[i][3]);
          *reinterpret_cast<uint2*>(bop + bop_off(2, bl)) = w;
        }
      }
      if (t < 0) break;
      if (tid == 0) GC_DBG(1, 2, it, 4);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      epi_bar_sync();
      if (tid == 0) {
        ptx::mbar_arrive(bop_full);
        if (!top) ptx::mbar_arrive(&dx_empty[s]);
      }
      if (tid == 0) GC_DBG(1, 2, it, 5);
      if (tid == 0) GC_DBG(1, 2, it, 6);
      ptx::mbar_wait(tmem_full, it & 1);
      ptx::tc_fence_after();
      if (tid == 0) GC_DBG(1, 2, it, 7);
      uint8_t* st = stage + (it & 1) * CS * SB;
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t v[NB];
        tmem_ld_cols<NB>(tmem_base + ((uint32_t)(q * 32) << 16) + ACC0 + (uint32_t)(mt * NB), v);
        // block for CTA (mt*4 + q): [4 quarters][32 units][BQ rows] bf16; lane = unit
        uint8_t* dst = st + (mt * 4 + q) * SB + lane * (BQ * 2);
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          if constexpr (BQ == 4) {
            uint2 w;
            w.x = pack_bf16x2(__uint_as_float(v[e4 * 4]), __uint_as_float(v[e4 * 4 + 1]));
            w.y = pack_bf16x2(__uint_as_float(v[e4 * 4 + 2]), __uint_as_float(v[e4 * 4 + 3]));
            *reinterpret_cast<uint2*>(dst + e4 * 32 * (BQ * 2)) = w;
          } else {
#pragma unroll
            for (int i = 0; i < BQ; i += 8) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(v[e4 * BQ + i]), __uint_as_float(v[e4 * BQ + i + 1]));
              w.y = pack_bf16x2(__uint_as_float(v[e4 * BQ + i + 2]), __uint_as_float(v[e4 * BQ + i + 3]));
              w.z = pack_bf16x2(__uint_as_float(v[e4 * BQ + i + 4]), __uint_as_float(v[e4 * BQ + i + 5]));
              w.w = pack_bf16x2(__uint_as_float(v[e4 * BQ + i + 6]), __uint_as_float(v[e4 * BQ + i + 7]));
              *reinterpret_cast<uint4*>(dst + e4 * 32 * (BQ * 2) + i * 2) = w;
            }
          }
        }
      }
      ptx::tc_fence_before();
      if (tid == 0) GC_DBG(1, 2, it, 8);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      epi_bar_sync();
      if (tid == 0) ptx::mbar_arrive_expect_tx(&rbar[it & 1], (uint32_t)CS * SB);
      if (sender) ptx::bulk_copy_s2c(dst_r[it & 1], ptx::smem_u32(st) + src_off, SB, dst_bar[it & 1]);
      if (tid == 0) GC_DBG(1, 2, it, 9);
      // ---- off the recurrent chain (hidden behind the exchange): the rows the weight-gradient GEMMs and the
      // projection below read, then the saved rows of the next iteration
      {
        const int64_t base = (int64_t)off_s[t] + m0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int bl = row0 + 16 * i;
          if (m0 + bl < Bt) {
            const int64_t o3 = (base + bl) * d3 + j0 + quad * 4;
            st4_bf16(dgi_k + o3, dar[i]);
            st4_bf16(dgi_k + o3 + d, daz[i]);
            st4_bf16(dgi_k + o3 + 2 * d, dan[i]);
            st4_bf16(dgh_k + o3, dar[i]);
            st4_bf16(dgh_k + o3 + d, daz[i]);
            st4_bf16(dgh_k + o3 + 2 * d, danr[i]);
          }
        }
      }
      sig_arrive(sig_ctr);                  // (a release: issued BEFORE the loads below so that it does not wait for them)
      if (t - 1 >= 0) prefetch(t - 1);
      if (t - 3 >= 0) prefetch_far(t - 3);
      if (tid == 0) GC_DBG(1, 2, it, 10);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS_R);
  ptx::cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------- host
struct ClusterPlan {
  int NB, nbt, CS, S_f, S_b, smem_f, smem_b;
};

static bool plan_cluster(int64_t d, int64_t bt0, int64_t nl, int64_t L, ClusterPlan* out) {
  if (d % 128 != 0 || d < 128 || d > 512 || bt0 <= 0 || nl < 1 || nl > GC_MAXL || L < 1 || L > 4096) return false;
  const int64_t meta = (2 * L * 4 + 1023) / 1024 * 1024;   // step table in shared memory
  const int CS = (int)(d / GC_DJ);
  if (CS > 16) return false;
  const int cand[3] = {16, 32, 64};
  for (int i = 0; i < 3; ++i) {
    const int NB = cand[i];
    const int nbt = (int)((bt0 + NB - 1) / NB);
    if (2 * nl * CS * nbt > kNumSMs) continue;
    const int64_t lim = 227 * 1024 - 1024 - meta;       // alignment slack, step table
    if (d / 2 + 2 * NB > 512 || (d / 128) * (48 + NB) > 512) continue;   // tensor memory: weight slice + accumulators
    // forward (the weight slices live in tensor memory)
    const int64_t rec_f = 2 * (int64_t)NB * d * 2 + 2 * NB * 64 + (int64_t)GC_GS * 96 * NB * 4 + (int64_t)NB * 100 * 4 + 256;
    const int64_t slot_f = d / 64 * NB * 128;
    int64_t S_f = (lim - 256) / slot_f;
    if (S_f > 4) S_f = 4;
    if (S_f < 2 || rec_f > lim) continue;
    const int64_t proj_f = S_f * slot_f + 256;
    // backward
    const int64_t rec_b = NB * 192 + 4 * (int64_t)CS * NB * 64 + (int64_t)GC_GS * 32 * NB * 4 + (int64_t)NB * 36 * 4 + 256;
    const int64_t wb = 3 * d / 64 * 4096;
    const int64_t slot_b = 3 * d / 64 * NB * 128;
    int64_t S_b = (lim - wb - 256) / slot_b;
    if (S_b > 4) S_b = 4;
    if (S_b < 2 || rec_b > lim || S_b * slot_b < 12288) continue;
    const int64_t proj_b = wb + S_b * slot_b + 256;
    out->NB = NB; out->nbt = nbt; out->CS = CS; out->S_f = (int)S_f; out->S_b = (int)S_b;
    out->smem_f = (int)((rec_f > proj_f ? rec_f : proj_f) + 1024 + meta);
    out->smem_b = (int)((rec_b > proj_b ? rec_b : proj_b) + 1024 + meta);
    return true;
  }
  return false;
}

template <typename Params, typename Kern>
static int launch_cluster(Kern kern, const Params& prm, dim3 grid, int cs, int smem, cudaStream_t s, const char* who,
                          bool query_only) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail((int)e, "%s: smem attribute (%d B): %s", who, smem, cudaGetErrorString(e)); }
  if (cs > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail((int)e, "%s: non-portable cluster size: %s", who, cudaGetErrorString(e)); }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GC_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = (unsigned)cs;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  if (query_only) {
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail((int)e, "%s: cudaOccupancyMaxActiveClusters: %s", who, cudaGetErrorString(e)); }
    const long long need = (long long)grid.x * grid.y * grid.z / cs;
    if (getenv("ARK_GRU_DEBUG"))
      fprintf(stderr, "[arkb200] %s grid=(%u,%u,%u) cluster=%d smem=%d: max active clusters %d (need %lld)\n", who, grid.x,
              grid.y, grid.z, cs, smem, n, need);
    return n >= need ? 0 : fail(ARK_E_SHAPE, "%s: only %d of %lld clusters can be co-resident", who, n, need);
  }
  // ARK_GRU_CLUSTER_NO_COOP=1 drops the cooperative attribute (Nsight Compute refuses cooperative cluster launches);
  // co-residency was established by the occupancy query of ark_gru_cluster_supported
  static int no_coop = -1;
  if (no_coop < 0) {
    const char* ev = getenv("ARK_GRU_CLUSTER_NO_COOP");
    no_coop = ev ? atoi(ev) : 0;
    if (!ev && under_profiler()) no_coop = 1;     // see common.cuh
  }
  cfg.numAttrs = no_coop ? 1 : 2;
  e = cudaLaunchKernelEx(&cfg, kern, prm);
  if (e != cudaSuccess && cfg.numAttrs == 2) {
    // a profiler / sanitizer that refuses cooperative cluster launches: the plain cluster launch is equivalent here
    // (co-residency of every cluster was established by ark_gru_cluster_supported's occupancy query)
    (void)cudaGetLastError();
    cudaStreamCaptureStatus cs_ = cudaStreamCaptureStatusNone;
    (void)cudaStreamIsCapturing(s, &cs_);
    if (cs_ == cudaStreamCaptureStatusNone) {
      cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, kern, prm);
      if (e == cudaSuccess) no_coop = 1;
    }
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail((int)e, "%s: launch grid=(%u,%u,%u) cluster=%d smem=%d: %s", who, grid.x, grid.y, grid.z, cs, smem,
                cudaGetErrorString(e));
  }
  count_launch();
  return 0;
}

template <typename Params, typename K16, typename K32, typename K64>
static int dispatch_nb(int NB, K16 k16, K32 k32, K64 k64, const Params& prm, dim3 grid, int cs, int smem, cudaStream_t s,
                       const char* who, bool query_only) {
  if (NB == 16) return launch_cluster(k16, prm, grid, cs, smem, s, who, query_only);
  if (NB == 32) return launch_cluster(k32, prm, grid, cs, smem, s, who, query_only);
  return launch_cluster(k64, prm, grid, cs, smem, s, who, query_only);
}

// ARK_GRU_CLUSTER_DBG=<first iteration>: clock64 timeline of 4 iterations (read back with ark_gru_cluster_debug_dump)
static long long* g_dbg = nullptr;
static int dbg_it0() {
  static int v = -2;
  if (v == -2) { const char* e = getenv("ARK_GRU_CLUSTER_DBG"); v = e ? atoi(e) : -1; }
  return v;
}
static long long* dbg_buffer() {
  if (dbg_it0() < 0) return nullptr;
  if (!g_dbg) {
    if (cudaMalloc(&g_dbg, sizeof(long long) * GC_DBG_WORDS) != cudaSuccess) { (void)cudaGetLastError(); g_dbg = nullptr; return nullptr; }
    cudaMemset(g_dbg, 0, sizeof(long long) * GC_DBG_WORDS);
  }
  return g_dbg;
}

static int swap_lbo_flag() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ARK_GRU_CLUSTER_SWAP_LBO"); v = e ? atoi(e) : 0; }
  return v;
}

}  // namespace ark

using namespace ark;

extern "C" int ark_gru_cluster_debug_dump(int64_t* out_host, int64_t n_words) {
  if (!g_dbg) return fail(ARK_E_BADARG, "gru_cluster_debug_dump: ARK_GRU_CLUSTER_DBG was not set");
  if (n_words > GC_DBG_WORDS) n_words = GC_DBG_WORDS;
  cudaError_t e = cudaMemcpy(out_host, g_dbg, sizeof(long long) * n_words, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? 0 : fail((int)e, "gru_cluster_debug_dump: %s", cudaGetErrorString(e));
}

extern "C" int ark_gru_cluster_supported(int64_t d, int64_t bt0, int64_t nl, int64_t L) {
  ClusterPlan pl;
  if (!plan_cluster(d, bt0, nl, L, &pl)) return 0;
  // the kernels spin on flags of other clusters: every cluster of the launch must be co-resident
  // memoised on what the launch configuration depends on (d, layers, tile rows / count, shared-memory sizes): ragged
  // batches change L every step but hit the same handful of plans
  struct Memo { int64_t d, nl; ClusterPlan pl; int ok; };
  static thread_local std::vector<Memo> memo;
  for (const Memo& m : memo)
    if (m.d == d && m.nl == nl && m.pl.NB == pl.NB && m.pl.nbt == pl.nbt && m.pl.smem_f == pl.smem_f && m.pl.smem_b == pl.smem_b)
      return m.ok ? pl.NB : 0;
  GruClFwdParams pf;
  GruClBwdParams pb;
  memset(&pf, 0, sizeof(pf));
  memset(&pb, 0, sizeof(pb));
  const dim3 gf((unsigned)pl.CS, (unsigned)pl.nbt, (unsigned)(2 * nl)), gb((unsigned)pl.CS, (unsigned)pl.nbt, (unsigned)(2 * nl - 1));
  int ok = dispatch_nb(pl.NB, gru_cluster_fwd_kernel<16>, gru_cluster_fwd_kernel<32>, gru_cluster_fwd_kernel<64>, pf, gf, pl.CS,
                       pl.smem_f, 0, "gru_cluster_fwd", true) == 0 &&
           dispatch_nb(pl.NB, gru_cluster_bwd_kernel<16>, gru_cluster_bwd_kernel<32>, gru_cluster_bwd_kernel<64>, pb, gb, pl.CS,
                       pl.smem_b, 0, "gru_cluster_bwd", true) == 0;
  memo.push_back(Memo{d, nl, pl, ok});
  return ok ? pl.NB : 0;
}

extern "C" int64_t ark_gru_cluster_workspace_bytes(int64_t L, int64_t bt0, int64_t d, int64_t nl) {
  ClusterPlan pl;
  if (!plan_cluster(d, bt0, nl, L, &pl)) return 0;
  // forward gi^T scratch (the backward dx^T scratch is a third of it and reuses the same buffer)
  return nl * L * (int64_t)pl.nbt * pl.NB * 3 * d * 4;
}

extern "C" int ark_gru_cluster_fwd(const uint16_t* x_b, uint16_t* hp_b, uint16_t* out_b, const float* h0,
                                   const uint16_t* const* Wih_b, const uint16_t* const* Whh_b,
                                   const float* const* b_ih, const float* const* b_hh, const int32_t* bt_dev,
                                   const int32_t* off_dev, int64_t L, int64_t bt0, int64_t N, int64_t d, int64_t nl,
                                   uint16_t* r, uint16_t* z, uint16_t* n, uint16_t* ghn, uint8_t* mask, float p_drop,
                                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int32_t* sync_ws,
                                   void* ws, int64_t ws_bytes, void* stream) {
  ARK_REQUIRE(x_b && hp_b && out_b && Wih_b && Whh_b && b_ih && b_hh && bt_dev && off_dev && sync_ws && ws, ARK_E_BADARG,
              "gru_cluster_fwd: null pointer");
  ARK_REQUIRE((r && z && n && ghn) || (!r && !z && !n && !ghn), ARK_E_BADARG,
              "gru_cluster_fwd: gate outputs must be all set or all NULL");
  ARK_REQUIRE(L > 0 && N > 0 && bt0 > 0, ARK_E_BADARG, "gru_cluster_fwd: bad sizes");
  ARK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, ARK_E_BADARG, "gru_cluster_fwd: dropout probability must be in [0,1)");
  ClusterPlan pl;
  ARK_REQUIRE(plan_cluster(d, bt0, nl, L, &pl), ARK_E_SHAPE,
              "gru_cluster_fwd: unsupported shape d=%lld bt0=%lld nl=%lld (need d in {128,256,384,512}, nl <= 4 and "
              "2*nl*(d/32)*ceil(bt0/NB) <= 148 CTAs)", (long long)d, (long long)bt0, (long long)nl);
  ARK_REQUIRE(ws_bytes >= ark_gru_cluster_workspace_bytes(L, bt0, d, nl), ARK_E_BADARG, "gru_cluster_fwd: workspace too small");
  ARK_REQUIRE(aligned16(x_b) && aligned16(hp_b) && aligned16(out_b) && aligned16(ws) && (!h0 || aligned16(h0)), ARK_E_ALIGN,
              "gru_cluster_fwd: 16-byte alignment");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sync_ws, 0, sizeof(int32_t) * 2 * pl.nbt * nl * GC_MAXCS, s);
  if (e != cudaSuccess) return fail((int)e, "gru_cluster_fwd: memset: %s", cudaGetErrorString(e));
  GruClFwdParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  const int64_t LS = N * d;
  for (int k = 0; k < nl; ++k) {
    const uint16_t* u = k == 0 ? x_b : out_b + (int64_t)(k - 1) * LS;
    if ((rc = make_tmap_kchunked_bf16(&prm.tmU[k], u, (uint64_t)d, (uint64_t)N, (uint64_t)d, pl.NB, (uint32_t)(d / 64)))) return rc;
    ARK_REQUIRE(Wih_b[k] && Whh_b[k] && b_ih[k] && b_hh[k], ARK_E_BADARG, "gru_cluster_fwd: null weight pointer (layer %d)", k);
    ARK_REQUIRE(aligned16(Wih_b[k]) && aligned16(Whh_b[k]), ARK_E_ALIGN, "gru_cluster_fwd: weights must be 16-byte aligned");
    prm.wih[k] = Wih_b[k];
    prm.whh[k] = Whh_b[k];
    prm.b_ih[k] = b_ih[k];
    prm.b_hh[k] = b_hh[k];
  }
  prm.bt = bt_dev; prm.off = off_dev; prm.sync = sync_ws; prm.h0 = h0; prm.git = (float*)ws; prm.hp_b = hp_b; prm.out_b = out_b;
  prm.r = r; prm.z = z; prm.n = n; prm.ghn = ghn; prm.mask = mask; prm.offset_dev = offset_dev;
  prm.seed = seed; prm.offset = offset; prm.drop_stride = (uint64_t)((N * d + 3) / 4); prm.layer_stride = LS;
  prm.p_drop = p_drop; prm.L = (int)L; prm.d = (int)d; prm.nl = (int)nl; prm.nbt = pl.nbt; prm.S = pl.S_f;
  prm.swap_lbo = swap_lbo_flag();
  prm.dbg = dbg_buffer(); prm.dbg_it0 = dbg_it0();
  const dim3 grid((unsigned)pl.CS, (unsigned)pl.nbt, (unsigned)(2 * nl));
  return dispatch_nb(pl.NB, gru_cluster_fwd_kernel<16>, gru_cluster_fwd_kernel<32>, gru_cluster_fwd_kernel<64>, prm, grid, pl.CS,
                     pl.smem_f, s, "gru_cluster_fwd", false);
}

extern "C" int ark_gru_cluster_bwd(const float* dy_top, const uint16_t* r, const uint16_t* z, const uint16_t* n,
                                   const uint16_t* ghn, const uint16_t* hp_b, const uint8_t* mask, float p_drop,
                                   const uint16_t* const* WhhT_b, const uint16_t* const* WihT_b, const int32_t* bt_dev,
                                   const int32_t* off_dev, int64_t L, int64_t bt0, int64_t N, int64_t d, int64_t nl,
                                   uint16_t* dgi_b, uint16_t* dgh_b, float* dh0, int32_t* sync_ws, void* ws,
                                   int64_t ws_bytes, void* stream) {
  ARK_REQUIRE(dy_top && r && z && n && ghn && hp_b && WhhT_b && WihT_b && bt_dev && off_dev && dgi_b && dgh_b && sync_ws && ws,
              ARK_E_BADARG, "gru_cluster_bwd: null pointer");
  ARK_REQUIRE(L > 0 && N > 0 && bt0 > 0, ARK_E_BADARG, "gru_cluster_bwd: bad sizes");
  ARK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, ARK_E_BADARG, "gru_cluster_bwd: dropout probability must be in [0,1)");
  ARK_REQUIRE(p_drop == 0.f || nl == 1 || mask, ARK_E_BADARG, "gru_cluster_bwd: dropout needs the forward keep mask");
  ClusterPlan pl;
  ARK_REQUIRE(plan_cluster(d, bt0, nl, L, &pl), ARK_E_SHAPE, "gru_cluster_bwd: unsupported shape d=%lld bt0=%lld nl=%lld",
              (long long)d, (long long)bt0, (long long)nl);
  ARK_REQUIRE(ws_bytes >= ark_gru_cluster_workspace_bytes(L, bt0, d, nl), ARK_E_BADARG, "gru_cluster_bwd: workspace too small");
  ARK_REQUIRE(aligned16(dy_top) && aligned16(dgi_b) && aligned16(dgh_b) && aligned16(ws) && (!dh0 || aligned16(dh0)), ARK_E_ALIGN,
              "gru_cluster_bwd: 16-byte alignment");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sync_ws, 0, sizeof(int32_t) * 2 * pl.nbt * nl * GC_MAXCS, s);
  if (e != cudaSuccess) return fail((int)e, "gru_cluster_bwd: memset: %s", cudaGetErrorString(e));
  if (dh0) {
    e = cudaMemsetAsync(dh0, 0, sizeof(float) * bt0 * d, s);
    if (e != cudaSuccess) return fail((int)e, "gru_cluster_bwd: memset dh0: %s", cudaGetErrorString(e));
  }
  GruClBwdParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  const int64_t LS = N * d;
  for (int k = 0; k < nl; ++k) {
    if ((rc = make_tmap_kchunked_bf16(&prm.tmDgi[k], dgi_b + (int64_t)k * LS * 3, (uint64_t)(3 * d), (uint64_t)N,
                                      (uint64_t)(3 * d), pl.NB, (uint32_t)(3 * d / 64)))) return rc;
    ARK_REQUIRE(WhhT_b[k] && (k == 0 || WihT_b[k]), ARK_E_BADARG, "gru_cluster_bwd: null weight pointer (layer %d)", k);
    ARK_REQUIRE(aligned16(WhhT_b[k]), ARK_E_ALIGN, "gru_cluster_bwd: weights must be 16-byte aligned");
    prm.whhT[k] = WhhT_b[k];
    if (k > 0 && (rc = make_tmap_2d_bf16(&prm.tmWihT[k], WihT_b[k], (uint64_t)(3 * d), (uint64_t)d, (uint64_t)(3 * d), 64,
                                         GC_DJ))) return rc;
  }
  prm.bt = bt_dev; prm.off = off_dev; prm.sync = sync_ws; prm.dy_top = dy_top; prm.r = r; prm.z = z; prm.n = n;
  prm.ghn = ghn; prm.hp_b = hp_b; prm.mask = (p_drop > 0.f) ? mask : nullptr; prm.dgi_b = dgi_b; prm.dgh_b = dgh_b;
  prm.dxt = (float*)ws; prm.dh0 = dh0; prm.layer_stride = LS; prm.p_drop = p_drop; prm.L = (int)L; prm.d = (int)d;
  prm.nl = (int)nl; prm.nbt = pl.nbt; prm.S = pl.S_b; prm.swap_lbo = swap_lbo_flag();
  prm.dbg = dbg_buffer(); prm.dbg_it0 = dbg_it0();
  const dim3 grid((unsigned)pl.CS, (unsigned)pl.nbt, (unsigned)(2 * nl - 1));
  return dispatch_nb(pl.NB, gru_cluster_bwd_kernel<16>, gru_cluster_bwd_kernel<32>, gru_cluster_bwd_kernel<64>, prm, grid, pl.CS,
                     pl.smem_b, s, "gru_cluster_bwd", false);
}
