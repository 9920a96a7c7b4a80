// K6 (wavefront): the WHOLE multi-layer GRU stack (kgvae/model/models.py:121-127,141; decoder-only variant
// :329-343) through all time steps in ONE cooperative launch per direction.
//
// gru_persist.cu runs one layer per launch, so a stack of nl layers over L steps is a chain of nl*L dependent
// steps (wd-articles: 3*637 = 1911 steps of ~4-6 us each = 88 % of the training step).  Layer k at step t
// only needs layer k-1 at step t and layer k at step t-1, so the layers can run as a diagonal wavefront:
// the chain shrinks to L + nl - 1 steps.  To make that possible the input projection W_ih u_t (a big batch
// GEMM between the layers in the per-layer design) moves INTO the recurrence:
//
//   grid = (d/DJ hidden slices) x (ceil(B/128) batch tiles) x (nl layers), one CTA per SM, all co-resident.
//   forward  CTA (ji,bi,k): resident in smem: rows {g*d + j0..j0+DJ} (g=r,z,n) of W_ih^k AND of W_hh^k.
//            step t: TMEM[r|z|n_i] = u^k_t W_ih^T (as soon as layer k-1 finished step t), then
//                    TMEM[r|z] += h^k_{t-1} W_hh^T, TMEM[n_h] = h^k_{t-1} W_hn^T (as soon as step t-1 is done);
//            gate math + inter-layer dropout (Philox, same stream as the stand-alone kernel) in the epilogue.
//   backward CTA (ji,bi,k): resident: rows j0..j0+DJ of W_hh^k^T and of W_ih^{k+1}^T ([d,3d] each).
//            step t: TMEM[dx] = dgi^{k+1}_t W_ih^{k+1} (from the layer above), TMEM[rec] = dgh^k_{t+1} W_hh^k.
//
// The A operand of a step is a TMA box of R = min(128, round-up(B)) rows; the UMMA always spans 128 rows, the
// rows beyond R are whatever lies behind the slot in shared memory (their TMEM lanes are never read).
// Flags: one release/acquire counter per (layer, batch tile).  TMEM accumulators are double-buffered by
// step parity so the input-side MMA of step t+1 overlaps the epilogue of step t.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "gru_math.cuh"
#include "gru_dev.cuh"
#include "philox.cuh"
#include <string.h>

namespace ark {

constexpr int GW_MAXL = 4;
constexpr int GW_BK = 64;

struct GruWaveFwdParams {
  CUtensorMap tmU[GW_MAXL];    // layer input rows u^k [N,d] bf16 (u^0 = token embeddings), box {64, R, C}
  CUtensorMap tmH[GW_MAXL];    // h_prev rows hp^k [N,d] bf16, box {64, R, C}
  CUtensorMap tmWih[GW_MAXL];  // W_ih^k [3d,d] bf16, box {64, DJ}
  CUtensorMap tmWhh[GW_MAXL];  // W_hh^k [3d,d] bf16, box {64, DJ}
  const float* b_ih[GW_MAXL];
  const float* b_hh[GW_MAXL];
  const int32_t* bt;
  const int32_t* off;
  int32_t* sync;               // [nl * n batch tiles], zeroed before launch
  const float* h0;             // [bt[0], d] fp32 initial state of every layer, or null (zeros)
  uint16_t* hp_b;              // [nl, N, d] bf16 h_prev rows (block 0 of every layer pre-filled with bf16(h0))
  uint16_t* out_b;             // [nl, N, d] bf16 layer outputs (after dropout for k < nl-1) = u^{k+1}
  uint16_t *r, *z, *n, *ghn;   // [nl, N, d] bf16 saved gates (all null in eval mode)
  uint8_t* mask;               // [nl-1, N, d] dropout keep mask (null when p_drop == 0)
  const uint64_t* offset_dev;  // optional device-resident Philox offset (CUDA-graph replay)
  uint64_t seed, offset, drop_stride;
  int64_t layer_stride;        // N * d
  float p_drop;
  int L, d, nl, R, n_groups, C;   // ring of n_groups groups, one TMA op loads C k-chunks
  int tile0, nbt_total;           // this launch covers batch tiles [tile0, tile0 + gridDim.y) of nbt_total
};

struct GruWaveBwdParams {
  CUtensorMap tmDgi[GW_MAXL];   // dgi^k [N,3d] bf16, box {64, R, C}   (read by layer k-1)
  CUtensorMap tmDgh[GW_MAXL];   // dgh^k [N,3d] bf16, box {64, R, C}   (read by layer k)
  CUtensorMap tmWhhT[GW_MAXL];  // W_hh^k^T [d,3d] bf16, box {64, DJ}
  CUtensorMap tmWihT[GW_MAXL];  // W_ih^k^T [d,3d] bf16, box {64, DJ} (read by layer k-1)
  const int32_t* bt;
  const int32_t* off;
  int32_t* sync;
  const float* dy_top;          // [N, d] fp32 gradient w.r.t. the top layer's outputs
  const uint16_t *r, *z, *n, *ghn, *hp_b;   // [nl, N, d] saved by the forward kernel
  const uint8_t* mask;          // [nl-1, N, d] or null
  uint16_t *dgi_b, *dgh_b;      // [nl, N, 3d] bf16
  float* dh0;                   // [bt[0], d] fp32, zeroed before launch; every layer adds its share; may be null
  int64_t layer_stride;         // N * d
  float p_drop;
  int L, d, nl, R, n_groups, C;   // ring of n_groups groups, one TMA op loads C k-chunks
  int tile0, nbt_total;           // this launch covers batch tiles [tile0, tile0 + gridDim.y) of nbt_total
};

__device__ __forceinline__ void st_mask4(uint8_t* p, const bool* keep) {
  *reinterpret_cast<uint32_t*>(p) = (keep[0] ? 1u : 0u) | (keep[1] ? 0x100u : 0u) | (keep[2] ? 0x10000u : 0u) |
                                    (keep[3] ? 0x1000000u : 0u);
}

// =====================================================================================================
// forward
// =====================================================================================================
template <int DJ>
__global__ void __launch_bounds__(192, 1) gru_wave_fwd_kernel(const __grid_constant__ GruWaveFwdParams p) {
  constexpr int NC = 6 * DJ;  // accumulator columns: input side r|z|n, recurrent side r|z|n (summed in the epilogue)
  constexpr uint32_t TMEM_COLS = 2 * NC <= 256 ? 256 : 512;
  constexpr int ACC_LD = 4 * DJ + 1;   // staged: r | z | n_i | n_h
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L, R = p.R, n_slots = p.n_groups, C = p.C;
  const int nkc = d / GW_BK, gpp = nkc / C;   // groups per phase
  const int chunk_bytes = R * 128, slot_bytes = C * chunk_bytes;
  const int w_bytes = 3 * DJ * d * 2;
  uint8_t* wih_sm = smem;
  uint8_t* whh_sm = smem + w_bytes;
  uint8_t* a_sm = whh_sm + w_bytes;
  // the UMMA of the last slot reads 128 rows: keep (128-R)*128 B of this CTA's smem behind the ring
  float* acc_sm = reinterpret_cast<float*>(a_sm + n_slots * slot_bytes);   // [R][ACC_LD]
  float* bias_sm = acc_sm + R * ACC_LD;                                     // [4][DJ]: r, z (ih+hh), n_i, n_h
  uint64_t* full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bias_sm + 4 * DJ) + 7) & ~(uintptr_t)7);
  uint64_t* empty_bar = full_bar + n_slots;
  uint64_t* w_bar = empty_bar + n_slots;
  uint64_t* tmem_full_bar = w_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ji = blockIdx.x, bi = blockIdx.y + p.tile0, k = blockIdx.z, ns = gridDim.x, nbt = p.nbt_total;
  const int j0 = ji * DJ, m0 = bi * 128;
  const CUtensorMap* tmU = &p.tmU[k];
  const CUtensorMap* tmH = &p.tmH[k];
  int32_t* my_sync = p.sync + k * nbt + bi;
  const int32_t* below_sync = p.sync + (k - 1) * nbt + bi;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(tmU);
    ptx::prefetch_tmap(tmH);
    for (int s = 0; s < n_slots; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(w_bar, 1);
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + DJ) {
    const int j = j0 + (threadIdx.x - 64);
    const float* bi_ = p.b_ih[k];
    const float* bh_ = p.b_hh[k];
    bias_sm[j - j0] = bi_[j] + bh_[j];
    bias_sm[DJ + j - j0] = bi_[d + j] + bh_[d + j];
    bias_sm[2 * DJ + j - j0] = bi_[2 * d + j];
    bias_sm[3 * DJ + j - j0] = bh_[2 * d + j];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)(2 * w_bytes));
      for (int kc = 0; kc < nkc; ++kc)
        for (int g = 0; g < 3; ++g) {
          ptx::tma_load_2d(wih_sm + kc * (3 * DJ * 128) + g * (DJ * 128), &p.tmWih[k], w_bar, kc * GW_BK, g * d + j0);
          ptx::tma_load_2d(whh_sm + kc * (3 * DJ * 128) + g * (DJ * 128), &p.tmWhh[k], w_bar, kc * GW_BK, g * d + j0);
        }
      int it = 0;
      for (int t = 0; t < L; ++t) {
        if (m0 >= p.bt[t]) break;
        const int row0 = p.off[t] + m0;
        // input side: u^k_t (token embeddings for k = 0, the layer below's step-t output otherwise)
        if (k > 0) {
          wait_counter(below_sync, (t + 1) * ns);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int g = 0; g < gpp; ++g, ++it) {
          const int s = it % n_slots;
          ptx::mbar_wait(&empty_bar[s], ((it / n_slots) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)slot_bytes);
          ptx::tma_load_3d(a_sm + s * slot_bytes, tmU, &full_bar[s], 0, row0, g * C);
        }
        // recurrent side: h^k_{t-1}
        if (t > 0) {
          wait_counter(my_sync, t * ns);
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int g = 0; g < gpp; ++g, ++it) {
          const int s = it % n_slots;
          ptx::mbar_wait(&empty_bar[s], ((it / n_slots) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)slot_bytes);
          ptx::tma_load_3d(a_sm + s * slot_bytes, tmH, &full_bar[s], 0, row0, g * C);
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc3 = ptx::make_idesc_bf16(128, 3 * DJ, 0, 0);
      ptx::mbar_wait(w_bar, 0);
      const uint32_t wih_addr = ptx::smem_u32(wih_sm), whh_addr = ptx::smem_u32(whh_sm), a_addr0 = ptx::smem_u32(a_sm);
      int it = 0;
      for (int t = 0; t < L; ++t) {
        if (m0 >= p.bt[t]) break;
        const uint32_t acc = tmem_base + (uint32_t)((t & 1) * NC);
        for (int g = 0; g < gpp; ++g, ++it) {
          const int s = it % n_slots;
          ptx::mbar_wait(&full_bar[s], (it / n_slots) & 1);
          ptx::tc_fence_after();
          for (int c = 0; c < C; ++c) {
            const int kc = g * C + c;
#pragma unroll
            for (int kk = 0; kk < GW_BK / 16; ++kk) {
              const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr0 + s * slot_bytes + c * chunk_bytes + kk * 32, 16, 1024);
              const uint64_t bdesc = ptx::make_smem_desc_sw128(wih_addr + kc * (3 * DJ * 128) + kk * 32, 16, 1024);
              ptx::umma_f16(acc, adesc, bdesc, idesc3, (kc | kk) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty_bar[s]);
        }
        for (int g = 0; g < gpp; ++g, ++it) {
          const int s = it % n_slots;
          ptx::mbar_wait(&full_bar[s], (it / n_slots) & 1);
          ptx::tc_fence_after();
          for (int c = 0; c < C; ++c) {
            const int kc = g * C + c;
#pragma unroll
            for (int kk = 0; kk < GW_BK / 16; ++kk) {
              const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr0 + s * slot_bytes + c * chunk_bytes + kk * 32, 16, 1024);
              const uint64_t bdesc = ptx::make_smem_desc_sw128(whh_addr + kc * (3 * DJ * 128) + kk * 32, 16, 1024);
              ptx::umma_f16(acc + 3 * DJ, adesc, bdesc, idesc3, (kc | kk) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty_bar[s]);
        }
        ptx::umma_commit(tmem_full_bar);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> smem, then (row, 4-unit group) work items over 128 threads =====
    constexpr int G = DJ / 4;
    const int q = warp & 3;
    const int tid = threadIdx.x - 64;
    const int64_t LS = p.layer_stride;
    uint16_t* hp_k = p.hp_b + (int64_t)k * LS;
    uint16_t* out_k = p.out_b + (int64_t)k * LS;
    uint16_t* r_k = p.r ? p.r + (int64_t)k * LS : nullptr;
    uint16_t* z_k = p.r ? p.z + (int64_t)k * LS : nullptr;
    uint16_t* n_k = p.r ? p.n + (int64_t)k * LS : nullptr;
    uint16_t* g_k = p.r ? p.ghn + (int64_t)k * LS : nullptr;
    const bool drop = p.p_drop > 0.f && k < p.nl - 1;
    uint8_t* mask_k = (drop && p.mask) ? p.mask + (int64_t)k * LS : nullptr;
    const float scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
    const uint64_t ctr0 = p.offset + (p.offset_dev ? *p.offset_dev : 0ull) + (uint64_t)k * p.drop_stride;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    float hreg[G][4];
    const int bt0 = p.bt[0];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int e = tid + 128 * i, bl = e / G, j = j0 + (e % G) * 4;
      const int b = m0 + bl;
      const float4 hv = (p.h0 && b < bt0) ? *reinterpret_cast<const float4*>(p.h0 + (int64_t)b * d + j)
                                          : make_float4(0, 0, 0, 0);
      hreg[i][0] = hv.x; hreg[i][1] = hv.y; hreg[i][2] = hv.z; hreg[i][3] = hv.w;
    }
    for (int t = 0; t < L; ++t) {
      const int Bt = p.bt[t];
      if (m0 >= Bt) break;
      const int Bn = (t + 1 < L) ? p.bt[t + 1] : 0;
      const int64_t base = (int64_t)p.off[t] + m0;
      const int64_t base_n = (t + 1 < L) ? (int64_t)p.off[t + 1] + m0 : 0;
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((t & 1) * NC);
      ptx::mbar_wait(tmem_full_bar, t & 1);
      ptx::tc_fence_after();
      if (m0 + q * 32 < Bt) {
        const int row = q * 32 + lane;
        float* dst = acc_sm + row * ACC_LD;
        // one gate per pass (input-side and recurrent-side columns in flight together, one wait per pass)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint32_t vi[DJ], vh[DJ];
#pragma unroll
          for (int c = 0; c < DJ; c += 16) {
            ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(g * DJ + c), vi + c);
            ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)(3 * DJ + g * DJ + c), vh + c);
          }
          ptx::tmem_ld_wait();
          if (row < R) {
            if (g < 2) {   // r, z: input + recurrent parts summed here
#pragma unroll
              for (int c = 0; c < DJ; ++c) dst[g * DJ + c] = __uint_as_float(vi[c]) + __uint_as_float(vh[c]);
            } else {       // n keeps them apart: n = tanh(i_n + r * h_n)
#pragma unroll
              for (int c = 0; c < DJ; ++c) {
                dst[2 * DJ + c] = __uint_as_float(vi[c]);
                dst[3 * DJ + c] = __uint_as_float(vh[c]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      epi_bar_sync();
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int e = tid + 128 * i, bl = e / G, jl = (e % G) * 4;
        const int b = m0 + bl;
        if (b < Bt) {
          const float* ap = acc_sm + bl * ACC_LD + jl;
          float o_r[4], o_z[4], o_n[4], o_g[4], o_h[4];
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            const GruFwd o = gru_fwd_math_fast(ap[kx] + bias_sm[jl + kx], ap[DJ + kx] + bias_sm[DJ + jl + kx],
                                               ap[2 * DJ + kx] + bias_sm[2 * DJ + jl + kx], 0.f, 0.f,
                                               ap[3 * DJ + kx] + bias_sm[3 * DJ + jl + kx], hreg[i][kx]);
            hreg[i][kx] = o.h;
            o_r[kx] = o.r; o_z[kx] = o.z; o_n[kx] = o.n; o_g[kx] = o.ghn; o_h[kx] = o.h;
          }
          const int64_t o = (base + bl) * d + j0 + jl;
          if (b < Bn) st4_bf16(hp_k + (base_n + bl) * d + j0 + jl, o_h);
          if (r_k) {
            st4_bf16(r_k + o, o_r);
            st4_bf16(z_k + o, o_z);
            st4_bf16(n_k + o, o_n);
            st4_bf16(g_k + o, o_g);
          }
          if (drop) {   // same draw as dropout_bf16_kernel over the [N, d] output of layer k
            const uint64_t c = ctr0 + (uint64_t)(o >> 2);
            const uint4 rn = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
            const uint32_t rr[4] = {rn.x, rn.y, rn.z, rn.w};
            bool keep[4];
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
              keep[kx] = (float)(rr[kx] >> 8) * (1.f / 16777216.f) >= p.p_drop;
              o_h[kx] = keep[kx] ? bf16_bits_to_f32(f32_to_bf16_bits(o_h[kx])) * scale : 0.f;
            }
            if (mask_k) st_mask4(mask_k + o, keep);
          }
          st4_bf16(out_k + o, o_h);
        }
      }
      epi_bar_sync();                                  // all stores of the tile issued, acc_sm free again
      if (tid == 0) red_release_add(my_sync, 1);       // release: cumulative over the barrier above
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// =====================================================================================================
// backward through time
// =====================================================================================================
template <int DJ>
__global__ void __launch_bounds__(192, 1) gru_wave_bwd_kernel(const __grid_constant__ GruWaveBwdParams p) {
  constexpr int NC = 2 * DJ;  // accumulator columns: rec | dx
  constexpr uint32_t TMEM_COLS = 2 * NC <= 64 ? 64 : 128;
  constexpr int ACC_LD = NC + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int d = p.d, L = p.L, R = p.R, n_slots = p.n_groups, C = p.C, nl = p.nl;
  const int nkc = 3 * d / GW_BK, gpp = nkc / C;
  const int chunk_bytes = R * 128, slot_bytes = C * chunk_bytes;
  const int w_bytes = 3 * DJ * d * 2;
  uint8_t* whh_sm = smem;               // rows j0.. of W_hh^k^T     [DJ x 3d]
  uint8_t* wih_sm = smem + w_bytes;     // rows j0.. of W_ih^{k+1}^T [DJ x 3d]
  uint8_t* a_sm = wih_sm + w_bytes;
  float* acc_sm = reinterpret_cast<float*>(a_sm + n_slots * slot_bytes);   // [R][ACC_LD]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(acc_sm + R * ACC_LD) + 7) & ~(uintptr_t)7);
  uint64_t* empty_bar = full_bar + n_slots;
  uint64_t* w_bar = empty_bar + n_slots;
  uint64_t* tmem_full_bar = w_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ji = blockIdx.x, bi = blockIdx.y + p.tile0, k = blockIdx.z, ns = gridDim.x, nbt = p.nbt_total;
  const int j0 = ji * DJ, m0 = bi * 128;
  const bool top = (k == nl - 1);
  const CUtensorMap* tmA1 = &p.tmDgi[top ? k : k + 1];   // dgi of the layer above
  const CUtensorMap* tmA2 = &p.tmDgh[k];
  int32_t* my_sync = p.sync + k * nbt + bi;
  const int32_t* above_sync = p.sync + (top ? k : k + 1) * nbt + bi;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(tmA1);
    ptx::prefetch_tmap(tmA2);
    for (int s = 0; s < n_slots; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
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
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Iterations t = t_first .. 0 (cell backward of step t) and t = -1 (gradient of the initial state); the
  // set of iterations of a batch tile is the same in every layer (same bt), so "iterations done" is the clock
  // the layers synchronise on.
  auto tile_active = [&](int t) { return m0 < p.bt[t < 0 ? 0 : t]; };
  auto has_rec = [&](int t) { return (t + 1 <= L - 1) && (m0 < p.bt[t + 1]); };
  auto has_dx = [&](int t) { return !top && t >= 0; };

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(w_bar, (uint32_t)((top ? 1 : 2) * w_bytes));
      for (int kc = 0; kc < nkc; ++kc) {
        ptx::tma_load_2d(whh_sm + kc * (DJ * 128), &p.tmWhhT[k], w_bar, kc * GW_BK, j0);
        if (!top) ptx::tma_load_2d(wih_sm + kc * (DJ * 128), &p.tmWihT[k + 1], w_bar, kc * GW_BK, j0);
      }
      int it = 0, done = 0;
      for (int t = L - 1; t >= -1; --t) {
        if (!tile_active(t)) continue;
        if (has_dx(t)) {
          wait_counter(above_sync, (done + 1) * ns);     // every slice of dgi^{k+1}_t is in global memory
          asm volatile("fence.proxy.async;" ::: "memory");
          const int row0 = p.off[t] + m0;
          for (int g = 0; g < gpp; ++g, ++it) {
            const int s = it % n_slots;
            ptx::mbar_wait(&empty_bar[s], ((it / n_slots) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)slot_bytes);
            ptx::tma_load_3d(a_sm + s * slot_bytes, tmA1, &full_bar[s], 0, row0, g * C);
          }
        }
        if (has_rec(t)) {
          wait_counter(my_sync, done * ns);              // every slice of dgh^k_{t+1} is in global memory
          asm volatile("fence.proxy.async;" ::: "memory");
          const int row0 = p.off[t + 1] + m0;
          for (int g = 0; g < gpp; ++g, ++it) {
            const int s = it % n_slots;
            ptx::mbar_wait(&empty_bar[s], ((it / n_slots) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)slot_bytes);
            ptx::tma_load_3d(a_sm + s * slot_bytes, tmA2, &full_bar[s], 0, row0, g * C);
          }
        }
        ++done;
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, DJ, 0, 0);
      ptx::mbar_wait(w_bar, 0);
      const uint32_t whh_addr = ptx::smem_u32(whh_sm), wih_addr = ptx::smem_u32(wih_sm), a_addr0 = ptx::smem_u32(a_sm);
      int it = 0, n_mma = 0;
      for (int t = L - 1; t >= -1; --t) {
        if (!tile_active(t)) continue;
        const bool dx = has_dx(t), rec = has_rec(t);
        if (!dx && !rec) continue;
        const uint32_t acc = tmem_base + (uint32_t)((n_mma & 1) * NC);
        if (dx) {
          for (int g = 0; g < gpp; ++g, ++it) {
            const int s = it % n_slots;
            ptx::mbar_wait(&full_bar[s], (it / n_slots) & 1);
            ptx::tc_fence_after();
            for (int c = 0; c < C; ++c) {
              const int kc = g * C + c;
#pragma unroll
              for (int kk = 0; kk < GW_BK / 16; ++kk) {
                const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr0 + s * slot_bytes + c * chunk_bytes + kk * 32, 16, 1024);
                const uint64_t bdesc = ptx::make_smem_desc_sw128(wih_addr + kc * (DJ * 128) + kk * 32, 16, 1024);
                ptx::umma_f16(acc + DJ, adesc, bdesc, idesc, (kc | kk) != 0 ? 1u : 0u);
              }
            }
            ptx::umma_commit(&empty_bar[s]);
          }
        }
        if (rec) {
          for (int g = 0; g < gpp; ++g, ++it) {
            const int s = it % n_slots;
            ptx::mbar_wait(&full_bar[s], (it / n_slots) & 1);
            ptx::tc_fence_after();
            for (int c = 0; c < C; ++c) {
              const int kc = g * C + c;
#pragma unroll
              for (int kk = 0; kk < GW_BK / 16; ++kk) {
                const uint64_t adesc = ptx::make_smem_desc_sw128(a_addr0 + s * slot_bytes + c * chunk_bytes + kk * 32, 16, 1024);
                const uint64_t bdesc = ptx::make_smem_desc_sw128(whh_addr + kc * (DJ * 128) + kk * 32, 16, 1024);
                ptx::umma_f16(acc, adesc, bdesc, idesc, (kc | kk) != 0 ? 1u : 0u);
              }
            }
            ptx::umma_commit(&empty_bar[s]);
          }
        }
        ptx::umma_commit(tmem_full_bar);
        ++n_mma;
      }
    }
  } else {
    constexpr int G = DJ / 4;
    const int q = warp & 3;
    const int tid = threadIdx.x - 64;
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
    float carry[G][4];   // dh_{t+1} * z_{t+1}: the direct path into h_t (valid for rows of step t+1)
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) carry[i][kx] = 0.f;
    float4 dyp[G];
    uint32_t mk[G];
    uint2 sp[G][5];
    auto prefetch = [&](int t) {
      const int Bt = p.bt[t];
      const int64_t base = (int64_t)p.off[t] + m0;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int e = tid + 128 * i, bl = e / G, jl = (e % G) * 4;
        if (m0 + bl < Bt) {
          const int64_t o = (base + bl) * d + j0 + jl;
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
    int t_first = L - 1;
    while (t_first >= 0 && !tile_active(t_first)) --t_first;
    if (t_first >= 0) prefetch(t_first);
    int n_mma = 0;
    for (int t = t_first; t >= -1; --t) {
      const bool dx = has_dx(t), rec = has_rec(t);
      const int B_next = (t + 1 <= L - 1) ? p.bt[t + 1] : 0;
      const int Bt = p.bt[t < 0 ? 0 : t];
      if (dx || rec) {
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((n_mma & 1) * NC);
        ptx::mbar_wait(tmem_full_bar, n_mma & 1);
        ptx::tc_fence_after();
        ++n_mma;
        if (m0 + q * 32 < Bt) {
          const int row = q * 32 + lane;
          float* dst = acc_sm + row * ACC_LD;
          uint32_t v[NC];
#pragma unroll
          for (int c = 0; c < NC; c += 16) ptx::tmem_ld_32x32b_x16(t_lane + (uint32_t)c, v + c);
          ptx::tmem_ld_wait();
          if (row < R) {
#pragma unroll
            for (int c = 0; c < NC; ++c) dst[c] = __uint_as_float(v[c]);
          }
        }
        ptx::tc_fence_before();
      }
      epi_bar_sync();
      const int64_t base = (t >= 0) ? (int64_t)p.off[t] + m0 : 0;
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const int e = tid + 128 * i, bl = e / G, jl = (e % G) * 4;
        const int b = m0 + bl;
        if (b >= Bt) continue;
        const bool from_next = b < B_next;
        float dh[4];
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
          dh[kx] = from_next ? carry[i][kx] + (rec ? acc_sm[bl * ACC_LD + jl + kx] : 0.f) : 0.f;
        if (t < 0) {
          if (p.dh0) red_add_v4(p.dh0 + (int64_t)b * d + j0 + jl, make_float4(dh[0], dh[1], dh[2], dh[3]));
          continue;
        }
        float dyv[4];
        if (top) {
          dyv[0] = dyp[i].x; dyv[1] = dyp[i].y; dyv[2] = dyp[i].z; dyv[3] = dyp[i].w;
        } else {
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) {
            const float m = drop ? (((mk[i] >> (8 * kx)) & 0xffu) ? scale : 0.f) : 1.f;
            dyv[kx] = m * acc_sm[bl * ACC_LD + DJ + jl + kx];
          }
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
        float dar[4], daz[4], dan[4], danr[4];
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const GruBwd w = gru_bwd_math(dh[kx] + dyv[kx], r[kx], z[kx], n[kx], g[kx], hp[kx]);
          dar[kx] = w.dar; daz[kx] = w.daz; dan[kx] = w.dan; danr[kx] = w.dan_r;
          carry[i][kx] = w.dh_prev;
        }
        const int64_t o3 = (base + bl) * d3 + j0 + jl;
        st4_bf16(dgi_k + o3, dar);
        st4_bf16(dgi_k + o3 + d, daz);
        st4_bf16(dgi_k + o3 + 2 * d, dan);
        st4_bf16(dgh_k + o3, dar);
        st4_bf16(dgh_k + o3 + d, daz);
        st4_bf16(dgh_k + o3 + 2 * d, danr);
      }
      epi_bar_sync();
      if (tid == 0) red_release_add(my_sync, 1);
      if (t - 1 >= 0) prefetch(t - 1);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------- host
struct WavePlan {
  int dj, R, n_groups, C, smem;
};

static int round_rows(int64_t bt0) {
  const int64_t b = bt0 < 128 ? bt0 : 128;
  int r = 16;
  while (r < b) r *= 2;
  return r;
}

// One phase of a step streams nkc k-chunks of [R x 64] through the ring (forward nkc = d/64, backward 3d/64; two
// phases per step).  A ring group = C chunks loaded by ONE TMA op; pick the largest C (a divisor of nkc) that still
// leaves >= 2 groups in shared memory.
static bool plan_wave(int64_t d, int64_t bt0, int64_t nl, bool bwd, WavePlan* out) {
  if (d % 64 != 0 || d < 64 || bt0 <= 0 || nl < 1 || nl > GW_MAXL) return false;
  const int R = round_rows(bt0);
  const int cand[2] = {16, 32};
  for (int i = 0; i < 2; ++i) {
    const int dj = cand[i];
    if (d % dj) continue;
    if ((d / dj) * nl > kNumSMs) continue;          // one batch tile of every layer must be co-resident
    const int64_t w = 2LL * 3 * dj * d * 2;
    const int nc = bwd ? 2 * dj : 4 * dj;
    const int64_t fixed = w + (int64_t)R * (nc + 1) * 4 + 4 * dj * 4 + 2048 + 1024;   // weights, acc, bias, barriers, align
    const int64_t tail = (int64_t)(128 - R) * 128;     // the last chunk's UMMA reads 128 rows
    const int nkc = (int)((bwd ? 3 : 1) * (d / 64));
    for (int C = nkc; C >= 1; --C) {
      if (nkc % C || C > 256) continue;
      const int64_t group = (int64_t)C * R * 128;
      int64_t n = (227 * 1024 - fixed - tail - 1024) / group;
      const int64_t want = 3 * (nkc / C);              // 1.5 steps in flight is plenty
      if (n > want) n = want;
      if (n > 32) n = 32;
      if (n < 2) continue;
      int64_t total = fixed + n * group;
      const int64_t need_end = w + n * group + tail + 1024;   // allocation must cover the overshoot
      if (total < need_end) total = need_end;
      total += n * 16;
      if (total > 227 * 1024) continue;
      out->dj = dj; out->R = R; out->n_groups = (int)n; out->C = C; out->smem = (int)total;
      return true;
    }
  }
  return false;
}

template <typename Params, typename Kern>
static int launch_wave(Kern kern, const Params& prm, dim3 grid, int smem, cudaStream_t s, const char* who) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return fail((int)e, "%s: smem attribute (%d B): %s", who, smem, cudaGetErrorString(e));
  void* args[] = {(void*)&prm};
  e = cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(192), args, (size_t)smem, s);
  if (e != cudaSuccess)
    return fail((int)e, "%s: cooperative launch grid=(%u,%u,%u) smem=%d: %s", who, grid.x, grid.y, grid.z, smem,
                cudaGetErrorString(e));
  count_launch();
  return 0;
}

}  // namespace ark

using namespace ark;

extern "C" int ark_gru_wave_supported(int64_t d, int64_t bt0, int64_t nl) {
  WavePlan f, b;
  if (!plan_wave(d, bt0, nl, false, &f) || !plan_wave(d, bt0, nl, true, &b)) return 0;
  return f.dj;
}

extern "C" int ark_gru_wave_fwd(const uint16_t* x_b, uint16_t* hp_b, uint16_t* out_b, const float* h0,
                                const uint16_t* const* Wih_b, const uint16_t* const* Whh_b,
                                const float* const* b_ih, const float* const* b_hh, const int32_t* bt_dev,
                                const int32_t* off_dev, int64_t L, int64_t bt0, int64_t N, int64_t d, int64_t nl,
                                uint16_t* r, uint16_t* z, uint16_t* n, uint16_t* ghn, uint8_t* mask, float p_drop,
                                uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int32_t* sync_ws,
                                void* stream) {
  ARK_REQUIRE(x_b && hp_b && out_b && Wih_b && Whh_b && b_ih && b_hh && bt_dev && off_dev && sync_ws, ARK_E_BADARG,
              "gru_wave_fwd: null pointer");
  ARK_REQUIRE((r && z && n && ghn) || (!r && !z && !n && !ghn), ARK_E_BADARG,
              "gru_wave_fwd: gate outputs must be all set or all NULL");
  ARK_REQUIRE(L > 0 && N > 0 && bt0 > 0, ARK_E_BADARG, "gru_wave_fwd: bad sizes");
  ARK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, ARK_E_BADARG, "gru_wave_fwd: dropout probability must be in [0,1)");
  WavePlan pl;
  ARK_REQUIRE(plan_wave(d, bt0, nl, false, &pl), ARK_E_SHAPE,
              "gru_wave_fwd: unsupported shape d=%lld bt0=%lld nl=%lld (need d %% 64 == 0, nl <= 4, the resident "
              "weights to fit shared memory and the grid to fit 148 SMs)", (long long)d, (long long)bt0, (long long)nl);
  ARK_REQUIRE(aligned16(x_b) && aligned16(hp_b) && aligned16(out_b) && (!h0 || aligned16(h0)), ARK_E_ALIGN,
              "gru_wave_fwd: 16-byte alignment");
  cudaStream_t s = (cudaStream_t)stream;
  const int nbt = (int)((bt0 + 127) / 128);
  cudaError_t e = cudaMemsetAsync(sync_ws, 0, sizeof(int32_t) * nbt * nl, s);
  if (e != cudaSuccess) return fail((int)e, "gru_wave_fwd: memset: %s", cudaGetErrorString(e));
  GruWaveFwdParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  const int64_t LS = N * d;
  for (int k = 0; k < nl; ++k) {
    const uint16_t* u = k == 0 ? x_b : out_b + (int64_t)(k - 1) * LS;
    if ((rc = make_tmap_kchunked_bf16(&prm.tmU[k], u, (uint64_t)d, (uint64_t)N, (uint64_t)d, pl.R, pl.C))) return rc;
    if ((rc = make_tmap_kchunked_bf16(&prm.tmH[k], hp_b + (int64_t)k * LS, (uint64_t)d, (uint64_t)N, (uint64_t)d, pl.R, pl.C)))
      return rc;
    ARK_REQUIRE(Wih_b[k] && Whh_b[k] && b_ih[k] && b_hh[k], ARK_E_BADARG, "gru_wave_fwd: null weight pointer (layer %d)", k);
    if ((rc = make_tmap_2d_bf16(&prm.tmWih[k], Wih_b[k], (uint64_t)d, (uint64_t)(3 * d), (uint64_t)d, GW_BK, pl.dj))) return rc;
    if ((rc = make_tmap_2d_bf16(&prm.tmWhh[k], Whh_b[k], (uint64_t)d, (uint64_t)(3 * d), (uint64_t)d, GW_BK, pl.dj))) return rc;
    prm.b_ih[k] = b_ih[k];
    prm.b_hh[k] = b_hh[k];
  }
  prm.bt = bt_dev; prm.off = off_dev; prm.sync = sync_ws; prm.h0 = h0; prm.hp_b = hp_b; prm.out_b = out_b;
  prm.r = r; prm.z = z; prm.n = n; prm.ghn = ghn; prm.mask = mask; prm.offset_dev = offset_dev;
  prm.seed = seed; prm.offset = offset; prm.drop_stride = (uint64_t)((N * d + 3) / 4); prm.layer_stride = LS;
  prm.p_drop = p_drop; prm.L = (int)L; prm.d = (int)d; prm.nl = (int)nl; prm.R = pl.R; prm.n_groups = pl.n_groups; prm.C = pl.C;
  // batch tiles are independent: when all of them do not fit the 148 SMs at once, run groups of tiles back to back
  const int tpl = (int)(kNumSMs / ((d / pl.dj) * nl));
  prm.nbt_total = nbt;
  for (int t0 = 0; t0 < nbt; t0 += tpl) {
    prm.tile0 = t0;
    dim3 grid((unsigned)(d / pl.dj), (unsigned)(nbt - t0 < tpl ? nbt - t0 : tpl), (unsigned)nl);
    rc = pl.dj == 16 ? launch_wave(gru_wave_fwd_kernel<16>, prm, grid, pl.smem, s, "gru_wave_fwd")
                     : launch_wave(gru_wave_fwd_kernel<32>, prm, grid, pl.smem, s, "gru_wave_fwd");
    if (rc) return rc;
  }
  return 0;
}

extern "C" int ark_gru_wave_bwd(const float* dy_top, const uint16_t* r, const uint16_t* z, const uint16_t* n,
                                const uint16_t* ghn, const uint16_t* hp_b, const uint8_t* mask, float p_drop,
                                const uint16_t* const* WhhT_b, const uint16_t* const* WihT_b, const int32_t* bt_dev,
                                const int32_t* off_dev, int64_t L, int64_t bt0, int64_t N, int64_t d, int64_t nl,
                                uint16_t* dgi_b, uint16_t* dgh_b, float* dh0, int32_t* sync_ws, void* stream) {
  ARK_REQUIRE(dy_top && r && z && n && ghn && hp_b && WhhT_b && WihT_b && bt_dev && off_dev && dgi_b && dgh_b && sync_ws,
              ARK_E_BADARG, "gru_wave_bwd: null pointer");
  ARK_REQUIRE(L > 0 && N > 0 && bt0 > 0, ARK_E_BADARG, "gru_wave_bwd: bad sizes");
  ARK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, ARK_E_BADARG, "gru_wave_bwd: dropout probability must be in [0,1)");
  ARK_REQUIRE(p_drop == 0.f || nl == 1 || mask, ARK_E_BADARG, "gru_wave_bwd: dropout needs the forward keep mask");
  WavePlan pl;
  ARK_REQUIRE(plan_wave(d, bt0, nl, true, &pl), ARK_E_SHAPE, "gru_wave_bwd: unsupported shape d=%lld bt0=%lld nl=%lld",
              (long long)d, (long long)bt0, (long long)nl);
  ARK_REQUIRE(aligned16(dy_top) && aligned16(dgi_b) && aligned16(dgh_b) && (!dh0 || aligned16(dh0)), ARK_E_ALIGN,
              "gru_wave_bwd: 16-byte alignment");
  cudaStream_t s = (cudaStream_t)stream;
  const int nbt = (int)((bt0 + 127) / 128);
  cudaError_t e = cudaMemsetAsync(sync_ws, 0, sizeof(int32_t) * nbt * nl, s);
  if (e != cudaSuccess) return fail((int)e, "gru_wave_bwd: memset: %s", cudaGetErrorString(e));
  if (dh0) {
    e = cudaMemsetAsync(dh0, 0, sizeof(float) * bt0 * d, s);
    if (e != cudaSuccess) return fail((int)e, "gru_wave_bwd: memset dh0: %s", cudaGetErrorString(e));
  }
  GruWaveBwdParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  const int64_t LS = N * d;
  for (int k = 0; k < nl; ++k) {
    if ((rc = make_tmap_kchunked_bf16(&prm.tmDgi[k], dgi_b + (int64_t)k * LS * 3, (uint64_t)(3 * d), (uint64_t)N,
                                      (uint64_t)(3 * d), pl.R, pl.C))) return rc;
    if ((rc = make_tmap_kchunked_bf16(&prm.tmDgh[k], dgh_b + (int64_t)k * LS * 3, (uint64_t)(3 * d), (uint64_t)N,
                                      (uint64_t)(3 * d), pl.R, pl.C))) return rc;
    ARK_REQUIRE(WhhT_b[k] && (k == 0 || WihT_b[k]), ARK_E_BADARG, "gru_wave_bwd: null weight pointer (layer %d)", k);
    if ((rc = make_tmap_2d_bf16(&prm.tmWhhT[k], WhhT_b[k], (uint64_t)(3 * d), (uint64_t)d, (uint64_t)(3 * d), GW_BK, pl.dj)))
      return rc;
    if (k > 0 && (rc = make_tmap_2d_bf16(&prm.tmWihT[k], WihT_b[k], (uint64_t)(3 * d), (uint64_t)d, (uint64_t)(3 * d), GW_BK,
                                         pl.dj))) return rc;
  }
  prm.bt = bt_dev; prm.off = off_dev; prm.sync = sync_ws; prm.dy_top = dy_top; prm.r = r; prm.z = z; prm.n = n;
  prm.ghn = ghn; prm.hp_b = hp_b; prm.mask = (p_drop > 0.f) ? mask : nullptr; prm.dgi_b = dgi_b; prm.dgh_b = dgh_b;
  prm.dh0 = dh0; prm.layer_stride = LS; prm.p_drop = p_drop; prm.L = (int)L; prm.d = (int)d; prm.nl = (int)nl;
  prm.R = pl.R; prm.n_groups = pl.n_groups; prm.C = pl.C;
  const int tpl = (int)(kNumSMs / ((d / pl.dj) * nl));
  prm.nbt_total = nbt;
  for (int t0 = 0; t0 < nbt; t0 += tpl) {
    prm.tile0 = t0;
    dim3 grid((unsigned)(d / pl.dj), (unsigned)(nbt - t0 < tpl ? nbt - t0 : tpl), (unsigned)nl);
    rc = pl.dj == 16 ? launch_wave(gru_wave_bwd_kernel<16>, prm, grid, pl.smem, s, "gru_wave_bwd")
                     : launch_wave(gru_wave_bwd_kernel<32>, prm, grid, pl.smem, s, "gru_wave_bwd");
    if (rc) return rc;
  }
  return 0;
}
