// K3: bf16 tensor-core GEMM for sm_100a —  C[M,N] = epi(A[M,K] . B[N,K]^T + bias[N])
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (UMMA 128 x BN x 16,
//   fp32 accumulator in TMEM) -> tcgen05.ld -> fused epilogue (bias, GELU/tanh, aux pre-activation,
//   f32|bf16 store, optional += for weight-gradient accumulation).
// Replaces the reference's nn.Linear / F.linear calls and their autograd GEMMs
// (kgvae/model/models.py:36,43-44,61-62,120,128,139,142): every dense contraction of the ELBO step.
// Operand majors: K-major (contraction contiguous) and MN-major (stored [K, M|N]) are both consumed
// directly, so dX = dY.W and dW = dY^T.X need no transposed copies in HBM.
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM owner + MMA issuer
// (one elected lane), warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).  One output tile per CTA;
// 96 KB of smem per CTA lets two CTAs share an SM so one tile's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "gemm_tc.cuh"

namespace ark {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;  // 64 bf16 = 128 B = one swizzle row

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

__device__ __forceinline__ float apply_act(float v, int epilogue) {
  if (epilogue == ARK_EPI_GELU) return gelu_erf(v);
  if (epilogue == ARK_EPI_TANH) return tanhf(v);
  return v;
}

template <int BN, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const EpiParams ep,
                                                      const int M, const int N, const int K,
                                                      const int a_row0, const int b_row0) {
  using L = TcSmem<BN, STAGES>;
  static_assert(TC_BM * (BN + 4) * 4 <= L::BAR_OFFSET, "epilogue staging must fit in the TMA ring");
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // swap_raster: the fast grid index walks the M tiles (A is the smaller operand), so each B tile is
  // fetched once and A stays L2-resident; otherwise the fast index walks the N tiles.
  const int n0 = (ep.swap_raster ? blockIdx.y : blockIdx.x) * BN, m0 = (ep.swap_raster ? blockIdx.x : blockIdx.y) * TC_BM;
  const int num_kb = (K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_s = smem + s * L::STAGE_BYTES;
        uint8_t* b_s = a_s + L::A_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < TC_BM / 64; ++j)
            ptx::tma_load_2d(a_s + j * (TC_BK * 128), &tmA, &full_bar[s], a_row0 + m0 + 64 * j, kb * TC_BK);
        } else {
          ptx::tma_load_2d(a_s, &tmA, &full_bar[s], kb * TC_BK, a_row0 + m0);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            ptx::tma_load_2d(b_s + j * (TC_BK * 128), &tmB, &full_bar[s], b_row0 + n0 + 64 * j, kb * TC_BK);
        } else {
          ptx::tma_load_2d(b_s, &tmB, &full_bar[s], kb * TC_BK, b_row0 + n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 16; ++kk) {
          // K-major: 16 bf16 = 32 B further along the swizzled 128 B row; rows of 8 are 1024 B apart.
          // MN-major: 16 k-rows of 128 B = 2048 B further; 64-element MN groups are TC_BK*128 B apart.
          const uint64_t adesc = A_MN ? ptx::make_smem_desc_sw128(a_addr + kk * 2048, TC_BK * 128, 1024)
                                      : ptx::make_smem_desc_sw128(a_addr + kk * 32, 16, 1024);
          const uint64_t bdesc = B_MN ? ptx::make_smem_desc_sw128(b_addr + kk * 2048, TC_BK * 128, 1024)
                                      : ptx::make_smem_desc_sw128(b_addr + kk * 32, 16, 1024);
          ptx::umma_f16(tmem_base, adesc, bdesc, idesc, (kb | kk) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lanes [32q, 32q+32)
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    // Phase 1: each warp drains its 32 TMEM lanes (tile rows), applies bias / activation in registers and parks
    // the finished values in shared memory (the TMA ring is idle by now: every k-block has been consumed).
    // Phase 2: the 128 epilogue threads write the tile out ROW-CONTIGUOUSLY — a warp stores 512 contiguous bytes
    // per instruction instead of 32 scattered 16-byte pieces.
    constexpr int LD = BN + 4;                       // +4 floats: conflict-free float4 row-per-lane writes
    float* stage = reinterpret_cast<float*>(smem);
    const int r_loc = q * 32 + lane;
    const int64_t row = (int64_t)m0 + r_loc;
    const bool row_ok = row < M;
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      uint32_t r[16];
      ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), r);
      ptx::tmem_ld_wait();
      const int nb = n0 + c * 16;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      if (ep.bias) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (nb + i < N) v[i] += __ldg(ep.bias + nb + i);
      }
      if (ep.aux && row_ok) {                          // pre-activation (small encoder GEMMs only): direct store
        const int64_t o = row * ep.ldc + nb;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (nb + i < N) ep.aux[o + i] = v[i];
      }
      if (ep.epilogue != ARK_EPI_NONE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], ep.epilogue);
      }
      float* sp = stage + r_loc * LD + c * 16;
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(sp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int tid = threadIdx.x - 64;
    const bool vec_ok = (ep.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.C) & 15) == 0);
    constexpr int F4_PER_ROW = BN / 4;
#pragma unroll 4
    for (int item = tid; item < TC_BM * F4_PER_ROW; item += 128) {
      const int rl = item / F4_PER_ROW, c4 = item % F4_PER_ROW;
      const int64_t grow = (int64_t)m0 + rl;
      const int col = n0 + c4 * 4;
      if (grow >= M || col >= N) continue;
      const float4 w = *reinterpret_cast<const float4*>(stage + rl * LD + c4 * 4);
      const int64_t o = grow * ep.ldc + col;
      if (ep.c_bf16) {
        uint16_t* cp = reinterpret_cast<uint16_t*>(ep.C) + o;
        if (vec_ok && col + 4 <= N) {
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
        if (vec_ok && col + 4 <= N) {
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
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN, int STAGES>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& ep, int M, int N, int K,
                     int a_row0, int b_row0, cudaStream_t s) {
  using L = TcSmem<BN, STAGES>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, STAGES>;
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return fail((int)e, "gemm_bf16_tc: smem attribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const unsigned nt = (unsigned)((N + BN - 1) / BN), mt = (unsigned)((M + TC_BM - 1) / TC_BM);
  EpiParams ep2 = ep;
  ep2.swap_raster = (M < N && nt <= 65535u) ? 1 : 0;
  dim3 grid(ep2.swap_raster ? mt : nt, ep2.swap_raster ? nt : mt);
  kern<<<grid, 192, L::TOTAL, s>>>(tmA, tmB, ep2, M, N, K, a_row0, b_row0);
  return launched("gemm_bf16_tc");
}

template <int BN, int STAGES>
static int dispatch_major(int a_major, int b_major, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const EpiParams& ep, int M, int N, int K, int a_row0, int b_row0, cudaStream_t s) {
  if (a_major == ARK_MAJOR_K && b_major == ARK_MAJOR_K)
    return launch_tc<BN, false, false, STAGES>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  if (a_major == ARK_MAJOR_K) return launch_tc<BN, false, true, STAGES>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  if (b_major == ARK_MAJOR_K) return launch_tc<BN, true, false, STAGES>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  return launch_tc<BN, true, true, STAGES>(tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
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
  if (BN == 128) return dispatch_major<128, 3>(a_major, b_major, tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
  return dispatch_major<64, 4>(a_major, b_major, tmA, tmB, ep, M, N, K, a_row0, b_row0, s);
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
  ARK_REQUIRE(epilogue >= ARK_EPI_NONE && epilogue <= ARK_EPI_TANH, ARK_E_BADARG, "gemm_bf16_tc: bad epilogue");
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
