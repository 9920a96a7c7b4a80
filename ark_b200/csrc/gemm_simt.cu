// K3 (SIMT variant): C[M,N] = epi(A[M,K] . B[N,K]^T + bias) with fp32 FMA accumulation for ANY alignment,
// leading dimension, operand major and tiny shape (K = d_latent = 10, N = 2*d_latent, M = 16 ...), and
// for fp32 operands (the drop-in fp32 inference path that reproduces the reference's sampled graphs).
// The bf16 training hot path uses the tcgen05 kernel in gemm_tc.cu; this kernel shares its contract so
// the two can be cross-checked element by element on the GPU.
#include "common.cuh"

namespace ark {

constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<uint16_t>(uint16_t v) { return bf16_bits_to_f32(v); }

// loads a [64 x 16] (rows x k) operand tile into smem as S[k][row]
template <typename T, bool MN_MAJOR>
__device__ __forceinline__ void load_tile(const T* __restrict__ P, int64_t ld, int64_t row0, int64_t nrows, int64_t k0,
                                          int64_t K, float (*S)[SB_M + 1]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r, k;
    if (MN_MAJOR) {
      r = tid & 63;
      k = (tid >> 6) + 4 * i;
    } else {
      k = tid & 15;
      r = (tid >> 4) + 16 * i;
    }
    const int64_t gr = row0 + r, gk = k0 + k;
    float v = 0.f;
    if (gr < nrows && gk < K) v = to_f32<T>(MN_MAJOR ? P[gk * ld + gr] : P[gr * ld + gk]);
    S[k][r] = v;
  }
}

template <typename T, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B,
                                                        int64_t ldb, void* __restrict__ C, int c_bf16, int64_t ldc,
                                                        int64_t M, int64_t N, int64_t K,
                                                        const float* __restrict__ bias, int epilogue, int accumulate,
                                                        float* __restrict__ aux) {
  __shared__ float As[SB_K][SB_M + 1];
  __shared__ float Bs[SB_K][SB_N + 1];
  const int64_t m0 = (int64_t)blockIdx.y * SB_M, n0 = (int64_t)blockIdx.x * SB_N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += SB_K) {
    load_tile<T, A_MN>(A, lda, m0, M, k0, K, As);
    load_tile<T, B_MN>(B, ldb, n0, N, k0, K, Bs);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (aux) aux[m * ldc + n] = v;
      if (epilogue == ARK_EPI_GELU) v = gelu_erf(v);
      else if (epilogue == ARK_EPI_TANH) v = tanhf(v);
      else if (epilogue == ARK_EPI_RELU) v = fmaxf(v, 0.f);
      if (c_bf16) {
        reinterpret_cast<uint16_t*>(C)[m * ldc + n] = f32_to_bf16_bits(v);
      } else {
        float* c = reinterpret_cast<float*>(C) + m * ldc + n;
        *c = accumulate ? (*c + v) : v;
      }
    }
  }
}

template <typename T>
static int launch_simt(const T* A, int a_major, int64_t lda, const T* B, int b_major, int64_t ldb, void* C, int c_dtype,
                       int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias, int epilogue, int accumulate,
                       float* aux, cudaStream_t s) {
  dim3 grid((unsigned)((N + SB_N - 1) / SB_N), (unsigned)((M + SB_M - 1) / SB_M));
  const int cb = c_dtype == ARK_BF16;
#define ARK_SIMT_GO(AM, BM) \
  gemm_simt_kernel<T, AM, BM><<<grid, 256, 0, s>>>(A, lda, B, ldb, C, cb, ldc, M, N, K, bias, epilogue, accumulate, aux)
  if (a_major == ARK_MAJOR_K && b_major == ARK_MAJOR_K) ARK_SIMT_GO(false, false);
  else if (a_major == ARK_MAJOR_K) ARK_SIMT_GO(false, true);
  else if (b_major == ARK_MAJOR_K) ARK_SIMT_GO(true, false);
  else ARK_SIMT_GO(true, true);
#undef ARK_SIMT_GO
  return launched("gemm_simt");
}

}  // namespace ark

using namespace ark;

extern "C" int ark_gemm_simt(const void* A, int a_major, int64_t lda, const void* B, int b_major, int64_t ldb,
                             int ab_dtype, void* C, int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K,
                             const float* bias, int epilogue, int accumulate, float* aux, void* stream) {
  ARK_REQUIRE(A && B && C, ARK_E_BADARG, "gemm_simt: null pointer");
  ARK_REQUIRE(M >= 0 && N >= 0 && K >= 0, ARK_E_BADARG, "gemm_simt: negative size");
  ARK_REQUIRE((a_major == ARK_MAJOR_K || a_major == ARK_MAJOR_MN) && (b_major == ARK_MAJOR_K || b_major == ARK_MAJOR_MN),
              ARK_E_BADARG, "gemm_simt: bad major");
  ARK_REQUIRE(lda >= (a_major == ARK_MAJOR_K ? K : M) && ldb >= (b_major == ARK_MAJOR_K ? K : N) && ldc >= N,
              ARK_E_BADARG, "gemm_simt: leading dimension too small");
  ARK_REQUIRE(c_dtype == ARK_F32 || c_dtype == ARK_BF16, ARK_E_BADARG, "gemm_simt: bad c_dtype");
  ARK_REQUIRE(!(accumulate && c_dtype != ARK_F32), ARK_E_BADARG, "gemm_simt: accumulate needs f32 C");
  ARK_REQUIRE(epilogue >= ARK_EPI_NONE && epilogue <= ARK_EPI_RELU, ARK_E_BADARG, "gemm_simt: bad epilogue");
  if (M == 0 || N == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (ab_dtype == ARK_F32)
    return launch_simt<float>((const float*)A, a_major, lda, (const float*)B, b_major, ldb, C, c_dtype, ldc, M, N, K,
                              bias, epilogue, accumulate, aux, s);
  if (ab_dtype == ARK_BF16)
    return launch_simt<uint16_t>((const uint16_t*)A, a_major, lda, (const uint16_t*)B, b_major, ldb, C, c_dtype, ldc,
                                 M, N, K, bias, epilogue, accumulate, aux, s);
  return fail(ARK_E_BADARG, "gemm_simt: unknown ab_dtype %d", ab_dtype);
}
