// Probe: where do the 64 rows of an M = 64 tcgen05.mma (cta_group::1, kind::f16) accumulator land in tensor memory?
// D[m][n] = A[m][:] . B[n][:] with A[m][k] = (k == 0) * (m + 1), B[n][k] = (k == 0)  ->  D[m][n] = m + 1.
// Every TMEM lane is zeroed first; the dump shows lane -> row.   nvcc -arch=sm_100a -I ark_b200/csrc tools/probe_m64.cu
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"
using namespace ark;

__device__ __forceinline__ uint16_t bf(float x) { return (uint16_t)(__float_as_uint(x) >> 16); }

template <int M>
__global__ void probe(float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint16_t* a_sm = reinterpret_cast<uint16_t*>(smem);            // [128 rows][64 k] bf16, 128B swizzle
  uint16_t* b_sm = reinterpret_cast<uint16_t*>(smem + 16384);    // [16 rows][64 k]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 64 + (((c / 8) ^ (r % 8)) * 8) + (c % 8);
    a_sm[off] = (c == 0) ? bf((float)(r + 1)) : 0;
  }
  for (int i = tid; i < 16 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 64 + (((c / 8) ^ (r % 8)) * 8) + (c % 8);
    b_sm[off] = (c == 0) ? bf(1.f) : 0;
  }
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_ptr, 32); ptx::tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  {   // zero 16 columns of every lane
    uint32_t z[16];
    for (int k = 0; k < 16; ++k) z[k] = 0;
    ptx::tmem_st_32x32b_x16(tb + ((uint32_t)(warp * 32) << 16), z);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(M, 16, 0, 0);
    const uint64_t ad = ptx::make_smem_desc_sw128(ptx::smem_u32(a_sm), 16, 1024);
    const uint64_t bd = ptx::make_smem_desc_sw128(ptx::smem_u32(b_sm), 16, 1024);
    ptx::umma_f16(tb, ad, bd, idesc, 0u);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  uint32_t v[16];
  ptx::tmem_ld_32x32b_x16(tb + ((uint32_t)(warp * 32) << 16), v);
  ptx::tmem_ld_wait();
  for (int k = 0; k < 16; ++k) out[(warp * 32 + lane) * 16 + k] = __uint_as_float(v[k]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 32);
}

template <int M>
void run(const char* name) {
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  cudaMemset(d, 0, 128 * 16 * 4);
  cudaFuncSetAttribute(probe<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<M><<<1, 128, 32768>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s: %s\n", name, cudaGetErrorString(e));
  float h[128 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int l = 0; l < 128; ++l) {
    printf("%3d:%3.0f/%3.0f ", l, h[l * 16], h[l * 16 + 15]);
    if (l % 8 == 7) printf("\n");
  }
  cudaFree(d);
}

int main() {
  run<128>("M=128 (lane l should hold row l+1)");
  run<64>("M=64");
  return 0;
}
