// Probe: cycles per tcgen05.mma (kind::f16, cta_group::1, operands in shared memory, 128B swizzle, K-major) as a function
// of the instruction shape, issued back to back by one thread.  Separates a fixed per-instruction cost from the
// shared-memory operand bandwidth.   nvcc -arch=sm_100a -I ark_b200/csrc tools/probe_mma_rate.cu
#include <cstdio>
#include <cstdint>
#include "ptx.cuh"
using namespace ark;

__global__ void probe(int M, int N, int n_mma, int distinct, int nacc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_ptr, 256); ptx::tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(M, N, 0, 0);
    const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 64 * 1024;    // A: up to 4 tiles of 16 KB; B: up to 3 tiles of 32 KB
    // descriptors precomputed: the timed loop is nothing but tcgen05.mma issues (4 k-steps x nacc accumulators)
    uint64_t ad[4], bd[4];
    for (int kk = 0; kk < 4; ++kk) {
      ad[kk] = ptx::make_smem_desc_sw128(a0 + kk * 32, 16, 1024);
      bd[kk] = ptx::make_smem_desc_sw128(b0 + kk * 32, 16, 1024);
    }
    const uint32_t d0 = tb, d1 = tb + (uint32_t)(nacc > 1 ? N : 0);
    ptx::umma_f16(d0, ad[0], bd[0], idesc, 0u);
    ptx::umma_f16(d1, ad[0], bd[0], idesc, 0u);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n_mma; i += 8) {
      ptx::umma_f16(d0, ad[0], bd[0], idesc, 1u);
      ptx::umma_f16(d1, ad[1], bd[1], idesc, 1u);
      ptx::umma_f16(d0, ad[2], bd[2], idesc, 1u);
      ptx::umma_f16(d1, ad[3], bd[3], idesc, 1u);
      ptx::umma_f16(d0, ad[0], bd[0], idesc, 1u);
      ptx::umma_f16(d1, ad[1], bd[1], idesc, 1u);
      ptx::umma_f16(d0, ad[2], bd[2], idesc, 1u);
      ptx::umma_f16(d1, ad[3], bd[3], idesc, 1u);
    }
    const long long t1 = clock64();
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int Ms[2] = {128, 64};
  const int Ns[7] = {16, 48, 64, 96, 128, 192, 256};
  const int n = 512;
  for (int mi = 0; mi < 2; ++mi)
    for (int ni = 0; ni < 7; ++ni) {
      for (int nacc = 1; nacc <= 2 && nacc * Ns[ni] <= 256; nacc *= 2) {
        const int distinct = 3;
        probe<<<1, 128, 200 * 1024>>>(Ms[mi], Ns[ni], n, distinct, nacc, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double per = (double)h[1] / n, math = (double)Ms[mi] * Ns[ni] * 16 / 4096.0;
        printf("M=%3d N=%3d nacc=%d: issue %.1f cyc/mma, complete %.1f cyc/mma (math at peak %.0f, operand bytes %d -> %.0f B/clk) %s\n",
               Ms[mi], Ns[ni], nacc, (double)h[0] / n, per, math, (Ms[mi] + Ns[ni]) * 32, (Ms[mi] + Ns[ni]) * 32 / per,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  return 0;
}
