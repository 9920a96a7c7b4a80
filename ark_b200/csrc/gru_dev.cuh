// Device helpers shared by the persistent GRU kernels (gru_persist.cu: one layer per launch;
// gru_wave.cu: all layers in one launch as a wavefront).
#pragma once
#include "common.cuh"
#include "gru_math.cuh"

namespace ark {

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int32_t* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int32_t* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Poll with RELAXED loads and order once at the end (one acquire fence): an acquire load per poll costs a fence per
// round trip on the recurrent chain.
__device__ __forceinline__ void wait_counter(const int32_t* p, int target) {
  for (ptx::SpinGuard g; ld_relaxed(p) < target;) {
    if (g.expired()) {
      printf("arkb200: gru_persist tile counter timed out (block %d,%d want %d have %d)\n", blockIdx.x, blockIdx.y,
             target, ld_acquire(p));
      __trap();
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void ld8_bf16(const uint16_t* p, float* x) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), e = unpack_bf16x2(v.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = e.x; x[7] = e.y;
}
__device__ __forceinline__ void st16_bf16(uint16_t* p, const float* x) {
  uint4 a, b;
  a.x = pack_bf16x2(x[0], x[1]); a.y = pack_bf16x2(x[2], x[3]); a.z = pack_bf16x2(x[4], x[5]); a.w = pack_bf16x2(x[6], x[7]);
  b.x = pack_bf16x2(x[8], x[9]); b.y = pack_bf16x2(x[10], x[11]); b.z = pack_bf16x2(x[12], x[13]); b.w = pack_bf16x2(x[14], x[15]);
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + 8) = b;
}
__device__ __forceinline__ void ld16_f32(const float* p, float* x) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + i);
    x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
  }
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// same expression tree as gru_fwd_math with MUFU tanh / sigmoid(x) = 0.5 + 0.5 tanh(x/2)
__device__ __forceinline__ GruFwd gru_fwd_math_fast(float gi_r, float gi_z, float gi_n, float gh_r, float gh_z,
                                                    float gh_n, float h_prev) {
  GruFwd o;
  o.r = fmaf(0.5f, tanh_fast(0.5f * (gi_r + gh_r)), 0.5f);
  o.z = fmaf(0.5f, tanh_fast(0.5f * (gi_z + gh_z)), 0.5f);
  o.ghn = gh_n;
  o.n = tanh_fast(fmaf(o.r, gh_n, gi_n));
  o.h = fmaf(o.z, h_prev - o.n, o.n);
  return o;
}
__device__ __forceinline__ void st4_bf16(uint16_t* p, const float* x) {
  uint2 v;
  v.x = pack_bf16x2(x[0], x[1]);
  v.y = pack_bf16x2(x[2], x[3]);
  *reinterpret_cast<uint2*>(p) = v;
}
__device__ __forceinline__ void ld4_bf16(const uint16_t* p, float* x) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}

}  // namespace ark
