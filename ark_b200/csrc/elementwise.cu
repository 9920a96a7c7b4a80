// Elementwise / reduction helpers of the ELBO step: activation backward, bias gradients (column sums),
// residual add, casts, inter-layer dropout, and K10 dense Adam over the flat parameter buffer
// (torch.optim.Adam defaults, kgvae/experiments/ablation_study.py:571).  All HBM-bound, 128-bit accesses.
#include "common.cuh"
#include "philox.cuh"

namespace ark {

template <int MODE>  // 0: gelu'(aux)   1: 1 - aux^2 (tanh, aux = output)
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dact, const float* __restrict__ aux,
                                                      int64_t n, float* __restrict__ dpre,
                                                      uint16_t* __restrict__ dpre_bf16) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float a[4], g[4], o[4];
  if (i + 4 <= n) {
    const float4 av = *reinterpret_cast<const float4*>(aux + i), gv = *reinterpret_cast<const float4*>(dact + i);
    a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
    g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
  } else {
    for (int k = 0; k < 4; ++k) {
      a[k] = (i + k < n) ? aux[i + k] : 0.f;
      g[k] = (i + k < n) ? dact[i + k] : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = g[k] * (MODE == 0 ? gelu_erf_grad(a[k]) : (1.f - a[k] * a[k]));
  if (i + 4 <= n) {
    if (dpre) *reinterpret_cast<float4*>(dpre + i) = make_float4(o[0], o[1], o[2], o[3]);
    if (dpre_bf16) {
      uint2 p;
      p.x = pack_bf16x2(o[0], o[1]);
      p.y = pack_bf16x2(o[2], o[3]);
      *reinterpret_cast<uint2*>(dpre_bf16 + i) = p;
    }
  } else {
    for (int k = 0; k < 4 && i + k < n; ++k) {
      if (dpre) dpre[i + k] = o[k];
      if (dpre_bf16) dpre_bf16[i + k] = f32_to_bf16_bits(o[k]);
    }
  }
}

// column sums: block = 32 x 8 threads, each thread owns one column (stride-1 across lanes, coalesced) and
// walks rows ty, ty+8*gridDim.y, ...; 8 partials reduced through smem, one atomicAdd per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, int64_t M, int N, int64_t ld,
                                                     float* __restrict__ out) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (col < N) {
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < M; r += (int64_t)gridDim.y * 8) {
      if constexpr (sizeof(T) == 2)
        acc += bf16_bits_to_f32(X[r * ld + col]);
      else
        acc += X[r * ld + col];
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][tx];
    atomicAdd(out + col, s);
  }
}

// vector variant: each thread owns 4 consecutive columns (8 B of bf16 / 16 B of f32 per row), a warp covers 128
// columns of a row contiguously; requires N % 4 == 0, ld % 4 == 0 and a 16-byte aligned base.
template <typename T>
__global__ void __launch_bounds__(256) colsum4_kernel(const T* __restrict__ X, int64_t M, int N, int64_t ld,
                                                      float* __restrict__ out) {
  __shared__ float4 part[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    const int64_t stride = (int64_t)gridDim.y * 8;
    int64_t r = (int64_t)blockIdx.y * 8 + ty;
    for (; r + 3 * stride < M; r += 4 * stride) {   // four independent loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if constexpr (sizeof(T) == 2) {
          const uint2 ua = *reinterpret_cast<const uint2*>(X + (r + u * stride) * ld + col);
          const float2 a0 = unpack_bf16x2(ua.x), a1 = unpack_bf16x2(ua.y);
          v[u] = make_float4(a0.x, a0.y, a1.x, a1.y);
        } else {
          v[u] = *reinterpret_cast<const float4*>(X + (r + u * stride) * ld + col);
        }
      }
      acc.x += (v[0].x + v[1].x) + (v[2].x + v[3].x); acc.y += (v[0].y + v[1].y) + (v[2].y + v[3].y);
      acc.z += (v[0].z + v[1].z) + (v[2].z + v[3].z); acc.w += (v[0].w + v[1].w) + (v[2].w + v[3].w);
    }
    for (; r + stride < M; r += 2 * stride) {       // two independent loads in flight
      float4 a, b;
      if constexpr (sizeof(T) == 2) {
        const uint2 ua = *reinterpret_cast<const uint2*>(X + r * ld + col);
        const uint2 ub = *reinterpret_cast<const uint2*>(X + (r + stride) * ld + col);
        const float2 a0 = unpack_bf16x2(ua.x), a1 = unpack_bf16x2(ua.y), b0 = unpack_bf16x2(ub.x), b1 = unpack_bf16x2(ub.y);
        a = make_float4(a0.x, a0.y, a1.x, a1.y);
        b = make_float4(b0.x, b0.y, b1.x, b1.y);
      } else {
        a = *reinterpret_cast<const float4*>(X + r * ld + col);
        b = *reinterpret_cast<const float4*>(X + (r + stride) * ld + col);
      }
      acc.x += a.x + b.x; acc.y += a.y + b.y; acc.z += a.z + b.z; acc.w += a.w + b.w;
    }
    for (; r < M; r += stride) {
      float4 a;
      if constexpr (sizeof(T) == 2) {
        const uint2 ua = *reinterpret_cast<const uint2*>(X + r * ld + col);
        const float2 a0 = unpack_bf16x2(ua.x), a1 = unpack_bf16x2(ua.y);
        a = make_float4(a0.x, a0.y, a1.x, a1.y);
      } else {
        a = *reinterpret_cast<const float4*>(X + r * ld + col);
      }
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < N) {
    float4 s4 = part[0][tx];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      s4.x += part[k][tx].x; s4.y += part[k][tx].y; s4.z += part[k][tx].z; s4.w += part[k][tx].w;
    }
    atomicAdd(out + col, s4.x);
    if (col + 1 < N) atomicAdd(out + col + 1, s4.y);    // N % 4 != 0: the last thread's tail columns are row padding
    if (col + 2 < N) atomicAdd(out + col + 2, s4.z);
    if (col + 3 < N) atomicAdd(out + col + 3, s4.w);
  }
}

__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                  float* __restrict__ y, uint16_t* __restrict__ yb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = a[i] + (b ? b[i] : 0.f);
  if (y) y[i] = v;
  if (yb) yb[i] = f32_to_bf16_bits(v);
}


__global__ void __launch_bounds__(256) dropout_fwd_kernel(const float* __restrict__ x, int64_t n, float p,
                                                          uint64_t seed, uint64_t offset, float* __restrict__ y,
                                                          uint16_t* __restrict__ yb, uint8_t* __restrict__ mask,
                                                          const uint64_t* __restrict__ offset_dev) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = q * 4;
  if (i >= n) return;
  const uint64_t c = offset + (offset_dev ? *offset_dev : 0ull) + (uint64_t)q;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
  const float scale = 1.f / (1.f - p);
  for (int k = 0; k < 4 && i + k < n; ++k) {
    const float u = (float)(rr[k] >> 8) * (1.f / 16777216.f);  // [0,1)
    const bool keep = u >= p;
    const float v = keep ? x[i + k] * scale : 0.f;
    if (mask) mask[i + k] = keep ? 1 : 0;
    if (y) y[i + k] = v;
    if (yb) yb[i + k] = f32_to_bf16_bits(v);
  }
}

__global__ void __launch_bounds__(256) dropout_bf16_kernel(const uint16_t* __restrict__ x, int64_t n, float p,
                                                           uint64_t seed, uint64_t offset, uint16_t* __restrict__ y,
                                                           uint8_t* __restrict__ mask,
                                                           const uint64_t* __restrict__ offset_dev) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = q * 4;
  if (i >= n) return;
  const uint64_t c = offset + (offset_dev ? *offset_dev : 0ull) + (uint64_t)q;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
  const float scale = 1.f / (1.f - p);
  if (i + 4 <= n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 7) == 0 &&
      (!mask || (reinterpret_cast<uintptr_t>(mask) & 3) == 0)) {
    // vector path: one 8-byte load / store of the four bf16 values and one 4-byte store of their keep flags
    const uint2 xv = *reinterpret_cast<const uint2*>(x + i);
    const float2 a = unpack_bf16x2(xv.x), b = unpack_bf16x2(xv.y);
    const float xs[4] = {a.x, a.y, b.x, b.y};
    uint16_t o[4];
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool keep = (float)(rr[k] >> 8) * (1.f / 16777216.f) >= p;
      m |= (keep ? 1u : 0u) << (8 * k);
      o[k] = keep ? f32_to_bf16_bits(xs[k] * scale) : (uint16_t)0;
    }
    if (mask) *reinterpret_cast<uint32_t*>(mask + i) = m;
    uint2 yv;
    yv.x = (uint32_t)o[0] | ((uint32_t)o[1] << 16);
    yv.y = (uint32_t)o[2] | ((uint32_t)o[3] << 16);
    *reinterpret_cast<uint2*>(y + i) = yv;
    return;
  }
  for (int k = 0; k < 4 && i + k < n; ++k) {
    const float u = (float)(rr[k] >> 8) * (1.f / 16777216.f);
    const bool keep = u >= p;
    if (mask) mask[i + k] = keep ? 1 : 0;
    y[i + k] = keep ? f32_to_bf16_bits(bf16_bits_to_f32(x[i + k]) * scale) : (uint16_t)0;
  }
}

__global__ void __launch_bounds__(256) dropout_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask,
                                                          int64_t n, float scale, float* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dx[i] = mask[i] ? dy[i] * scale : 0.f;
}
// four elements per thread: 16-byte gradient accesses, one 4-byte load of the keep flags (n % 4 == 0, aligned bases)
__global__ void __launch_bounds__(256) dropout_bwd4_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask,
                                                           int64_t n4, float scale, float* __restrict__ dx) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n4) return;
  const float4 g = *reinterpret_cast<const float4*>(dy + 4 * q);
  const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + 4 * q);
  float4 o;
  o.x = (m & 0xffu) ? g.x * scale : 0.f;
  o.y = (m & 0xff00u) ? g.y * scale : 0.f;
  o.z = (m & 0xff0000u) ? g.z * scale : 0.f;
  o.w = (m & 0xff000000u) ? g.w * scale : 0.f;
  *reinterpret_cast<float4*>(dx + 4 * q) = o;
}

// Adam: 16 B p + 16 B g + 16 B m + 16 B v read, 16+16+16 written (+8 B bf16 shadow) per 4 parameters.
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        uint16_t* __restrict__ shadow, int64_t n, float step_size,
                                                        float beta1, float beta2, float eps, float inv_sqrt_bc2,
                                                        float grad_scale, const float* __restrict__ hyper) {
  if (hyper) {   // CUDA-graph replay: step-dependent scalars live in device memory
    step_size = hyper[0];
    inv_sqrt_bc2 = hyper[1];
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
      float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vv4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = gg[k] * grad_scale;
        mm[k] = beta1 * mm[k] + (1.f - beta1) * gk;
        vv4[k] = beta2 * vv4[k] + (1.f - beta2) * gk * gk;
        pp[k] -= step_size * mm[k] / (sqrtf(vv4[k]) * inv_sqrt_bc2 + eps);
      }
      *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv4[0], vv4[1], vv4[2], vv4[3]);
      if (shadow) {
        uint2 s;
        s.x = pack_bf16x2(pp[0], pp[1]);
        s.y = pack_bf16x2(pp[2], pp[3]);
        *reinterpret_cast<uint2*>(shadow + i) = s;
      }
    } else {
      for (int64_t j = i; j < n; ++j) {
        const float gk = g[j] * grad_scale;
        const float mj = beta1 * m[j] + (1.f - beta1) * gk;
        const float vj = beta2 * v[j] + (1.f - beta2) * gk * gk;
        const float pj = p[j] - step_size * mj / (sqrtf(vj) * inv_sqrt_bc2 + eps);
        m[j] = mj; v[j] = vj; p[j] = pj;
        if (shadow) shadow[j] = f32_to_bf16_bits(pj);
      }
    }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, int64_t n, uint16_t* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      uint2 s;
      s.x = pack_bf16x2(v.x, v.y);
      s.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(y + i) = s;
    } else {
      for (int64_t j = i; j < n; ++j) y[j] = f32_to_bf16_bits(x[j]);
    }
  }
}

}  // namespace ark

using namespace ark;

static inline unsigned blocks_for(int64_t n, int per_block, int64_t cap) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (cap > 0 && b > cap) b = cap;
  return (unsigned)b;
}

extern "C" int ark_gelu_bwd(const float* dact, const float* pre, int64_t n, float* dpre, uint16_t* dpre_bf16,
                            void* stream) {
  ARK_REQUIRE(dact && pre && (dpre || dpre_bf16), ARK_E_BADARG, "gelu_bwd: null pointer");
  ARK_REQUIRE(aligned16(dact) && aligned16(pre) && (!dpre || aligned16(dpre)) && (!dpre_bf16 || aligned16(dpre_bf16)),
              ARK_E_ALIGN, "gelu_bwd: 16-byte alignment");
  if (n <= 0) return 0;
  act_bwd_kernel<0><<<blocks_for(n, 1024, 0), 256, 0, (cudaStream_t)stream>>>(dact, pre, n, dpre, dpre_bf16);
  return launched("gelu_bwd");
}

extern "C" int ark_tanh_bwd(const float* dh, const float* h, int64_t n, float* dpre, uint16_t* dpre_bf16,
                            void* stream) {
  ARK_REQUIRE(dh && h && (dpre || dpre_bf16), ARK_E_BADARG, "tanh_bwd: null pointer");
  ARK_REQUIRE(aligned16(dh) && aligned16(h) && (!dpre || aligned16(dpre)) && (!dpre_bf16 || aligned16(dpre_bf16)),
              ARK_E_ALIGN, "tanh_bwd: 16-byte alignment");
  if (n <= 0) return 0;
  act_bwd_kernel<1><<<blocks_for(n, 1024, 0), 256, 0, (cudaStream_t)stream>>>(dh, h, n, dpre, dpre_bf16);
  return launched("tanh_bwd");
}

extern "C" int ark_colsum(const void* X, int dtype, int64_t M, int64_t N, int64_t ld, float* out, int accumulate,
                          void* stream) {
  ARK_REQUIRE(X && out, ARK_E_BADARG, "colsum: null pointer");
  ARK_REQUIRE(M >= 0 && N > 0 && ld >= N, ARK_E_BADARG, "colsum: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  if (accumulate != 1) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s);
    if (e != cudaSuccess) return fail((int)e, "colsum: memset: %s", cudaGetErrorString(e));
  }
  if (M == 0) return 0;
  // vector path: 4 columns per thread; a ragged N is fine when the row padding up to the next multiple of 4 exists
  const bool vec = (ld % 4 == 0) && aligned16(X) && ((N + 3) / 4 * 4 <= ld);
  const int64_t gx = vec ? (N + 127) / 128 : (N + 31) / 32;
  int64_t gy = (M + 63) / 64;
  const int64_t want = 4 * kNumSMs;  // enough CTAs to cover the machine, few enough atomics
  if (gx * gy > want) gy = (want + gx - 1) / gx;
  if (gy < 1) gy = 1;
  if (accumulate == 2) gy = 1;       // deterministic: one CTA per column strip, fixed summation order (no cross-CTA atomics)
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (dtype == ARK_BF16) {
    if (vec) colsum4_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)X, M, (int)N, ld, out);
    else colsum_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)X, M, (int)N, ld, out);
  } else if (dtype == ARK_F32) {
    if (vec) colsum4_kernel<float><<<grid, 256, 0, s>>>((const float*)X, M, (int)N, ld, out);
    else colsum_kernel<float><<<grid, 256, 0, s>>>((const float*)X, M, (int)N, ld, out);
  } else {
    return fail(ARK_E_BADARG, "colsum: unknown dtype %d", dtype);
  }
  return launched("colsum");
}

extern "C" int ark_add_f32(const float* a, const float* b, int64_t n, float* y, uint16_t* y_bf16, void* stream) {
  ARK_REQUIRE(a && (y || y_bf16), ARK_E_BADARG, "add_f32: null pointer");
  if (n <= 0) return 0;
  add_kernel<<<blocks_for(n, 256, 0), 256, 0, (cudaStream_t)stream>>>(a, b, n, y, y_bf16);
  return launched("add_f32");
}

extern "C" int ark_cast_f32_to_bf16(const float* x, int64_t n, uint16_t* y, void* stream) {
  ARK_REQUIRE(x && y, ARK_E_BADARG, "cast: null pointer");
  ARK_REQUIRE(aligned16(x) && aligned16(y), ARK_E_ALIGN, "cast: 16-byte alignment");
  if (n <= 0) return 0;
  cast_bf16_kernel<<<blocks_for(n, 1024, 8 * kNumSMs), 256, 0, (cudaStream_t)stream>>>(x, n, y);
  return launched("cast_f32_to_bf16");
}

extern "C" int ark_dropout_fwd(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset, float* y,
                               uint16_t* y_bf16, uint8_t* mask, const uint64_t* offset_dev, void* stream) {
  ARK_REQUIRE(x && (y || y_bf16), ARK_E_BADARG, "dropout_fwd: null pointer");
  ARK_REQUIRE(p >= 0.f && p < 1.f, ARK_E_BADARG, "dropout_fwd: p must be in [0,1)");
  if (n <= 0) return 0;
  dropout_fwd_kernel<<<blocks_for((n + 3) / 4, 256, 0), 256, 0, (cudaStream_t)stream>>>(x, n, p, seed, offset, y,
                                                                                       y_bf16, mask, offset_dev);
  return launched("dropout_fwd");
}

extern "C" int ark_dropout_bf16(const uint16_t* x, int64_t n, float p, uint64_t seed, uint64_t offset, uint16_t* y,
                                uint8_t* mask, const uint64_t* offset_dev, void* stream) {
  ARK_REQUIRE(x && y, ARK_E_BADARG, "dropout_bf16: null pointer");
  ARK_REQUIRE(p >= 0.f && p < 1.f, ARK_E_BADARG, "dropout_bf16: p must be in [0,1)");
  if (n <= 0) return 0;
  dropout_bf16_kernel<<<blocks_for((n + 3) / 4, 256, 0), 256, 0, (cudaStream_t)stream>>>(x, n, p, seed, offset, y, mask,
                                                                                        offset_dev);
  return launched("dropout_bf16");
}

extern "C" int ark_dropout_bwd(const float* dy, const uint8_t* mask, int64_t n, float p, float* dx, void* stream) {
  ARK_REQUIRE(dy && mask && dx, ARK_E_BADARG, "dropout_bwd: null pointer");
  ARK_REQUIRE(p >= 0.f && p < 1.f, ARK_E_BADARG, "dropout_bwd: p must be in [0,1)");
  if (n <= 0) return 0;
  if (n % 4 == 0 && aligned16(dy) && aligned16(dx) && (reinterpret_cast<uintptr_t>(mask) & 3) == 0)
    dropout_bwd4_kernel<<<blocks_for(n / 4, 256, 0), 256, 0, (cudaStream_t)stream>>>(dy, mask, n / 4, 1.f / (1.f - p), dx);
  else
    dropout_bwd_kernel<<<blocks_for(n, 256, 0), 256, 0, (cudaStream_t)stream>>>(dy, mask, n, 1.f / (1.f - p), dx);
  return launched("dropout_bwd");
}

extern "C" int ark_adam_flat(float* p, const float* g, float* m, float* v, uint16_t* shadow, int64_t n, float lr,
                             float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream) {
  ARK_REQUIRE(p && g && m && v, ARK_E_BADARG, "adam_flat: null pointer");
  ARK_REQUIRE(step >= 1, ARK_E_BADARG, "adam_flat: step is 1-based");
  ARK_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!shadow || aligned16(shadow)),
              ARK_E_ALIGN, "adam_flat: 16-byte alignment");
  if (n <= 0) return 0;
  // torch.optim.Adam: step_size = lr / bc1; denom = sqrt(v)/sqrt(bc2) + eps   (computed in double on the host)
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  adam_flat_kernel<<<blocks_for(n, 1024, 16 * kNumSMs), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, shadow, n, step_size, beta1, beta2, eps, inv_sqrt_bc2, grad_scale, nullptr);
  return launched("adam_flat");
}

extern "C" int ark_adam_flat_dyn(float* p, const float* g, float* m, float* v, uint16_t* shadow, int64_t n,
                                 const float* hyper, float beta1, float beta2, float eps, float grad_scale,
                                 void* stream) {
  ARK_REQUIRE(p && g && m && v && hyper, ARK_E_BADARG, "adam_flat_dyn: null pointer");
  ARK_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!shadow || aligned16(shadow)),
              ARK_E_ALIGN, "adam_flat_dyn: 16-byte alignment");
  if (n <= 0) return 0;
  adam_flat_kernel<<<blocks_for(n, 1024, 16 * kNumSMs), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, shadow, n, 0.f, beta1, beta2, eps, 1.f, grad_scale, hyper);
  return launched("adam_flat_dyn");
}
