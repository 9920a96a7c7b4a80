// K1/K2: entity/relation embedding gather fused with masked mean-pooling, and its scatter-add
// backward.  K4: reparameterisation + analytic KL, forward and backward.
// Replaces kgvae/model/models.py:47-58 (gather/concat/mask/mean), :62-63 (clamp, reparam) and
// :199-200 (kl_mean) of the reference.  All HBM-bound: 128-bit loads, one pass, nothing staged.
#include "common.cuh"

namespace ark {

// grid (B, 3): CTA (b, slot) pools column block `slot` (head | relation | tail) of graph perm[b].
// Threads form a [TY t-lanes] x [CX column lanes] grid: t-lane ty walks triples ty, ty+TY, ... (UNROLL independent
// 16-byte row loads in flight), column lane tx owns float4 columns tx, tx+CX, ...; the TY partial sums meet in
// shared memory.  With few graphs per GPU (wd-articles: B = 16, T = 212) the parallelism over triples is what
// keeps enough loads in flight; triples are read through L1 (broadcast across the column lanes).
constexpr int GP_MAX_THREADS = 1024;

template <int UNROLL>
__global__ void __launch_bounds__(GP_MAX_THREADS) gather_pool_fwd_kernel(
    const int64_t* __restrict__ triples, const int32_t* __restrict__ perm, const float* __restrict__ E,
    const float* __restrict__ R, int T, int d, long long pad_rid, int CX, float* __restrict__ g,
    uint16_t* __restrict__ g_bf16, float* __restrict__ inv_cnt) {
  extern __shared__ float4 gp_red[];   // [TY][CX]
  __shared__ int cnt_s;
  const int b = blockIdx.x, slot = blockIdx.y;
  const int src = perm ? perm[b] : b;
  const int64_t* tri = triples + (int64_t)src * T * 3;
  const float* table = (slot == 1) ? R : E;
  const int d4 = d >> 2;
  const int tx = threadIdx.x % CX, ty = threadIdx.x / CX, TY = blockDim.x / CX;

  if (threadIdx.x == 0) cnt_s = 0;
  __syncthreads();
  if (pad_rid >= 0) {   // number of valid triples: one warp-aggregated atomic per warp
    int c = 0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) c += (tri[t * 3 + 1] != pad_rid);
    c = (int)warp_sum((float)c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&cnt_s, c);
  } else if (threadIdx.x == 0) {
    cnt_s = T;
  }
  __syncthreads();
  const int cnt = cnt_s;
  const float inv = 1.f / (float)(cnt > 0 ? cnt : 1);
  if (slot == 0 && threadIdx.x == 0) inv_cnt[b] = inv;

  for (int c0 = 0; c0 < d4; c0 += CX) {
    const int c = c0 + tx;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < d4) {
      int t = ty;
      for (; t + (UNROLL - 1) * TY < T; t += UNROLL * TY) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int tt = t + u * TY;
          const bool ok = (pad_rid < 0) || (tri[tt * 3 + 1] != pad_rid);
          const int64_t idx = tri[tt * 3 + slot];
          v[u] = ok ? __ldg(reinterpret_cast<const float4*>(table + idx * d) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
        }
      }
      for (; t < T; t += TY) {
        if (pad_rid >= 0 && tri[t * 3 + 1] == pad_rid) continue;
        const int64_t idx = tri[t * 3 + slot];
        const float4 v = __ldg(reinterpret_cast<const float4*>(table + idx * d) + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    gp_red[ty * CX + tx] = acc;
    __syncthreads();
    if (ty == 0 && c < d4) {
      for (int y = 1; y < TY; ++y) {
        const float4 v = gp_red[y * CX + tx];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      const int64_t o = (int64_t)b * 3 * d + (int64_t)slot * d + c * 4;
      if (g) *reinterpret_cast<float4*>(g + o) = acc;
      if (g_bf16) {
        uint2 pk;
        pk.x = pack_bf16x2(acc.x, acc.y);
        pk.y = pack_bf16x2(acc.z, acc.w);
        *reinterpret_cast<uint2*>(g_bf16 + o) = pk;
      }
    }
    __syncthreads();
  }
}

// grid (B, 3, Z): every valid triple of graph perm[b] receives the SAME slice dg[b, slot]/cnt_b, so the value is
// loaded once into registers and pushed with one 16-byte L2 reduction per (triple, float4); the Z CTAs of a
// (graph, slot) split the triples.
constexpr int kRelHist = 64;    // relation tables up to this many rows are accumulated through a per-CTA histogram
__global__ void __launch_bounds__(256) gather_pool_bwd_kernel(
    const float* __restrict__ dg, const int64_t* __restrict__ triples, const int32_t* __restrict__ perm,
    const float* __restrict__ inv_cnt, int T, int d, long long pad_rid, long long pad_eid, int n_rel,
    float* __restrict__ dE, float* __restrict__ dR) {
  const int b = blockIdx.x, slot = blockIdx.y;
  const int src = perm ? perm[b] : b;
  const int64_t* tri = triples + (int64_t)src * T * 3;
  float* table = (slot == 1) ? dR : dE;
  const int d4 = d >> 2;
  const float inv = inv_cnt[b];
  if (slot == 1 && n_rel > 0 && n_rel <= kRelHist) {
    // the relation table has a handful of rows (3..7 in the IntelliGraphs sets) and every triple of the graph adds the
    // SAME vector: count the graph's triples per relation in shared memory and push count * vector ONCE per relation
    // present — n_rel reductions per (graph, column chunk) instead of one per triple, all of them onto the same few L2 lines
    __shared__ int cnt[kRelHist];
    for (int r = threadIdx.x; r < n_rel; r += blockDim.x) cnt[r] = 0;
    __syncthreads();
    for (int t = blockIdx.z * blockDim.x + threadIdx.x; t < T; t += gridDim.z * blockDim.x) {
      const int64_t rel = tri[t * 3 + 1];
      if (rel >= 0 && rel < n_rel && !(pad_rid >= 0 && rel == pad_rid)) atomicAdd(&cnt[(int)rel], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d4; c += blockDim.x) {
      float4 v = *reinterpret_cast<const float4*>(dg + (int64_t)b * 3 * d + (int64_t)d + c * 4);
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      for (int r = 0; r < n_rel; ++r) {
        const float k = (float)cnt[r];
        if (k != 0.f) red_add_v4(dR + (int64_t)r * d + c * 4, make_float4(v.x * k, v.y * k, v.z * k, v.w * k));
      }
    }
    return;
  }
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(dg + (int64_t)b * 3 * d + (int64_t)slot * d + c * 4);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    for (int t = blockIdx.z; t < T; t += gridDim.z) {
      const int64_t rel = tri[t * 3 + 1];
      if (pad_rid >= 0 && rel == pad_rid) continue;
      const int64_t idx = tri[t * 3 + slot];
      if (slot != 1 && idx == pad_eid) continue;  // padding_idx row never accumulates
      red_add_v4(table + idx * d + c * 4, v);
    }
  }
}

// grid-stride, 4 elements per thread per trip when dz % 4 == 0 (128-bit accesses), ONE atomic per CTA
template <bool VEC>
__global__ void __launch_bounds__(256) reparam_kl_fwd_kernel(
    const float* __restrict__ heads, int ld_heads, const float* __restrict__ eps, const int32_t* __restrict__ perm,
    int B, int dz, int clamp_logv, float kl_scale, float* __restrict__ z, uint16_t* __restrict__ z_bf16,
    int ld_zb, float* __restrict__ kl_acc) {
  __shared__ float red[33];
  constexpr int W = VEC ? 4 : 1;
  const int64_t n = (int64_t)B * dz / W;
  float term = 0.f;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q * W;
    const int b = (int)(i / dz), j = (int)(i - (int64_t)b * dz);
    float mu[W], lv[W], e[W], zz[W];
    const float* hp = heads + (int64_t)b * ld_heads + j;
    const float* ep = eps + (int64_t)(perm ? perm[b] : b) * dz + j;
    if (VEC) {
      const float4 a = *reinterpret_cast<const float4*>(hp), c = *reinterpret_cast<const float4*>(hp + dz);
      const float4 r = *reinterpret_cast<const float4*>(ep);
      mu[0] = a.x; mu[W > 1 ? 1 : 0] = a.y; mu[W > 2 ? 2 : 0] = a.z; mu[W > 3 ? 3 : 0] = a.w;
      lv[0] = c.x; lv[W > 1 ? 1 : 0] = c.y; lv[W > 2 ? 2 : 0] = c.z; lv[W > 3 ? 3 : 0] = c.w;
      e[0] = r.x; e[W > 1 ? 1 : 0] = r.y; e[W > 2 ? 2 : 0] = r.z; e[W > 3 ? 3 : 0] = r.w;
    } else {
      mu[0] = hp[0]; lv[0] = hp[dz]; e[0] = ep[0];
    }
#pragma unroll
    for (int k = 0; k < W; ++k) {
      if (clamp_logv) lv[k] = fminf(fmaxf(lv[k], -10.f), 10.f);
      zz[k] = fmaf(e[k], expf(0.5f * lv[k]), mu[k]);
      term += -0.5f * (1.f + lv[k] - mu[k] * mu[k] - expf(lv[k]));
    }
    if (VEC) {
      *reinterpret_cast<float4*>(z + i) = make_float4(zz[0], zz[W > 1 ? 1 : 0], zz[W > 2 ? 2 : 0], zz[W > 3 ? 3 : 0]);
      if (z_bf16) {
        uint2 pk;
        pk.x = pack_bf16x2(zz[0], zz[W > 1 ? 1 : 0]);
        pk.y = pack_bf16x2(zz[W > 2 ? 2 : 0], zz[W > 3 ? 3 : 0]);
        *reinterpret_cast<uint2*>(z_bf16 + (int64_t)b * ld_zb + j) = pk;
      }
    } else {
      z[i] = zz[0];
      if (z_bf16) z_bf16[(int64_t)b * ld_zb + j] = f32_to_bf16_bits(zz[0]);
    }
  }
  const float s = block_sum(term, red);
  if (threadIdx.x == 0 && kl_acc) atomicAdd(kl_acc, s * kl_scale);
}

// W = 4: four consecutive latent units per thread with 128-bit accesses (dz, strides % 4 == 0); W = 1: any shape
template <int W>
__global__ void __launch_bounds__(256) reparam_kl_bwd_kernel(
    const float* __restrict__ heads, int ld_heads, const float* __restrict__ eps, const int32_t* __restrict__ perm,
    const float* __restrict__ dz_in, int B, int dz, int clamp_logv, float bk, const float* __restrict__ beta_dev,
    const float* __restrict__ dmu_ext, const float* __restrict__ dlogv_ext, float* __restrict__ dheads,
    uint16_t* __restrict__ dheads_bf16, int ld_dh) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * W;
  if (i >= (int64_t)B * dz) return;
  if (beta_dev) bk *= *beta_dev;      // replayed CUDA graphs: beta lives in device memory (changes every epoch)
  const int b = (int)(i / dz), j = (int)(i - (int64_t)b * dz);
  const int64_t src = (int64_t)(perm ? perm[b] : b) * dz + j;
  float mu[W], raw[W], e[W], dzv[W], xm[W], xl[W];
  const float* hp = heads + (int64_t)b * ld_heads + j;
  if constexpr (W == 4) {
    const float4 a = *reinterpret_cast<const float4*>(hp), c = *reinterpret_cast<const float4*>(hp + dz);
    const float4 r = *reinterpret_cast<const float4*>(eps + src), g = *reinterpret_cast<const float4*>(dz_in + i);
    mu[0] = a.x; mu[W - 3] = a.y; mu[W - 2] = a.z; mu[W - 1] = a.w;
    raw[0] = c.x; raw[W - 3] = c.y; raw[W - 2] = c.z; raw[W - 1] = c.w;
    e[0] = r.x; e[W - 3] = r.y; e[W - 2] = r.z; e[W - 1] = r.w;
    dzv[0] = g.x; dzv[W - 3] = g.y; dzv[W - 2] = g.z; dzv[W - 1] = g.w;
  } else {
    mu[0] = hp[0]; raw[0] = hp[dz]; e[0] = eps[src]; dzv[0] = dz_in[i];
  }
#pragma unroll
  for (int k = 0; k < W; ++k) {
    xm[k] = dmu_ext ? dmu_ext[src + k] : 0.f;      // upstream gradients of the returned (mu, logv) (autograd forward())
    xl[k] = dlogv_ext ? dlogv_ext[src + k] : 0.f;
  }
  float dmu[W], dlv[W];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    float lv = raw[k];
    bool pass = true;
    if (clamp_logv) {
      lv = fminf(fmaxf(raw[k], -10.f), 10.f);
      pass = (raw[k] >= -10.f) && (raw[k] <= 10.f);  // torch.clamp passes the gradient on the closed interval
    }
    dmu[k] = fmaf(bk, mu[k], dzv[k]) + xm[k];
    dlv[k] = 0.5f * dzv[k] * e[k] * expf(0.5f * lv) + 0.5f * bk * (expf(lv) - 1.f) + xl[k];
    if (!pass) dlv[k] = 0.f;
  }
  const int64_t o = (int64_t)b * ld_dh + j;
  if constexpr (W == 4) {
    if (dheads) {
      *reinterpret_cast<float4*>(dheads + o) = make_float4(dmu[0], dmu[W - 3], dmu[W - 2], dmu[W - 1]);
      *reinterpret_cast<float4*>(dheads + o + dz) = make_float4(dlv[0], dlv[W - 3], dlv[W - 2], dlv[W - 1]);
    }
    if (dheads_bf16) {
      *reinterpret_cast<uint2*>(dheads_bf16 + o) = make_uint2(pack_bf16x2(dmu[0], dmu[W - 3]), pack_bf16x2(dmu[W - 2], dmu[W - 1]));
      *reinterpret_cast<uint2*>(dheads_bf16 + o + dz) = make_uint2(pack_bf16x2(dlv[0], dlv[W - 3]), pack_bf16x2(dlv[W - 2], dlv[W - 1]));
    }
  } else {
    if (dheads) { dheads[o] = dmu[0]; dheads[o + dz] = dlv[0]; }
    if (dheads_bf16) { dheads_bf16[o] = f32_to_bf16_bits(dmu[0]); dheads_bf16[o + dz] = f32_to_bf16_bits(dlv[0]); }
  }
}

}  // namespace ark

using namespace ark;

extern "C" int ark_gather_pool_fwd(const int64_t* triples, const int32_t* perm, const float* E, const float* R,
                                   int64_t B, int64_t T, int64_t d, int64_t pad_rid, float* g, uint16_t* g_bf16,
                                   float* inv_cnt, void* stream) {
  ARK_REQUIRE(triples && E && R && inv_cnt && (g || g_bf16), ARK_E_BADARG, "gather_pool_fwd: null pointer");
  ARK_REQUIRE(B > 0 && T > 0 && d > 0, ARK_E_BADARG, "gather_pool_fwd: B,T,d must be positive");
  ARK_REQUIRE(d % 4 == 0, ARK_E_SHAPE, "gather_pool_fwd: d=%lld must be a multiple of 4", (long long)d);
  ARK_REQUIRE(aligned16(E) && aligned16(R) && (!g || aligned16(g)) && (!g_bf16 || aligned16(g_bf16)), ARK_E_ALIGN,
              "gather_pool_fwd: tables/outputs must be 16-byte aligned");
  // column lanes: a warp multiple covering d/4 (at most 128); t-lanes: as many as give ~4 triples per lane
  const int d4 = (int)(d / 4);
  const int cx = d4 >= 128 ? 128 : (d4 + 31) / 32 * 32;
  int ty = (int)((T + 3) / 4);
  if (ty > GP_MAX_THREADS / cx) ty = GP_MAX_THREADS / cx;
  if (ty < 1) ty = 1;
  dim3 grid((unsigned)B, 3);
  gather_pool_fwd_kernel<4><<<grid, cx * ty, (size_t)cx * ty * sizeof(float4), (cudaStream_t)stream>>>(
      triples, perm, E, R, (int)T, (int)d, (long long)pad_rid, cx, g, g_bf16, inv_cnt);
  return launched("gather_pool_fwd");
}

extern "C" int ark_gather_pool_bwd(const float* dg, const int64_t* triples, const int32_t* perm,
                                   const float* inv_cnt, int64_t B, int64_t T, int64_t d, int64_t pad_rid,
                                   int64_t pad_eid, int64_t n_rel, float* dE, float* dR, void* stream) {
  ARK_REQUIRE(dg && triples && inv_cnt && dE && dR, ARK_E_BADARG, "gather_pool_bwd: null pointer");
  ARK_REQUIRE(B > 0 && T > 0 && d > 0, ARK_E_BADARG, "gather_pool_bwd: B,T,d must be positive");
  ARK_REQUIRE(d % 4 == 0, ARK_E_SHAPE, "gather_pool_bwd: d must be a multiple of 4");
  ARK_REQUIRE(aligned16(dg) && aligned16(dE) && aligned16(dR), ARK_E_ALIGN, "gather_pool_bwd: 16-byte alignment");
  const int threads = (int)((d / 4 + 31) / 32 * 32 < 256 ? (d / 4 + 31) / 32 * 32 : 256);
  // split the triples of a graph over Z CTAs until the grid has a few waves (B = 16 graphs alone is 48 CTAs)
  int z = (int)((4 * kNumSMs + 3 * B - 1) / (3 * B));
  if (z > T) z = (int)T;
  if (z > 32) z = 32;
  if (z < 1) z = 1;
  dim3 grid((unsigned)B, 3, (unsigned)z);
  gather_pool_bwd_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(dg, triples, perm, inv_cnt, (int)T, (int)d,
                                                                     (long long)pad_rid, (long long)pad_eid, (int)n_rel,
                                                                     dE, dR);
  return launched("gather_pool_bwd");
}

extern "C" int ark_reparam_kl_fwd(const float* heads, int64_t ld_heads, const float* eps, const int32_t* perm,
                                  int64_t B, int64_t dz, int clamp_logv, float kl_scale, float* z,
                                  uint16_t* z_bf16, int64_t ld_zb, float* kl_acc, void* stream) {
  ARK_REQUIRE(heads && eps && z, ARK_E_BADARG, "reparam_kl_fwd: null pointer");
  ARK_REQUIRE(B > 0 && dz > 0 && ld_heads >= 2 * dz && (!z_bf16 || ld_zb >= dz), ARK_E_BADARG,
              "reparam_kl_fwd: bad sizes");
  const bool vec = dz % 4 == 0 && ld_heads % 4 == 0 && (!z_bf16 || ld_zb % 4 == 0) && aligned16(heads) && aligned16(eps) &&
                   aligned16(z) && (!z_bf16 || (reinterpret_cast<uintptr_t>(z_bf16) & 7) == 0);
  const int64_t work = B * dz / (vec ? 4 : 1);
  int64_t blocks = (work + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (vec)
    reparam_kl_fwd_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        heads, (int)ld_heads, eps, perm, (int)B, (int)dz, clamp_logv, kl_scale, z, z_bf16, (int)ld_zb, kl_acc);
  else
    reparam_kl_fwd_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        heads, (int)ld_heads, eps, perm, (int)B, (int)dz, clamp_logv, kl_scale, z, z_bf16, (int)ld_zb, kl_acc);
  return launched("reparam_kl_fwd");
}

extern "C" int ark_reparam_kl_bwd(const float* heads, int64_t ld_heads, const float* eps, const int32_t* perm,
                                  const float* dz_in, int64_t B, int64_t dz, int clamp_logv, float beta_kl_scale,
                                  const float* beta_dev, const float* dmu_ext, const float* dlogv_ext,
                                  float* dheads, uint16_t* dheads_bf16, int64_t ld_dh, void* stream) {
  ARK_REQUIRE(heads && eps && dz_in && (dheads || dheads_bf16), ARK_E_BADARG, "reparam_kl_bwd: null pointer");
  ARK_REQUIRE(B > 0 && dz > 0 && ld_heads >= 2 * dz && ld_dh >= 2 * dz, ARK_E_BADARG, "reparam_kl_bwd: bad sizes");
  const int64_t n = B * dz;
  const bool vec = dz % 4 == 0 && ld_heads % 4 == 0 && ld_dh % 4 == 0 && aligned16(heads) && aligned16(eps) && aligned16(dz_in) &&
                   (!dheads || aligned16(dheads)) && (!dheads_bf16 || (reinterpret_cast<uintptr_t>(dheads_bf16) & 7) == 0) &&
                   (!dmu_ext || aligned16(dmu_ext)) && (!dlogv_ext || aligned16(dlogv_ext));
  if (vec)
    reparam_kl_bwd_kernel<4><<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        heads, (int)ld_heads, eps, perm, dz_in, (int)B, (int)dz, clamp_logv, beta_kl_scale, beta_dev, dmu_ext, dlogv_ext,
        dheads, dheads_bf16, (int)ld_dh);
  else
    reparam_kl_bwd_kernel<1><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        heads, (int)ld_heads, eps, perm, dz_in, (int)B, (int)dz, clamp_logv, beta_kl_scale, beta_dev, dmu_ext, dlogv_ext,
        dheads, dheads_bf16, (int)ld_dh);
  return launched("reparam_kl_bwd");
}
