// t-SAIL (Transformer encoder / decoder, kgvae/model/models.py:66-114) building blocks over RAGGED, PAD-free,
// graph-major packed rows (graph b owns rows cu[b] .. cu[b+1]):
//
//   K8  attention as batched ragged GEMMs + row softmax (nn.MultiheadAttention inside
//       nn.TransformerEncoderLayer / nn.TransformerDecoderLayer: scores / sqrt(hd), key-padding mask == the
//       PAD rows that do not exist here, causal triu mask, softmax, attention dropout, P.V), forward + backward;
//   K9  residual add + dropout + LayerNorm (post-LN, eps 1e-5), forward + backward;
//   plus the glue: triple embedding rows (no pooling), token+position embedding in fp32, masked mean-pool as a
//   segment mean, the collapsed cross-attention (memory = z_proj(z) repeated L times, models.py:111).
//
// This first version computes the [n_b x n_b] score blocks with an fp32-FMA tiled kernel (generic strides so
// the six products of attention forward/backward share one kernel); the dense projections around it run on the
// tcgen05 GEMM.  It is the parity vehicle for SURVEY.md §8 rows a11/a12, not yet a roofline kernel.
#include "common.cuh"
#include "philox.cuh"

namespace ark {

// ------------------------------------------------------------------------------------------------
// batched ragged GEMM:  C_p[M x N] = alpha * A_p[M x K] . B_p[K x N],  p = (graph b, head h)
// Operand kinds: TOK = rows of a token matrix [N_tok, ld] restricted to graph b and head h's hd columns;
//                SQ  = the [n_b x n_b] block of (b, h) inside a packed buffer (offset sq_off[b]*H + h*n_b^2).
// ------------------------------------------------------------------------------------------------
constexpr int BG_K = 16;   // k-step; the output tile BG_T x BG_T (64 / 32 / 16) follows the longest graph of the batch

struct BgOperand {
  const void* ptr;   // bf16 or f32
  int64_t ld;        // TOK: row stride (elements); SQ: unused
  int col0;          // TOK: first column of head 0
  int kind;          // 0 TOK, 1 SQ
  int trans;         // 0: (row, col) as stored; 1: transposed view
  int is_f32;
};
struct BgParams {
  BgOperand A, B;
  void* C;
  int64_t c_ld;
  int c_col0, c_kind, c_f32;
  const int32_t* cu;       // [nb+1]
  const int64_t* sq_off;   // [nb+1] prefix sums of n_b^2
  int H, hd;
  int mode;                // 0: M=N=n,K=hd (scores)   1: M=n,N=hd,K=n (apply)
  int causal;              // scores: skip tiles above the diagonal; apply: restrict the k range
  int a_lower;             // apply with A = SQ^T (dV, dK): A(j,i) != 0 only for i >= j
  float alpha;
};

__device__ __forceinline__ float bg_load(const void* p, int64_t off, int is_f32) {
  return is_f32 ? reinterpret_cast<const float*>(p)[off] : bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(p)[off]);
}

template <int BG_T>
__global__ void __launch_bounds__(256) bgemm_kernel(const BgParams p) {
  constexpr int R = BG_T / 16;   // outputs per thread per dimension
  __shared__ float As[BG_K][BG_T + 1];
  __shared__ float Bs[BG_K][BG_T + 1];
  const int b = blockIdx.z / p.H, h = blockIdx.z % p.H;
  const int r0 = p.cu[b], n = p.cu[b + 1] - r0;
  const int M = n, N = p.mode == 0 ? n : p.hd, K = p.mode == 0 ? p.hd : n;
  const int m0 = blockIdx.y * BG_T, n0 = blockIdx.x * BG_T;
  if (m0 >= M || n0 >= N) return;
  if (p.mode == 0 && p.causal && n0 > m0 + BG_T - 1) return;   // whole tile above the diagonal: never read
  const int64_t sq_base = p.sq_off[b] * p.H + (int64_t)h * n * n;
  // element (i,k) of A and (k,j) of B as base + i*rs + k*cs
  int64_t a_base, a_rs, a_cs, b_base, b_rs, b_cs;
  if (p.A.kind == 0) { a_base = (int64_t)r0 * p.A.ld + p.A.col0 + h * p.hd; a_rs = p.A.ld; a_cs = 1; }
  else { a_base = sq_base; a_rs = n; a_cs = 1; }
  if (p.A.trans) { const int64_t t = a_rs; a_rs = a_cs; a_cs = t; }
  if (p.B.kind == 0) { b_base = (int64_t)r0 * p.B.ld + p.B.col0 + h * p.hd; b_rs = p.B.ld; b_cs = 1; }
  else { b_base = sq_base; b_rs = n; b_cs = 1; }
  if (p.B.trans) { const int64_t t = b_rs; b_rs = b_cs; b_cs = t; }
  int k_lo = 0, k_hi = K;
  if (p.mode == 1 && p.causal) {
    if (p.a_lower) k_lo = (m0 / BG_K) * BG_K;          // A(j,i) = P[i,j]: zero for i < j
    else k_hi = min(K, m0 + BG_T);                     // A(i,j) = P[i,j]: zero for j > i
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;
  // loader mapping: the thread's fastest index follows the operand's unit-stride dimension
  const bool a_kfast = (a_cs == 1), b_nfast = (b_cs == 1);
  for (int k0 = k_lo; k0 < k_hi; k0 += BG_K) {
#pragma unroll
    for (int e = 0; e < R; ++e) {
      const int idx = threadIdx.x + 256 * e;              // 0 .. BG_T*16-1
      int mi, kk;
      if (a_kfast) { mi = idx >> 4; kk = idx & 15; } else { kk = idx / BG_T; mi = idx % BG_T; }
      const int gm = m0 + mi, gk = k0 + kk;
      As[kk][mi] = (gm < M && gk < k_hi) ? bg_load(p.A.ptr, a_base + gm * a_rs + gk * a_cs, p.A.is_f32) : 0.f;
      int ni, kb;
      if (b_nfast) { kb = idx / BG_T; ni = idx % BG_T; } else { ni = idx >> 4; kb = idx & 15; }
      const int gn = n0 + ni, gk2 = k0 + kb;
      Bs[kb][ni] = (gn < N && gk2 < k_hi) ? bg_load(p.B.ptr, b_base + gk2 * b_rs + gn * b_cs, p.B.is_f32) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BG_K; ++k) {
      float a[R], bb[R];
#pragma unroll
      for (int i = 0; i < R; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < R; ++j) bb[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  int64_t c_base, c_rs;
  if (p.c_kind == 0) { c_base = (int64_t)r0 * p.c_ld + p.c_col0 + h * p.hd; c_rs = p.c_ld; }
  else { c_base = sq_base; c_rs = n; }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int nn = n0 + tx + 16 * j;
      if (nn >= N) continue;
      const float v = acc[i][j] * p.alpha;
      const int64_t o = c_base + m * c_rs + nn;
      if (p.c_f32) reinterpret_cast<float*>(p.C)[o] = v;
      else reinterpret_cast<uint16_t*>(p.C)[o] = f32_to_bf16_bits(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// row softmax over the score blocks.  One warp per (token row, head).
// fwd: P = softmax(S[i, :lim]) (lim = i+1 causal, n otherwise), zeros beyond; optional dropout -> Pd.
// bwd: dS = P * (dP - sum_j P dP), dP = dPd * keep/(1-p) when dropout was applied (keep <=> Pd != 0 or P == 0).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(
    const float* __restrict__ S, const int32_t* __restrict__ cu, const int64_t* __restrict__ sq_off,
    const int32_t* __restrict__ tok_graph, int64_t n_tok, int H, int causal, float p_drop, uint64_t seed,
    uint64_t offset, const uint64_t* __restrict__ offset_dev, uint16_t* __restrict__ P, uint16_t* __restrict__ Pd) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_tok * H) return;
  const int lane = threadIdx.x & 31;
  const int64_t tok = w / H;
  const int h = (int)(w % H);
  const int b = tok_graph[tok];
  const int r0 = cu[b], n = cu[b + 1] - r0, i = (int)(tok - r0);
  const int64_t base = sq_off[b] * H + ((int64_t)h * n + i) * n;
  const int lim = causal ? i + 1 : n;
  float mx = -INFINITY;
  for (int j = lane; j < lim; j += 32) mx = fmaxf(mx, S[base + j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < lim; j += 32) sum += __expf(S[base + j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  const float scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint64_t ctr0 = offset + (offset_dev ? *offset_dev : 0ull);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int j = lane; j < n; j += 32) {
    const float pj = j < lim ? __expf(S[base + j] - mx) * inv : 0.f;
    P[base + j] = f32_to_bf16_bits(pj);
    if (Pd) {
      const uint64_t c = ctr0 + (uint64_t)((base + j) >> 2);
      const uint4 rn = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
      const uint32_t rr[4] = {rn.x, rn.y, rn.z, rn.w};
      const bool keep = (float)(rr[(base + j) & 3] >> 8) * (1.f / 16777216.f) >= p_drop;
      Pd[base + j] = keep ? f32_to_bf16_bits(pj * scale) : (uint16_t)0;
    }
  }
}

// fp32 inference path: S <- softmax(S) in place (zeros beyond the causal limit), no dropout, nothing saved
__global__ void __launch_bounds__(256) attn_softmax_inplace_kernel(
    float* __restrict__ S, const int32_t* __restrict__ cu, const int64_t* __restrict__ sq_off,
    const int32_t* __restrict__ tok_graph, int64_t n_tok, int H, int causal) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_tok * H) return;
  const int lane = threadIdx.x & 31;
  const int64_t tok = w / H;
  const int h = (int)(w % H);
  const int b = tok_graph[tok];
  const int r0 = cu[b], n = cu[b + 1] - r0, i = (int)(tok - r0);
  const int64_t base = sq_off[b] * H + ((int64_t)h * n + i) * n;
  const int lim = causal ? i + 1 : n;
  float mx = -INFINITY;
  for (int j = lane; j < lim; j += 32) mx = fmaxf(mx, S[base + j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < lim; j += 32) sum += expf(S[base + j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < n; j += 32) S[base + j] = j < lim ? expf(S[base + j] - mx) * inv : 0.f;
}

__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(
    const uint16_t* __restrict__ P, const uint16_t* __restrict__ Pd, const float* __restrict__ dP,
    const int32_t* __restrict__ cu, const int64_t* __restrict__ sq_off, const int32_t* __restrict__ tok_graph,
    int64_t n_tok, int H, int causal, float p_drop, float alpha, uint16_t* __restrict__ dS) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_tok * H) return;
  const int lane = threadIdx.x & 31;
  const int64_t tok = w / H;
  const int h = (int)(w % H);
  const int b = tok_graph[tok];
  const int r0 = cu[b], n = cu[b + 1] - r0, i = (int)(tok - r0);
  const int64_t base = sq_off[b] * H + ((int64_t)h * n + i) * n;
  const int lim = causal ? i + 1 : n;
  const float scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float dot = 0.f;
  for (int j = lane; j < lim; j += 32) {
    const float pj = bf16_bits_to_f32(P[base + j]);
    const bool keep = !Pd || Pd[base + j] != 0;
    dot += keep ? pj * dP[base + j] * scale : 0.f;
  }
  dot = warp_sum(dot);
  for (int j = lane; j < n; j += 32) {
    float v = 0.f;
    if (j < lim) {
      const float pj = bf16_bits_to_f32(P[base + j]);
      const bool keep = !Pd || Pd[base + j] != 0;
      v = pj * ((keep ? dP[base + j] * scale : 0.f) - dot) * alpha;
    }
    dS[base + j] = f32_to_bf16_bits(v);
  }
}

// ------------------------------------------------------------------------------------------------
// K9: s = res + dropout(branch[rows[r] or r]);  y = LN(s) * gamma + beta.   One CTA (128 threads) per row.
// `branch` is overwritten with s (the backward needs it); mask u8 saved when p_drop > 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) add_layernorm_fwd_kernel(
    float* __restrict__ branch, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, int D, float eps, float p_drop, uint64_t seed, uint64_t offset,
    const uint64_t* __restrict__ offset_dev, uint8_t* __restrict__ mask, float* __restrict__ y,
    uint16_t* __restrict__ y_bf16, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ float red[33];
  const int64_t row = blockIdx.x;
  float* br = branch + row * D;
  const float* rs = res + row * D;
  const float scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const uint64_t ctr0 = offset + (offset_dev ? *offset_dev : 0ull);
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  float sum = 0.f;
  for (int c = threadIdx.x * 4; c < D; c += 512) {
    float4 v = *reinterpret_cast<const float4*>(br + c);
    const float4 r = *reinterpret_cast<const float4*>(rs + c);
    if (p_drop > 0.f) {
      const uint64_t ct = ctr0 + (uint64_t)((row * D + c) >> 2);
      const uint4 rn = philox4x32_10(make_uint4((uint32_t)ct, (uint32_t)(ct >> 32), 0u, 0u), key);
      const bool k0 = (float)(rn.x >> 8) * (1.f / 16777216.f) >= p_drop, k1 = (float)(rn.y >> 8) * (1.f / 16777216.f) >= p_drop;
      const bool k2 = (float)(rn.z >> 8) * (1.f / 16777216.f) >= p_drop, k3 = (float)(rn.w >> 8) * (1.f / 16777216.f) >= p_drop;
      v.x = k0 ? v.x * scale : 0.f; v.y = k1 ? v.y * scale : 0.f; v.z = k2 ? v.z * scale : 0.f; v.w = k3 ? v.w * scale : 0.f;
      *reinterpret_cast<uint32_t*>(mask + row * D + c) = (k0 ? 1u : 0u) | (k1 ? 0x100u : 0u) | (k2 ? 0x10000u : 0u) | (k3 ? 0x1000000u : 0u);
    }
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    *reinterpret_cast<float4*>(br + c) = v;
    sum += v.x + v.y + v.z + v.w;
  }
  const float mean = block_sum(sum, red) / (float)D;
  float var = 0.f;
  for (int c = threadIdx.x * 4; c < D; c += 512) {
    const float4 v = *reinterpret_cast<const float4*>(br + c);
    const float a = v.x - mean, b2 = v.y - mean, c2 = v.z - mean, d2 = v.w - mean;
    var += a * a + b2 * b2 + c2 * c2 + d2 * d2;
  }
  const float rstd = rsqrtf(block_sum(var, red) / (float)D + eps);
  if (threadIdx.x == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  for (int c = threadIdx.x * 4; c < D; c += 512) {
    const float4 v = *reinterpret_cast<const float4*>(br + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + bt.x; o.y = (v.y - mean) * rstd * g.y + bt.y;
    o.z = (v.z - mean) * rstd * g.z + bt.z; o.w = (v.w - mean) * rstd * g.w + bt.w;
    *reinterpret_cast<float4*>(y + row * D + c) = o;
    if (y_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(y_bf16 + row * D + c) = pk;
    }
  }
}

// ds = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.   d_res = ds (f32); d_branch = ds * keep/(1-p)
// written as bf16 (operand of the branch's backward GEMMs) and/or f32.  dgamma/dbeta: per-CTA rows, atomics.
__global__ void __launch_bounds__(128) add_layernorm_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ s, const float* __restrict__ mean_in,
    const float* __restrict__ rstd_in, const float* __restrict__ gamma, int D, float p_drop,
    const uint8_t* __restrict__ mask, float* __restrict__ d_res, float* __restrict__ d_branch_f32,
    uint16_t* __restrict__ d_branch_bf16, int64_t n_rows, int rows_per_cta, float* __restrict__ dgamma,
    float* __restrict__ dbeta) {
  __shared__ float red[33];
  const float scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row_end = min(n_rows, row_begin + rows_per_cta);
  // per-thread dgamma/dbeta partials for the columns this thread owns (D <= 512*MAXC*... handled by loop below)
  constexpr int MAXC = 8;   // D <= 4096
  float4 dg_acc[MAXC], db_acc[MAXC];
#pragma unroll
  for (int u = 0; u < MAXC; ++u) dg_acc[u] = db_acc[u] = make_float4(0, 0, 0, 0);
  for (int64_t row = row_begin; row < row_end; ++row) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int c = threadIdx.x * 4 + u * 512;
      if (c < D) {
        const float4 d4 = *reinterpret_cast<const float4*>(dy + row * D + c);
        const float4 v = *reinterpret_cast<const float4*>(s + row * D + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        const float x0 = (v.x - mean) * rstd, x1 = (v.y - mean) * rstd, x2 = (v.z - mean) * rstd, x3 = (v.w - mean) * rstd;
        sg += d4.x * g.x + d4.y * g.y + d4.z * g.z + d4.w * g.w;
        sgx += d4.x * g.x * x0 + d4.y * g.y * x1 + d4.z * g.z * x2 + d4.w * g.w * x3;
        dg_acc[u].x += d4.x * x0; dg_acc[u].y += d4.y * x1; dg_acc[u].z += d4.z * x2; dg_acc[u].w += d4.w * x3;
        db_acc[u].x += d4.x; db_acc[u].y += d4.y; db_acc[u].z += d4.z; db_acc[u].w += d4.w;
      }
    }
    const float mg = block_sum(sg, red) / (float)D;
    const float mgx = block_sum(sgx, red) / (float)D;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int c = threadIdx.x * 4 + u * 512;
      if (c < D) {
        const float4 d4 = *reinterpret_cast<const float4*>(dy + row * D + c);
        const float4 v = *reinterpret_cast<const float4*>(s + row * D + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        float4 o;
        o.x = rstd * (d4.x * g.x - mg - (v.x - mean) * rstd * mgx);
        o.y = rstd * (d4.y * g.y - mg - (v.y - mean) * rstd * mgx);
        o.z = rstd * (d4.z * g.z - mg - (v.z - mean) * rstd * mgx);
        o.w = rstd * (d4.w * g.w - mg - (v.w - mean) * rstd * mgx);
        *reinterpret_cast<float4*>(d_res + row * D + c) = o;
        float4 bq = o;
        if (p_drop > 0.f) {
          const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + row * D + c);
          bq.x = (m & 0xffu) ? o.x * scale : 0.f; bq.y = (m & 0xff00u) ? o.y * scale : 0.f;
          bq.z = (m & 0xff0000u) ? o.z * scale : 0.f; bq.w = (m & 0xff000000u) ? o.w * scale : 0.f;
        }
        if (d_branch_f32) *reinterpret_cast<float4*>(d_branch_f32 + row * D + c) = bq;
        if (d_branch_bf16) {
          uint2 pk;
          pk.x = pack_bf16x2(bq.x, bq.y);
          pk.y = pack_bf16x2(bq.z, bq.w);
          *reinterpret_cast<uint2*>(d_branch_bf16 + row * D + c) = pk;
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int c = threadIdx.x * 4 + u * 512;
    if (c < D) {
      red_add_v4(dgamma + c, dg_acc[u]);
      red_add_v4(dbeta + c, db_acc[u]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// glue kernels
// ------------------------------------------------------------------------------------------------
// X[r, slot*d + c] = (slot == 1 ? R : E)[idx[r, slot], c]   (models.py:80-83, real triples only)
__global__ void __launch_bounds__(256) triple_embed_fwd_kernel(const int32_t* __restrict__ idx, const float* __restrict__ E,
                                                               const float* __restrict__ R, int64_t n, int d,
                                                               float* __restrict__ X, uint16_t* __restrict__ Xb) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n * 3) return;
  const int lane = threadIdx.x & 31;
  const int64_t r = w / 3;
  const int slot = (int)(w % 3);
  const float* src = (slot == 1 ? R : E) + (int64_t)idx[w] * d;
  const int64_t o = r * 3 * d + (int64_t)slot * d;
  for (int c = lane * 4; c < d; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
    *reinterpret_cast<float4*>(X + o + c) = v;
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(Xb + o + c) = pk;
  }
}
__global__ void __launch_bounds__(256) triple_embed_bwd_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dX,
                                                               int64_t n, int d, float* __restrict__ dE,
                                                               float* __restrict__ dR) {
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n * 3) return;
  const int lane = threadIdx.x & 31;
  const int64_t r = w / 3;
  const int slot = (int)(w % 3);
  float* dst = (slot == 1 ? dR : dE) + (int64_t)idx[w] * d;
  const int64_t o = r * 3 * d + (int64_t)slot * d;
  for (int c = lane * 4; c < d; c += 128) red_add_v4(dst + c, *reinterpret_cast<const float4*>(dX + o + c));
}

// X[r] = W[tok[r]] + P[pos[r]] in fp32 (+ bf16 copy)   (models.py:109-110)
__global__ void __launch_bounds__(256) embed_sum_fwd_kernel(const float* __restrict__ W, const float* __restrict__ Pm,
                                                            const int32_t* __restrict__ tok, const int32_t* __restrict__ pos,
                                                            int64_t n, int d, float* __restrict__ X, uint16_t* __restrict__ Xb) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  const float* a = W + (int64_t)tok[r] * d;
  const float* b = Pm + (int64_t)pos[r] * d;
  for (int c = lane * 4; c < d; c += 128) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(a + c)), v = __ldg(reinterpret_cast<const float4*>(b + c));
    const float4 o = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    *reinterpret_cast<float4*>(X + r * d + c) = o;
    uint2 pk;
    pk.x = pack_bf16x2(o.x, o.y);
    pk.y = pack_bf16x2(o.z, o.w);
    *reinterpret_cast<uint2*>(Xb + r * d + c) = pk;
  }
}

// segment reduce over the rows of each graph: out[b] = scale_b * sum_{r in graph b} w_r * X[r]   (scale_b = 1/n_b for
// the masked mean-pool of models.py:88-89; w_r = optional per-(row, head) weights for the cross-attention backward)
__global__ void __launch_bounds__(256) seg_reduce_kernel(const float* __restrict__ X, const int32_t* __restrict__ cu, int D,
                                                         int mean, const float* __restrict__ wts, int H,
                                                         float* __restrict__ out, uint16_t* __restrict__ out_bf16) {
  const int b = blockIdx.x;
  const int r0 = cu[b], r1 = cu[b + 1];
  const float sc = mean ? 1.f / (float)max(r1 - r0, 1) : 1.f;
  const int hd = wts ? D / H : D;
  for (int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4; c < D; c += gridDim.y * blockDim.x * 4) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int r = r0; r < r1; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(X + (int64_t)r * D + c);
      const float w = wts ? wts[(int64_t)r * H + c / hd] : 1.f;
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
    if (out) *reinterpret_cast<float4*>(out + (int64_t)b * D + c) = acc;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(acc.x, acc.y);
      pk.y = pack_bf16x2(acc.z, acc.w);
      *reinterpret_cast<uint2*>(out_bf16 + (int64_t)b * D + c) = pk;
    }
  }
}

// out[r] = scale_b * w_r * src[graph(r)]   (broadcast of a per-graph row to its token rows; mean-pool backward with
// scale_b = 1/n_b; collapsed cross-attention forward with w = attention-dropout weights, models.py:111)
__global__ void __launch_bounds__(256) seg_broadcast_kernel(const float* __restrict__ src, const int32_t* __restrict__ cu,
                                                            const int32_t* __restrict__ tok_graph, int64_t n, int D,
                                                            int mean, const float* __restrict__ wts, int H,
                                                            float* __restrict__ out, uint16_t* __restrict__ out_bf16) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  const int b = tok_graph[r];
  const float sc = mean ? 1.f / (float)max(cu[b + 1] - cu[b], 1) : 1.f;
  const int hd = wts ? D / H : D;
  for (int c = lane * 4; c < D; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(src + (int64_t)b * D + c);
    const float w = sc * (wts ? wts[r * H + c / hd] : 1.f);
    v.x *= w; v.y *= w; v.z *= w; v.w *= w;
    if (out) *reinterpret_cast<float4*>(out + r * D + c) = v;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(v.x, v.y);
      pk.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(out_bf16 + r * D + c) = pk;
    }
  }
}

// attention-dropout weights of the collapsed cross-attention: every query attends uniformly to n_keys identical
// memory rows, so dropout(p) on the weights turns the output into v * Binomial(n_keys, 1-p) / ((1-p) n_keys).
__global__ void __launch_bounds__(256) xattn_weights_kernel(int64_t n_items, int n_keys, float p_drop, uint64_t seed,
                                                            uint64_t offset, const uint64_t* __restrict__ offset_dev,
                                                            float* __restrict__ wts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const uint64_t per = (uint64_t)(n_keys + 3) / 4;
  const uint64_t ctr0 = offset + (offset_dev ? *offset_dev : 0ull) + (uint64_t)i * per;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  int kept = 0;
  for (int j = 0; j < n_keys; j += 4) {
    const uint64_t c = ctr0 + (uint64_t)(j >> 2);
    const uint4 rn = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
    const uint32_t rr[4] = {rn.x, rn.y, rn.z, rn.w};
    for (int k = 0; k < 4 && j + k < n_keys; ++k) kept += ((float)(rr[k] >> 8) * (1.f / 16777216.f) >= p_drop) ? 1 : 0;
  }
  wts[i] = (float)kept / ((1.f - p_drop) * (float)n_keys);
}

// dpre = (out > 0) ? d * scale : 0  — ReLU backward; with the FFN dropout applied IN PLACE on `out`, out > 0 iff the
// unit was active AND kept, so the same test folds the dropout backward (scale = 1/(1-p)).
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ d, const uint16_t* __restrict__ out,
                                                       int64_t n, float scale, uint16_t* __restrict__ dpre) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 g = *reinterpret_cast<const float4*>(d + i);
  const uint2 o = *reinterpret_cast<const uint2*>(out + i);
  const float2 a = unpack_bf16x2(o.x), b = unpack_bf16x2(o.y);
  uint2 pk;
  pk.x = pack_bf16x2(a.x > 0.f ? g.x * scale : 0.f, a.y > 0.f ? g.y * scale : 0.f);
  pk.y = pack_bf16x2(b.x > 0.f ? g.z * scale : 0.f, b.y > 0.f ? g.w * scale : 0.f);
  *reinterpret_cast<uint2*>(dpre + i) = pk;
}

}  // namespace ark

using namespace ark;

static int fill_operand(BgOperand* o, const void* ptr, int64_t ld, int64_t col0, int kind, int trans, int is_f32) {
  o->ptr = ptr; o->ld = ld; o->col0 = (int)col0; o->kind = kind; o->trans = trans; o->is_f32 = is_f32;
  return 0;
}

extern "C" int ark_attn_bgemm(const void* A, int a_kind, int a_trans, int a_f32, int64_t a_ld, int64_t a_col0,
                              const void* B, int b_kind, int b_trans, int b_f32, int64_t b_ld, int64_t b_col0,
                              void* C, int c_kind, int c_f32, int64_t c_ld, int64_t c_col0, const int32_t* cu,
                              const int64_t* sq_off, int64_t n_graphs, int64_t n_max, int64_t H, int64_t hd, int mode,
                              int causal, float alpha, void* stream) {
  ARK_REQUIRE(A && B && C && cu && sq_off, ARK_E_BADARG, "attn_bgemm: null pointer");
  ARK_REQUIRE(n_graphs > 0 && n_max > 0 && H > 0 && hd > 0 && (mode == 0 || mode == 1), ARK_E_BADARG, "attn_bgemm: bad sizes");
  ARK_REQUIRE(n_graphs * H <= 65535, ARK_E_SHAPE, "attn_bgemm: graphs x heads = %lld exceeds grid.z", (long long)(n_graphs * H));
  BgParams p;
  fill_operand(&p.A, A, a_ld, a_col0, a_kind, a_trans, a_f32);
  fill_operand(&p.B, B, b_ld, b_col0, b_kind, b_trans, b_f32);
  p.C = C; p.c_ld = c_ld; p.c_col0 = (int)c_col0; p.c_kind = c_kind; p.c_f32 = c_f32; p.cu = cu; p.sq_off = sq_off;
  p.H = (int)H; p.hd = (int)hd; p.mode = mode; p.causal = causal; p.a_lower = (mode == 1 && a_trans) ? 1 : 0; p.alpha = alpha;
  const int64_t Mx = n_max, Nx = mode == 0 ? n_max : hd;
  const int tt = n_max <= 16 ? 16 : (n_max <= 32 ? 32 : 64);     // short graphs (syn-*: 3..16 rows): small tiles
  dim3 grid((unsigned)((Nx + tt - 1) / tt), (unsigned)((Mx + tt - 1) / tt), (unsigned)(n_graphs * H));
  if (tt == 16) bgemm_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else if (tt == 32) bgemm_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  else bgemm_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return launched("attn_bgemm");
}

extern "C" int ark_attn_softmax_fwd(const float* S, const int32_t* cu, const int64_t* sq_off, const int32_t* tok_graph,
                                    int64_t n_tok, int64_t H, int causal, float p_drop, uint64_t seed, uint64_t offset,
                                    const uint64_t* offset_dev, uint16_t* P, uint16_t* P_drop, void* stream) {
  ARK_REQUIRE(S && cu && sq_off && tok_graph && P, ARK_E_BADARG, "attn_softmax_fwd: null pointer");
  ARK_REQUIRE((p_drop > 0.f) == (P_drop != nullptr), ARK_E_BADARG, "attn_softmax_fwd: P_drop iff p_drop > 0");
  if (n_tok == 0) return 0;
  attn_softmax_fwd_kernel<<<(unsigned)((n_tok * H + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      S, cu, sq_off, tok_graph, n_tok, (int)H, causal, p_drop, seed, offset, offset_dev, P, P_drop);
  return launched("attn_softmax_fwd");
}

extern "C" int ark_attn_softmax_inplace(float* S, const int32_t* cu, const int64_t* sq_off, const int32_t* tok_graph,
                                        int64_t n_tok, int64_t H, int causal, void* stream) {
  ARK_REQUIRE(S && cu && sq_off && tok_graph, ARK_E_BADARG, "attn_softmax_inplace: null pointer");
  if (n_tok == 0) return 0;
  attn_softmax_inplace_kernel<<<(unsigned)((n_tok * H + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, cu, sq_off, tok_graph,
                                                                                                 n_tok, (int)H, causal);
  return launched("attn_softmax_inplace");
}

extern "C" int ark_attn_softmax_bwd(const uint16_t* P, const uint16_t* P_drop, const float* dP, const int32_t* cu,
                                    const int64_t* sq_off, const int32_t* tok_graph, int64_t n_tok, int64_t H, int causal,
                                    float p_drop, float alpha, uint16_t* dS, void* stream) {
  ARK_REQUIRE(P && dP && cu && sq_off && tok_graph && dS, ARK_E_BADARG, "attn_softmax_bwd: null pointer");
  if (n_tok == 0) return 0;
  attn_softmax_bwd_kernel<<<(unsigned)((n_tok * H + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      P, P_drop, dP, cu, sq_off, tok_graph, n_tok, (int)H, causal, p_drop, alpha, dS);
  return launched("attn_softmax_bwd");
}

extern "C" int ark_add_layernorm_fwd(float* branch, const float* res, const float* gamma, const float* beta, int64_t n,
                                     int64_t D, float eps, float p_drop, uint64_t seed, uint64_t offset,
                                     const uint64_t* offset_dev, uint8_t* mask, float* y, uint16_t* y_bf16, float* mean,
                                     float* rstd, void* stream) {
  ARK_REQUIRE(branch && res && gamma && beta && y && mean && rstd, ARK_E_BADARG, "add_layernorm_fwd: null pointer");
  ARK_REQUIRE(D > 0 && D % 4 == 0 && D <= 4096, ARK_E_SHAPE, "add_layernorm_fwd: D must be a multiple of 4, <= 4096");
  ARK_REQUIRE(p_drop == 0.f || mask, ARK_E_BADARG, "add_layernorm_fwd: dropout needs a mask buffer");
  if (n == 0) return 0;
  add_layernorm_fwd_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(branch, res, gamma, beta, (int)D, eps, p_drop, seed,
                                                                           offset, offset_dev, mask, y, y_bf16, mean, rstd);
  return launched("add_layernorm_fwd");
}

extern "C" int ark_add_layernorm_bwd(const float* dy, const float* s, const float* mean, const float* rstd,
                                     const float* gamma, int64_t n, int64_t D, float p_drop, const uint8_t* mask,
                                     float* d_res, float* d_branch_f32, uint16_t* d_branch_bf16, float* dgamma,
                                     float* dbeta, void* stream) {
  ARK_REQUIRE(dy && s && mean && rstd && gamma && d_res && dgamma && dbeta, ARK_E_BADARG, "add_layernorm_bwd: null pointer");
  ARK_REQUIRE(D > 0 && D % 4 == 0 && D <= 4096, ARK_E_SHAPE, "add_layernorm_bwd: D must be a multiple of 4, <= 4096");
  ARK_REQUIRE(p_drop == 0.f || mask, ARK_E_BADARG, "add_layernorm_bwd: dropout needs the forward mask");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dgamma, 0, sizeof(float) * D, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * D, st);
  if (e != cudaSuccess) return fail((int)e, "add_layernorm_bwd: memset: %s", cudaGetErrorString(e));
  if (n == 0) return 0;
  const int ctas = (int)min((int64_t)(4 * kNumSMs), n);
  const int rpc = (int)((n + ctas - 1) / ctas);
  add_layernorm_bwd_kernel<<<(unsigned)((n + rpc - 1) / rpc), 128, 0, st>>>(dy, s, mean, rstd, gamma, (int)D, p_drop, mask,
                                                                            d_res, d_branch_f32, d_branch_bf16, n, rpc,
                                                                            dgamma, dbeta);
  return launched("add_layernorm_bwd");
}

extern "C" int ark_triple_embed_fwd(const int32_t* idx, const float* E, const float* R, int64_t n, int64_t d, float* X,
                                    uint16_t* X_bf16, void* stream) {
  ARK_REQUIRE(idx && E && R && X && X_bf16, ARK_E_BADARG, "triple_embed_fwd: null pointer");
  ARK_REQUIRE(d > 0 && d % 4 == 0, ARK_E_SHAPE, "triple_embed_fwd: d must be a multiple of 4");
  if (n == 0) return 0;
  triple_embed_fwd_kernel<<<(unsigned)((n * 3 + 7) / 8), 256, 0, (cudaStream_t)stream>>>(idx, E, R, n, (int)d, X, X_bf16);
  return launched("triple_embed_fwd");
}

extern "C" int ark_triple_embed_bwd(const int32_t* idx, const float* dX, int64_t n, int64_t d, float* dE, float* dR,
                                    void* stream) {
  ARK_REQUIRE(idx && dX && dE && dR, ARK_E_BADARG, "triple_embed_bwd: null pointer");
  ARK_REQUIRE(d > 0 && d % 4 == 0, ARK_E_SHAPE, "triple_embed_bwd: d must be a multiple of 4");
  if (n == 0) return 0;
  triple_embed_bwd_kernel<<<(unsigned)((n * 3 + 7) / 8), 256, 0, (cudaStream_t)stream>>>(idx, dX, n, (int)d, dE, dR);
  return launched("triple_embed_bwd");
}

extern "C" int ark_embed_sum_fwd(const float* W, const float* P, const int32_t* tok, const int32_t* pos, int64_t n,
                                 int64_t d, float* X, uint16_t* X_bf16, void* stream) {
  ARK_REQUIRE(W && P && tok && pos && X && X_bf16, ARK_E_BADARG, "embed_sum_fwd: null pointer");
  ARK_REQUIRE(d > 0 && d % 4 == 0, ARK_E_SHAPE, "embed_sum_fwd: d must be a multiple of 4");
  if (n == 0) return 0;
  embed_sum_fwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(W, P, tok, pos, n, (int)d, X, X_bf16);
  return launched("embed_sum_fwd");
}

extern "C" int ark_seg_reduce(const float* X, const int32_t* cu, int64_t n_graphs, int64_t D, int mean, const float* wts,
                              int64_t H, float* out, uint16_t* out_bf16, void* stream) {
  ARK_REQUIRE(X && cu && (out || out_bf16), ARK_E_BADARG, "seg_reduce: null pointer");
  ARK_REQUIRE(D > 0 && D % 4 == 0 && (!wts || (H > 0 && D % H == 0 && (D / H) % 4 == 0)), ARK_E_SHAPE, "seg_reduce: bad D / H");
  if (n_graphs == 0) return 0;
  dim3 grid((unsigned)n_graphs, (unsigned)((D / 4 + 255) / 256));
  seg_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, cu, (int)D, mean, wts, (int)(H > 0 ? H : 1), out, out_bf16);
  return launched("seg_reduce");
}

extern "C" int ark_seg_broadcast(const float* src, const int32_t* cu, const int32_t* tok_graph, int64_t n, int64_t D,
                                 int mean, const float* wts, int64_t H, float* out, uint16_t* out_bf16, void* stream) {
  ARK_REQUIRE(src && cu && tok_graph && (out || out_bf16), ARK_E_BADARG, "seg_broadcast: null pointer");
  ARK_REQUIRE(D > 0 && D % 4 == 0 && (!wts || (H > 0 && D % H == 0 && (D / H) % 4 == 0)), ARK_E_SHAPE, "seg_broadcast: bad D / H");
  if (n == 0) return 0;
  seg_broadcast_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(src, cu, tok_graph, n, (int)D, mean, wts,
                                                                                  (int)(H > 0 ? H : 1), out, out_bf16);
  return launched("seg_broadcast");
}

extern "C" int ark_xattn_weights(int64_t n_items, int64_t n_keys, float p_drop, uint64_t seed, uint64_t offset,
                                 const uint64_t* offset_dev, float* wts, void* stream) {
  ARK_REQUIRE(wts && n_keys > 0 && p_drop > 0.f && p_drop < 1.f, ARK_E_BADARG, "xattn_weights: bad arguments");
  if (n_items == 0) return 0;
  xattn_weights_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_items, (int)n_keys, p_drop, seed,
                                                                                            offset, offset_dev, wts);
  return launched("xattn_weights");
}

extern "C" int ark_relu_bwd(const float* d, const uint16_t* out, int64_t n, float scale, uint16_t* dpre, void* stream) {
  ARK_REQUIRE(d && out && dpre, ARK_E_BADARG, "relu_bwd: null pointer");
  ARK_REQUIRE(n % 4 == 0, ARK_E_SHAPE, "relu_bwd: n must be a multiple of 4");
  if (n == 0) return 0;
  relu_bwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, out, n, scale, dpre);
  return launched("relu_bwd");
}
