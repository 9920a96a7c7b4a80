// a1/K5: packed (PAD-skipping) token ids, token-embedding gather and its scatter-add backward.
// Replaces the reference's nn.Embedding lookup of the decoder (kgvae/model/models.py:138) and
// autograd's embedding_dense_backward; the token layout is utils.py:102-108.
#include "common.cuh"

namespace ark {

// grid (L): block t writes rows off[t] .. off[t]+bt[t]
__global__ void __launch_bounds__(256) pack_tokens_kernel(
    const int64_t* __restrict__ seq, const int32_t* __restrict__ perm, const int32_t* __restrict__ bt,
    const int32_t* __restrict__ off, int seq_len, int32_t* __restrict__ tok_in, int32_t* __restrict__ tgt,
    int32_t* __restrict__ row_t) {
  const int t = blockIdx.x;
  const int n = bt[t], base = off[t];
  for (int b = threadIdx.x; b < n; b += blockDim.x) {
    const int64_t* row = seq + (int64_t)(perm ? perm[b] : b) * seq_len;
    tok_in[base + b] = (int32_t)row[t];
    tgt[base + b] = (int32_t)row[t + 1];
    if (row_t) row_t[base + b] = t;
  }
}

// X[row] = W[tok[row]] + P[pos[row]]  (decoder-only ARK: token + position embedding, models.py:340-342)
__global__ void __launch_bounds__(256) tok_pos_gather_kernel(
    const uint16_t* __restrict__ W, const uint16_t* __restrict__ P, const int32_t* __restrict__ tok,
    const int32_t* __restrict__ pos, int64_t N, int d, uint16_t* __restrict__ Xb) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const int64_t sw = (int64_t)tok[row] * d, sp = (int64_t)pos[row] * d;
  for (int c = lane * 8; c < d; c += 256) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(W + sw + c));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(P + sp + c));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = unpack_bf16x2(aw[k]), y = unpack_bf16x2(bw[k]);
      o[k] = pack_bf16x2(x.x + y.x, x.y + y.y);
    }
    *reinterpret_cast<uint4*>(Xb + row * d + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one warp per token row; 16-byte chunks
template <typename TW>
__global__ void __launch_bounds__(256) tok_gather_kernel(
    const TW* __restrict__ W, const int32_t* __restrict__ tok, int64_t N, int d, float* __restrict__ Xf,
    uint16_t* __restrict__ Xb) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const int64_t src = (int64_t)tok[row] * d;
  if constexpr (sizeof(TW) == 2) {
    // bf16 table: 8 elements per 16 bytes
    for (int c = lane * 8; c < d; c += 256) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(W + src + c));
      if (Xb) *reinterpret_cast<uint4*>(Xb + row * d + c) = v;
      if (Xf) {
        const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), e = unpack_bf16x2(v.z), f = unpack_bf16x2(v.w);
        *reinterpret_cast<float4*>(Xf + row * d + c) = make_float4(a.x, a.y, b.x, b.y);
        *reinterpret_cast<float4*>(Xf + row * d + c + 4) = make_float4(e.x, e.y, f.x, f.y);
      }
    }
  } else {
    for (int c = lane * 4; c < d; c += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(W + src + c));
      if (Xf) *reinterpret_cast<float4*>(Xf + row * d + c) = v;
      if (Xb) {
        uint2 p;
        p.x = pack_bf16x2(v.x, v.y);
        p.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(Xb + row * d + c) = p;
      }
    }
  }
}

__global__ void __launch_bounds__(256) tok_scatter_add_kernel(
    const float* __restrict__ dX, const int32_t* __restrict__ tok, int64_t N, int d, float* __restrict__ dW) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  float* dst = dW + (int64_t)tok[row] * d;
  for (int c = lane * 4; c < d; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(dX + row * d + c);
    red_add_v4(dst + c, v);
  }
}

}  // namespace ark

using namespace ark;

extern "C" int ark_pack_tokens(const int64_t* seq, const int32_t* perm, const int32_t* bt, const int32_t* off,
                               int64_t B, int64_t seq_len, int64_t L, int32_t* tok_in, int32_t* tgt, int32_t* row_t,
                               void* stream) {
  ARK_REQUIRE(seq && bt && off && tok_in && tgt, ARK_E_BADARG, "pack_tokens: null pointer");
  ARK_REQUIRE(B > 0 && L > 0 && L < seq_len, ARK_E_BADARG, "pack_tokens: need 0 < L < seq_len");
  pack_tokens_kernel<<<(unsigned)L, 256, 0, (cudaStream_t)stream>>>(seq, perm, bt, off, (int)seq_len, tok_in, tgt,
                                                                    row_t);
  return launched("pack_tokens");
}

extern "C" int ark_tok_gather_fwd(const void* W, int w_dtype, const int32_t* tok, int64_t N, int64_t d, int64_t V,
                                  float* X_f32, uint16_t* X_bf16, void* stream) {
  ARK_REQUIRE(W && tok && (X_f32 || X_bf16), ARK_E_BADARG, "tok_gather_fwd: null pointer");
  ARK_REQUIRE(N >= 0 && d > 0 && V > 0, ARK_E_BADARG, "tok_gather_fwd: bad sizes");
  ARK_REQUIRE(d % 8 == 0, ARK_E_SHAPE, "tok_gather_fwd: d=%lld must be a multiple of 8", (long long)d);
  ARK_REQUIRE(aligned16(W) && (!X_f32 || aligned16(X_f32)) && (!X_bf16 || aligned16(X_bf16)), ARK_E_ALIGN,
              "tok_gather_fwd: 16-byte alignment");
  if (N == 0) return 0;
  const unsigned grid = (unsigned)((N + 7) / 8);
  if (w_dtype == ARK_BF16)
    tok_gather_kernel<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)W, tok, N, (int)d, X_f32, X_bf16);
  else if (w_dtype == ARK_F32)
    tok_gather_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)W, tok, N, (int)d, X_f32, X_bf16);
  else
    return fail(ARK_E_BADARG, "tok_gather_fwd: unknown dtype %d", w_dtype);
  return launched("tok_gather_fwd");
}

extern "C" int ark_tok_pos_gather_fwd(const uint16_t* W, const uint16_t* P, const int32_t* tok, const int32_t* pos,
                                      int64_t N, int64_t d, uint16_t* X_bf16, void* stream) {
  ARK_REQUIRE(W && P && tok && pos && X_bf16, ARK_E_BADARG, "tok_pos_gather_fwd: null pointer");
  ARK_REQUIRE(N >= 0 && d > 0 && d % 8 == 0, ARK_E_SHAPE, "tok_pos_gather_fwd: d must be a positive multiple of 8");
  ARK_REQUIRE(aligned16(W) && aligned16(P) && aligned16(X_bf16), ARK_E_ALIGN, "tok_pos_gather_fwd: 16-byte alignment");
  if (N == 0) return 0;
  tok_pos_gather_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(W, P, tok, pos, N, (int)d, X_bf16);
  return launched("tok_pos_gather_fwd");
}

extern "C" int ark_tok_scatter_add(const float* dX, const int32_t* tok, int64_t N, int64_t d, int64_t V, float* dW,
                                   void* stream) {
  ARK_REQUIRE(dX && tok && dW, ARK_E_BADARG, "tok_scatter_add: null pointer");
  ARK_REQUIRE(N >= 0 && d > 0 && V > 0, ARK_E_BADARG, "tok_scatter_add: bad sizes");
  ARK_REQUIRE(d % 4 == 0, ARK_E_SHAPE, "tok_scatter_add: d must be a multiple of 4");
  ARK_REQUIRE(aligned16(dX) && aligned16(dW), ARK_E_ALIGN, "tok_scatter_add: 16-byte alignment");
  if (N == 0) return 0;
  tok_scatter_add_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dX, tok, N, (int)d, dW);
  return launched("tok_scatter_add");
}
