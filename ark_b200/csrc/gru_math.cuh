// Gate arithmetic of one GRU cell element, shared by the pointwise kernels and the fused tensor-core
// step kernels so both evaluate exactly the same expression tree.
#pragma once
#include "common.cuh"

namespace ark {

struct GruFwd { float r, z, n, ghn, h; };
struct GruBwd { float dar, daz, dan, dan_r, dh_prev; };

// gi_* include b_ih; gh_* include b_hh.
__device__ __forceinline__ GruFwd gru_fwd_math(float gi_r, float gi_z, float gi_n, float gh_r, float gh_z, float gh_n,
                                               float h_prev) {
  GruFwd o;
  o.r = sigmoidf_(gi_r + gh_r);
  o.z = sigmoidf_(gi_z + gh_z);
  o.ghn = gh_n;
  o.n = tanhf(fmaf(o.r, gh_n, gi_n));
  o.h = fmaf(o.z, h_prev - o.n, o.n);  // (1-z)*n + z*h_prev
  return o;
}

// dh = total gradient flowing into h_t.
__device__ __forceinline__ GruBwd gru_bwd_math(float dh, float r, float z, float n, float ghn, float h_prev) {
  GruBwd o;
  const float dn = dh * (1.f - z);
  const float dzg = dh * (h_prev - n);
  o.dan = dn * (1.f - n * n);
  o.dar = o.dan * ghn * r * (1.f - r);
  o.daz = dzg * z * (1.f - z);
  o.dan_r = o.dan * r;
  o.dh_prev = dh * z;
  return o;
}

}  // namespace ark
