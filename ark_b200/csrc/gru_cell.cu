// K6 (pointwise part): one GRU time step's gate math, forward and backward, given the two projections.
// torch.nn.GRU semantics (kgvae/model/models.py:121-127,141):
//   r = sig(gi_r + gh_r + b_hr); z = sig(gi_z + gh_z + b_hz); n = tanh(gi_n + r*(gh_n + b_hn));
//   h = (1-z)*n + z*h_prev                      (gi already contains b_ih; gh = W_hh h_prev WITHOUT bias)
// Used by the SIMT/fp32 path and as the numerical twin of the fused tcgen05 step kernel's epilogue
// (gru_step_tc.cu), which must produce bit-identical gate arithmetic given the same gh.
#include "common.cuh"
#include "gru_math.cuh"

namespace ark {

__global__ void __launch_bounds__(256) gru_cell_fwd_kernel(
    const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ b_hh,
    const float* __restrict__ h_prev, int Bt, int d, float* __restrict__ h, uint16_t* __restrict__ h_bf16,
    float* __restrict__ hp_next, uint16_t* __restrict__ hp_next_bf16, int Bt_next, float* __restrict__ r_out,
    float* __restrict__ z_out, float* __restrict__ n_out, float* __restrict__ ghn_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Bt * d) return;
  const int b = (int)(i / d), j = (int)(i - (int64_t)b * d);
  const int64_t g3 = (int64_t)b * 3 * d;
  GruFwd o = gru_fwd_math(gi[g3 + j], gi[g3 + d + j], gi[g3 + 2 * d + j], gh[g3 + j] + b_hh[j],
                          gh[g3 + d + j] + b_hh[d + j], gh[g3 + 2 * d + j] + b_hh[2 * d + j], h_prev[i]);
  h[i] = o.h;
  if (h_bf16) h_bf16[i] = f32_to_bf16_bits(o.h);
  if (b < Bt_next) {
    if (hp_next) hp_next[i] = o.h;
    if (hp_next_bf16) hp_next_bf16[i] = f32_to_bf16_bits(o.h);
  }
  if (r_out) {
    r_out[i] = o.r; z_out[i] = o.z; n_out[i] = o.n; ghn_out[i] = o.ghn;
  }
}

__global__ void __launch_bounds__(256) gru_cell_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ dh_carry, int Bt_carry, const float* __restrict__ r,
    const float* __restrict__ z, const float* __restrict__ n, const float* __restrict__ ghn,
    const float* __restrict__ h_prev, int Bt, int d, float* __restrict__ dgi, float* __restrict__ dgh,
    uint16_t* __restrict__ dgi_bf16, uint16_t* __restrict__ dgh_bf16, float* __restrict__ dh_direct) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Bt * d) return;
  const int b = (int)(i / d), j = (int)(i - (int64_t)b * d);
  float dh = dy ? dy[i] : 0.f;
  if (dh_carry && b < Bt_carry) dh += dh_carry[i];
  GruBwd o = gru_bwd_math(dh, r[i], z[i], n[i], ghn[i], h_prev[i]);
  const int64_t g3 = (int64_t)b * 3 * d;
  if (dgi) {
    dgi[g3 + j] = o.dar; dgi[g3 + d + j] = o.daz; dgi[g3 + 2 * d + j] = o.dan;
  }
  if (dgh) {
    dgh[g3 + j] = o.dar; dgh[g3 + d + j] = o.daz; dgh[g3 + 2 * d + j] = o.dan_r;
  }
  if (dgi_bf16) {
    dgi_bf16[g3 + j] = f32_to_bf16_bits(o.dar);
    dgi_bf16[g3 + d + j] = f32_to_bf16_bits(o.daz);
    dgi_bf16[g3 + 2 * d + j] = f32_to_bf16_bits(o.dan);
  }
  if (dgh_bf16) {
    dgh_bf16[g3 + j] = f32_to_bf16_bits(o.dar);
    dgh_bf16[g3 + d + j] = f32_to_bf16_bits(o.daz);
    dgh_bf16[g3 + 2 * d + j] = f32_to_bf16_bits(o.dan_r);
  }
  dh_direct[i] = o.dh_prev;
}

}  // namespace ark

using namespace ark;

extern "C" int ark_gru_cell_fwd(const float* gi, const float* gh, const float* b_hh, const float* h_prev, int64_t Bt,
                                int64_t d, float* h, uint16_t* h_bf16, float* hp_next, uint16_t* hp_next_bf16,
                                int64_t Bt_next, float* r, float* z, float* n, float* ghn, void* stream) {
  ARK_REQUIRE(gi && gh && b_hh && h_prev && h, ARK_E_BADARG, "gru_cell_fwd: null pointer");
  ARK_REQUIRE((r && z && n && ghn) || (!r && !z && !n && !ghn), ARK_E_BADARG,
              "gru_cell_fwd: gate outputs must be all set or all NULL");
  ARK_REQUIRE(Bt >= 0 && d > 0 && Bt_next <= Bt, ARK_E_BADARG, "gru_cell_fwd: bad sizes");
  if (Bt == 0) return 0;
  const int64_t n_el = Bt * d;
  gru_cell_fwd_kernel<<<(unsigned)((n_el + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      gi, gh, b_hh, h_prev, (int)Bt, (int)d, h, h_bf16, hp_next, hp_next_bf16, (int)Bt_next, r, z, n, ghn);
  return launched("gru_cell_fwd");
}

extern "C" int ark_gru_cell_bwd(const float* dy, const float* dh_carry, int64_t Bt_carry, const float* r,
                                const float* z, const float* n, const float* ghn, const float* h_prev, int64_t Bt,
                                int64_t d, float* dgi, float* dgh, uint16_t* dgi_bf16, uint16_t* dgh_bf16,
                                float* dh_direct, void* stream) {
  ARK_REQUIRE(r && z && n && ghn && h_prev && dh_direct, ARK_E_BADARG, "gru_cell_bwd: null pointer");
  ARK_REQUIRE(Bt >= 0 && d > 0 && Bt_carry >= 0, ARK_E_BADARG, "gru_cell_bwd: bad sizes");
  if (Bt == 0) return 0;
  const int64_t n_el = Bt * d;
  gru_cell_bwd_kernel<<<(unsigned)((n_el + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      dy, dh_carry, (int)Bt_carry, r, z, n, ghn, h_prev, (int)Bt, (int)d, dgi, dgh, dgi_bf16, dgh_bf16, dh_direct);
  return launched("gru_cell_bwd");
}
