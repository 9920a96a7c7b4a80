// K6: one GRU layer over a packed, length-sorted batch — forward and backward THROUGH TIME.
// Replaces nn.GRU's per-layer recurrence (kgvae/model/models.py:121-127,141; cuDNN/ATen RNN in the reference)
// for rows packed time-major: step t owns rows [off[t], off[t]+bt[t]) and bt is non-increasing (graphs
// sorted by decreasing length), so the active set of step t is a prefix of step t-1's (PAD positions are
// never computed — exact, SURVEY.md finding 7).
//
// The time loop lives here, in native code, not in Python: per step it enqueues the recurrent projection
// (tcgen05 GEMM on a row window of ONE tensor map built per layer, or the SIMT GEMM on the fp32 path)
// and the fused gate kernel.  bt/off are HOST arrays (they are known when the batch is assembled).
#include "common.cuh"
#include "gemm_tc.cuh"

extern "C" int ark_gru_layer_fwd(void* hp, float* hp_f32, const void* Whh, int w_dtype, const float* gi,
                                 const float* b_hh, const int32_t* bt_host, const int32_t* off_host, int64_t L,
                                 int64_t d, float* y, uint16_t* y_bf16, float* r, float* z, float* n, float* ghn,
                                 float* gh_ws, int use_tc, void* stream) {
  using namespace ark;
  ARK_REQUIRE(hp_f32 && Whh && gi && b_hh && bt_host && off_host && y && gh_ws, ARK_E_BADARG,
              "gru_layer_fwd: null pointer");
  ARK_REQUIRE(L > 0 && d > 0, ARK_E_BADARG, "gru_layer_fwd: bad sizes");
  ARK_REQUIRE(w_dtype == ARK_F32 || (w_dtype == ARK_BF16 && hp), ARK_E_BADARG,
              "gru_layer_fwd: bf16 weights need the bf16 h_prev buffer");
  ARK_REQUIRE(!(use_tc && w_dtype != ARK_BF16), ARK_E_BADARG, "gru_layer_fwd: tensor-core path needs bf16 weights");
  for (int64_t t = 0; t + 1 < L; ++t)
    ARK_REQUIRE(bt_host[t + 1] <= bt_host[t], ARK_E_SHAPE, "gru_layer_fwd: bt must be non-increasing (sort by length)");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_rows = (int64_t)off_host[L - 1] + bt_host[L - 1];
  uint16_t* hp_b = reinterpret_cast<uint16_t*>(hp);

  CUtensorMap tmA, tmB;
  int BN = 64, rc;
  if (use_tc) {
    if ((rc = tc_check_operands("gru_layer_fwd", hp_b, ARK_MAJOR_K, d, Whh, ARK_MAJOR_K, d, bt_host[0], 3 * d, d))) return rc;
    BN = tc_pick_bn(bt_host[0], 3 * d);
    if ((rc = tc_make_operand_map(&tmA, hp_b, ARK_MAJOR_K, n_rows, d, d, 128))) return rc;
    if ((rc = tc_make_operand_map(&tmB, (const uint16_t*)Whh, ARK_MAJOR_K, 3 * d, d, d, BN))) return rc;
  }
  for (int64_t t = 0; t < L; ++t) {
    const int64_t Bt = bt_host[t], o = off_host[t];
    if (Bt == 0) break;
    const int64_t Bn = (t + 1 < L) ? bt_host[t + 1] : 0, on = (t + 1 < L) ? off_host[t + 1] : 0;
    // gh = h_prev . W_hh^T  (no bias; the gate kernel adds b_hh)
    if (use_tc) {
      EpiParams ep;
      ep.C = gh_ws; ep.aux = nullptr; ep.bias = nullptr; ep.ldc = 3 * d; ep.c_bf16 = 0; ep.epilogue = ARK_EPI_NONE;
      ep.accumulate = 0;
      if ((rc = tc_enqueue(tmA, tmB, ARK_MAJOR_K, ARK_MAJOR_K, BN, ep, (int)Bt, (int)(3 * d), (int)d, (int)o, 0, s))) return rc;
    } else {
      const void* a = (w_dtype == ARK_BF16) ? (const void*)(hp_b + o * d) : (const void*)(hp_f32 + o * d);
      if ((rc = ark_gemm_simt(a, ARK_MAJOR_K, d, Whh, ARK_MAJOR_K, d, w_dtype, gh_ws, ARK_F32, 3 * d, Bt, 3 * d, d,
                              nullptr, ARK_EPI_NONE, 0, nullptr, stream)))
        return rc;
    }
    if ((rc = ark_gru_cell_fwd(gi + o * 3 * d, gh_ws, b_hh, hp_f32 + o * d, Bt, d, y + o * d,
                               y_bf16 ? y_bf16 + o * d : nullptr, Bn ? hp_f32 + on * d : nullptr,
                               (Bn && hp_b) ? hp_b + on * d : nullptr, Bn, r ? r + o * d : nullptr,
                               z ? z + o * d : nullptr, n ? n + o * d : nullptr, ghn ? ghn + o * d : nullptr, stream)))
      return rc;
  }
  return 0;
}

extern "C" int ark_gru_layer_bwd(const float* dy, const float* r, const float* z, const float* n, const float* ghn,
                                 const float* hp_f32, const void* Whh, int w_dtype, const int32_t* bt_host,
                                 const int32_t* off_host, int64_t L, int64_t d, void* dgi, void* dgh, float* dh_a,
                                 float* dh_b, float** dh0_out, int use_tc, void* stream) {
  using namespace ark;
  ARK_REQUIRE(dy && r && z && n && ghn && hp_f32 && Whh && bt_host && off_host && dgi && dgh && dh_a && dh_b && dh0_out,
              ARK_E_BADARG, "gru_layer_bwd: null pointer");
  ARK_REQUIRE(L > 0 && d > 0, ARK_E_BADARG, "gru_layer_bwd: bad sizes");
  ARK_REQUIRE(!(use_tc && w_dtype != ARK_BF16), ARK_E_BADARG, "gru_layer_bwd: tensor-core path needs bf16 weights");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n_rows = (int64_t)off_host[L - 1] + bt_host[L - 1];
  const bool bf = (w_dtype == ARK_BF16);
  // bf16 path: dgi/dgh are bf16 [N,3d] (GEMM operands); fp32 path: f32 [N,3d]
  uint16_t* dgi_b = bf ? reinterpret_cast<uint16_t*>(dgi) : nullptr;
  uint16_t* dgh_b = bf ? reinterpret_cast<uint16_t*>(dgh) : nullptr;
  float* dgi_f = bf ? nullptr : reinterpret_cast<float*>(dgi);
  float* dgh_f = bf ? nullptr : reinterpret_cast<float*>(dgh);

  CUtensorMap tmA, tmB;
  int BN = 64, rc;
  if (use_tc) {
    // dh_prev[Bt,d] += dgh[Bt,3d] . W_hh[3d,d] : A K-major (K=3d), B stored [K=3d, N=d] = MN-major
    if ((rc = tc_check_operands("gru_layer_bwd", dgh_b, ARK_MAJOR_K, 3 * d, Whh, ARK_MAJOR_MN, d, bt_host[0], d, 3 * d))) return rc;
    BN = tc_pick_bn(bt_host[0], d);
    if ((rc = tc_make_operand_map(&tmA, dgh_b, ARK_MAJOR_K, n_rows, 3 * d, 3 * d, 128))) return rc;
    if ((rc = tc_make_operand_map(&tmB, (const uint16_t*)Whh, ARK_MAJOR_MN, d, 3 * d, d, BN))) return rc;
  }
  float* cur = dh_a;   // receives dh_prev of the step being processed
  float* prev = dh_b;  // holds the carry produced by step t+1
  int64_t carry_rows = 0;
  for (int64_t t = L - 1; t >= 0; --t) {
    const int64_t Bt = bt_host[t], o = off_host[t];
    if (Bt == 0) continue;
    if ((rc = ark_gru_cell_bwd(dy + o * d, carry_rows ? prev : nullptr, carry_rows, r + o * d, z + o * d, n + o * d,
                               ghn + o * d, hp_f32 + o * d, Bt, d, dgi_f ? dgi_f + o * 3 * d : nullptr,
                               dgh_f ? dgh_f + o * 3 * d : nullptr, dgi_b ? dgi_b + o * 3 * d : nullptr,
                               dgh_b ? dgh_b + o * 3 * d : nullptr, cur, stream)))
      return rc;
    if (use_tc) {
      EpiParams ep;
      ep.C = cur; ep.aux = nullptr; ep.bias = nullptr; ep.ldc = d; ep.c_bf16 = 0; ep.epilogue = ARK_EPI_NONE;
      ep.accumulate = 1;
      if ((rc = tc_enqueue(tmA, tmB, ARK_MAJOR_K, ARK_MAJOR_MN, BN, ep, (int)Bt, (int)d, (int)(3 * d), (int)o, 0, s))) return rc;
    } else {
      const void* a = bf ? (const void*)(dgh_b + o * 3 * d) : (const void*)(dgh_f + o * 3 * d);
      if ((rc = ark_gemm_simt(a, ARK_MAJOR_K, 3 * d, Whh, ARK_MAJOR_MN, d, w_dtype, cur, ARK_F32, d, Bt, d, 3 * d, nullptr,
                              ARK_EPI_NONE, 1, nullptr, stream)))
        return rc;
    }
    carry_rows = Bt;
    float* tmp = cur; cur = prev; prev = tmp;
  }
  *dh0_out = prev;  // [bt[0], d] gradient w.r.t. this layer's initial state
  return 0;
}
