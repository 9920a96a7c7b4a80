/*
 * arkb200.h — C ABI of libarkb200.so: hand-written sm_100a kernels for the KG-VAE (SAIL) ELBO
 * training step of thiviyanT/ARK.
 *
 * The reference has no FFI of its own (it is pure Python on PyTorch); each entry
 * point below replaces the PyTorch op sequence cited next to it (paths relative to the reference
 * checkout).  The reference-side binding a maintainer would add is the ctypes stub shown in
 * INTEGRATION.md (this repository's copy of it is ark_b200/_C.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates, frees
 *     or retains one.  No torch types cross this boundary.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*), is re-entrant
 *     and keeps no state (autograd's backward thread may call concurrently with the main thread).
 *   - return value: 0 on success, a positive cudaError_t if the CUDA runtime reported one, a
 *     negative ARK_E_* code for an argument/shape/alignment violation detected BEFORE launching.
 *     ark_last_error() returns a thread-local description.
 *   - "rows" of token-level tensors are PACKED non-PAD decoder positions, time-major over graphs
 *     sorted by decreasing length: row = off[t] + b  (ark_b200/layout.py; SURVEY.md finding 7).
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t), f32 = IEEE float, indices int32 unless noted.
 */
#ifndef ARKB200_H_
#define ARKB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARK_ABI_VERSION 1

enum {
  ARK_E_BADARG = -1,    /* null pointer / negative size */
  ARK_E_ALIGN = -2,     /* pointer or leading dimension violates the stated alignment */
  ARK_E_SHAPE = -3,     /* shape outside what the kernel supports */
  ARK_E_NODRIVER = -4,  /* cuTensorMapEncodeTiled could not be resolved from the driver */
};

/* element types for GEMM operands / results */
enum { ARK_F32 = 0, ARK_BF16 = 1 };
/* operand storage: K-major = the contraction index is contiguous (nn.Linear weight [N,K], activations
 * [M,K]); MN-major = the M (or N) index is contiguous, i.e. the operand is stored as [K, M|N]. */
enum { ARK_MAJOR_K = 0, ARK_MAJOR_MN = 1 };
/* GEMM epilogues (applied to acc + bias) */
enum { ARK_EPI_NONE = 0, ARK_EPI_GELU = 1, ARK_EPI_TANH = 2, ARK_EPI_RELU = 3 };

int ark_abi_version(void);
const char* ark_last_error(void);
/* number of kernel launches this thread has enqueued since the last reset (bench.py: gpu_launches) */
int64_t ark_launch_count(void);
void ark_launch_count_reset(void);

/* ---- K1/K2: encoder gather + masked mean-pool (models.py:47-58) and its scatter-add backward ----
 * triples int64 [B,T,3] (the reference's LongTensor, read as is); perm int32 [B] or NULL maps output
 * row b -> input graph perm[b]; E f32 [nE,d], R f32 [nR,d]; pad_rid < 0 means "no padding" (plain mean
 * over T, models.py:58).  Writes g (f32 [B,3d], may be NULL), g_bf16 (may be NULL) and inv_cnt f32 [B]
 * (= 1/max(#valid,1)).  d % 4 == 0. */
int ark_gather_pool_fwd(const int64_t* triples, const int32_t* perm, const float* E, const float* R,
                        int64_t B, int64_t T, int64_t d, int64_t pad_rid,
                        float* g, uint16_t* g_bf16, float* inv_cnt, void* stream);
/* dE[nE,d] / dR[nR,d] += dg[b, slot*d:(slot+1)*d] * inv_cnt[b] for every valid triple; caller zeroes
 * dE/dR first.  Rows pad_eid / pad_rid are never touched (nn.Embedding padding_idx semantics).  n_rel = rows of dR
 * (<= 64: the relation slot is accumulated through a per-CTA histogram, one reduction per relation present; 0 = one
 * reduction per triple as for the entity slots). */
int ark_gather_pool_bwd(const float* dg, const int64_t* triples, const int32_t* perm, const float* inv_cnt,
                        int64_t B, int64_t T, int64_t d, int64_t pad_rid, int64_t pad_eid, int64_t n_rel,
                        float* dE, float* dR, void* stream);

/* ---- a1: packed token ids (utils.py:102-108 layout; PAD-skip) ----
 * seq int64 [B, seq_len]; for t in [0,L), b in [0,bt[t]): row = off[t]+b,
 * tok_in[row] = seq[perm[b], t], tgt[row] = seq[perm[b], t+1]; row_t[row] = t (optional, may be NULL). */
int ark_pack_tokens(const int64_t* seq, const int32_t* perm, const int32_t* bt, const int32_t* off,
                    int64_t B, int64_t seq_len, int64_t L, int32_t* tok_in, int32_t* tgt, int32_t* row_t,
                    void* stream);

/* ---- K5: token-embedding gather (models.py:138) and scatter-add backward ----
 * W is f32 or bf16 [V,d] (w_dtype); X f32 and/or bf16 [N,d] (either may be NULL). d % 8 == 0. */
int ark_tok_gather_fwd(const void* W, int w_dtype, const int32_t* tok, int64_t N, int64_t d, int64_t V,
                       float* X_f32, uint16_t* X_bf16, void* stream);
/* Decoder-only ARK input: X[i] = W[tok[i]] + P[pos[i]] in bf16 (tok_emb + pos_emb, models.py:340-342). */
int ark_tok_pos_gather_fwd(const uint16_t* W, const uint16_t* P, const int32_t* tok, const int32_t* pos,
                           int64_t N, int64_t d, uint16_t* X_bf16, void* stream);
/* dW[tok[i], :] += dX[i, :]  (red.global.add.v4.f32); dW is NOT zeroed here (tied weight: it already
 * holds dLogits^T Y, models.py:130-132). */
int ark_tok_scatter_add(const float* dX, const int32_t* tok, int64_t N, int64_t d, int64_t V,
                        float* dW, void* stream);

/* ---- K4: reparameterisation + analytic KL (models.py:62-63, 199-200) ----
 * heads f32 [B, 2*dz] = [mu | logv_raw] (row stride ld_heads); eps f32 [B,dz] indexed through perm
 * (eps row perm[b] belongs to output row b) or directly if perm == NULL.
 * z f32 [B,dz]; z_bf16 [B, ld_zb] zero-padded to ld_zb columns (may be NULL); kl_acc += kl_scale *
 * sum(-0.5*(1+logv-mu^2-e^logv))  with kl_scale = 1/(B_global*dz); clamp_logv != 0 applies
 * clamp(-10,10) (SAIL's MLP encoder; t-SAIL's has none, models.py:93). */
int ark_reparam_kl_fwd(const float* heads, int64_t ld_heads, const float* eps, const int32_t* perm,
                       int64_t B, int64_t dz, int clamp_logv, float kl_scale,
                       float* z, uint16_t* z_bf16, int64_t ld_zb, float* kl_acc, void* stream);
/* dheads [B,2dz] = [dz + bk*mu + dmu_ext | clampmask*(0.5*dz*eps*sigma + 0.5*bk*(e^logv-1) + dlogv_ext)],
 * bk = beta_kl_scale * (beta_dev ? *beta_dev : 1)   (beta_dev: device scalar, so a replayed CUDA graph follows the
 * per-epoch beta schedule of ablation_study.py:589-591).  dmu_ext / dlogv_ext (NULL or f32 [B,dz], ORIGINAL graph
 * order like eps): upstream gradients of the mu / logv tensors SAIL.forward returns (models.py:317-320). */
int ark_reparam_kl_bwd(const float* heads, int64_t ld_heads, const float* eps, const int32_t* perm,
                       const float* dz_in, int64_t B, int64_t dz, int clamp_logv, float beta_kl_scale,
                       const float* beta_dev, const float* dmu_ext, const float* dlogv_ext,
                       float* dheads, uint16_t* dheads_bf16, int64_t ld_dh, void* stream);

/* ---- K7: fused softmax cross-entropy forward+backward (ablation_study.py:64-69) ----
 * logits [N, ldv] (bf16 or f32 per `dtype`), V valid columns; tgt int32 [N] (never PAD: rows are
 * packed).  loss_acc += grad_scale * sum_i (lse_i - logits_i[tgt_i]); when write_grad != 0 the row is
 * overwritten IN PLACE by grad_scale*(softmax - onehot) (columns [V,ldv) zeroed); lse f32 [N] may be
 * NULL.  The probability matrix never exists in HBM. */
int ark_softmax_ce(void* logits, int dtype, int64_t N, int64_t V, int64_t ldv, const int32_t* tgt,
                   float grad_scale, int write_grad, float* loss_acc, float* lse, void* stream);

/* ---- K3: GEMM  C[M,N] = epi(A[M,K] . B[N,K]^T + bias[N])  (nn.Linear, models.py:36,43-44,120,128) ----
 * tcgen05/TMA tensor-core path: A,B bf16; a_major/b_major as above; lda/ldb/ldc in elements; all
 * leading dimensions * element size and base pointers must be multiples of 16 bytes.
 * c_dtype f32|bf16.  accumulate != 0: C += result (f32 C only).  aux (may be NULL, f32 [M,ldc]):
 * receives the PRE-activation when epilogue != NONE (needed by the backward pass). */
int ark_gemm_bf16_tc(const uint16_t* A, int a_major, int64_t lda, const uint16_t* B, int b_major, int64_t ldb,
                     void* C, int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K,
                     const float* bias, int epilogue, int accumulate, float* aux, void* stream);
/* SIMT path with the same contract for any alignment / tiny shapes and for the fp32 inference path:
 * ab_dtype f32|bf16 (both operands), fp32 FMA accumulation. */
int ark_gemm_simt(const void* A, int a_major, int64_t lda, const void* B, int b_major, int64_t ldb, int ab_dtype,
                  void* C, int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K,
                  const float* bias, int epilogue, int accumulate, float* aux, void* stream);

/* ---- K6: GRU (nn.GRU gate order r,z,n; models.py:121-127,141) ----
 * Pointwise cell, one time step, rows [0,Bt): given gi = W_ih x + b_ih (f32 [Bt,3d], row stride 3d) and
 * gh = W_hh h_prev (NO bias; f32 [Bt,3d]) computes r,z,n, h = (1-z)n + z h_prev and stores
 * h (f32), h_bf16 (optional), and the gates for backward: r,z,n,ghn(=gh_n+b_hn) each f32 [Bt,d].
 * h_next_prev(_bf16): optional second destination, the packed "h_prev" rows of step t+1 (first
 * Bt_next rows only). */
int ark_gru_cell_fwd(const float* gi, const float* gh, const float* b_hh, const float* h_prev,
                     int64_t Bt, int64_t d, float* h, uint16_t* h_bf16,
                     float* hp_next, uint16_t* hp_next_bf16, int64_t Bt_next,
                     float* r, float* z, float* n, float* ghn, void* stream);
/* Backward of one step: dh_total = dy + dh_carry (dh_carry rows >= Bt_carry are treated as 0; may be
 * NULL).  Writes dgi,dgh f32 [Bt,3d] (+ optional bf16 copies) and dh_prev_direct = dh_total*z. */
int ark_gru_cell_bwd(const float* dy, const float* dh_carry, int64_t Bt_carry,
                     const float* r, const float* z, const float* n, const float* ghn, const float* h_prev,
                     int64_t Bt, int64_t d, float* dgi, float* dgh, uint16_t* dgi_bf16, uint16_t* dgh_bf16,
                     float* dh_direct, void* stream);

/* One GRU layer through time over the packed batch (the time loop is native, not Python).
 * bt_host / off_host are HOST int32 arrays of length L (rows of step t: [off[t], off[t]+bt[t]), bt
 * non-increasing).  hp (bf16, may be NULL on the fp32 path) / hp_f32: [N,d] packed "h_prev" rows; the
 * caller fills block 0 with h0, the layer fills blocks 1..L-1.  gi f32 [N,3d] = W_ih x + b_ih.
 * Whh: [3d,d] in w_dtype (ARK_BF16 training path, ARK_F32 inference path).  Outputs y (f32 [N,d]),
 * y_bf16 (optional) and the saved gates r,z,n,ghn (all NULL for inference).  gh_ws f32 [bt[0],3d] scratch.
 * use_tc != 0 runs the recurrent projection on tcgen05 (bf16 only), else on the SIMT GEMM. */
int ark_gru_layer_fwd(void* hp, float* hp_f32, const void* Whh, int w_dtype, const float* gi, const float* b_hh,
                      const int32_t* bt_host, const int32_t* off_host, int64_t L, int64_t d,
                      float* y, uint16_t* y_bf16, float* r, float* z, float* n, float* ghn,
                      float* gh_ws, int use_tc, void* stream);
/* Backward through time.  dy f32 [N,d] (gradient w.r.t. the layer's outputs); writes dgi/dgh [N,3d]
 * (bf16 when w_dtype == ARK_BF16, else f32); dh_a/dh_b: two f32 [bt[0],d] scratch buffers; *dh0_out
 * (HOST pointer to a device pointer) receives whichever of them holds d(loss)/d(h0) at the end. */
int ark_gru_layer_bwd(const float* dy, const float* r, const float* z, const float* n, const float* ghn,
                      const float* hp_f32, const void* Whh, int w_dtype, const int32_t* bt_host,
                      const int32_t* off_host, int64_t L, int64_t d, void* dgi, void* dgh,
                      float* dh_a, float* dh_b, float** dh0_out, int use_tc, void* stream);

/* Persistent GRU layer (cooperative launch, W_hh slice resident in shared memory for all L steps, tcgen05
 * MMA + gate math fused, recurrent state in registers; see ark_b200/csrc/gru_persist.cu).  bt_dev / off_dev
 * are DEVICE int32 arrays [L].  Saved tensors are bf16.  ark_gru_persist_supported returns the hidden-slice
 * width the launcher would use (16/32/64) or 0 when the shape is not supported (d % 64 != 0, or the
 * (d/slice) x ceil(bt0/128) grid does not fit the 148 SMs): callers then use ark_gru_layer_fwd/bwd. */
int ark_gru_persist_supported(int64_t d, int64_t bt0);
/* hp_b [N,d]: block 0 pre-filled with bf16(h0); blocks 1.. are written here.  h0 f32 [bt0,d].
 * gi f32 [N,3d]; outputs y_b and (optional, all or none) r,z,n,ghn bf16 [N,d]; sync_ws int32 [2*ceil(bt0/128)].
 * p_drop > 0: y_b receives the layer output AFTER the inter-layer dropout (nn.GRU(dropout=p), models.py:121-127), drawn
 * from the Philox stream of ark_dropout_bf16 (seed, offset [+ *offset_dev]) over the [N,d] tensor, keep mask -> mask
 * u8 [N,d] (may be NULL); hp_b always holds the undropped state. */
int ark_gru_persist_fwd(uint16_t* hp_b, const float* h0, const uint16_t* Whh_b, const float* gi, const float* b_hh,
                        const int32_t* bt_dev, const int32_t* off_dev, int64_t L, int64_t bt0, int64_t N, int64_t d,
                        uint16_t* y_b, uint16_t* r, uint16_t* z, uint16_t* n, uint16_t* ghn,
                        uint8_t* mask, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                        int32_t* sync_ws, void* stream);
/* dy f32 [N,d].  Writes dgi_b/dgh_b bf16 [N,3d] and dh0 f32 [bt0,d] (+= when dh0_accumulate != 0).
 * Weights: WhhT_b = W_hh^T bf16 [d,3d] (ark_transpose_bf16) for the N-sliced kernel; Whh_b = W_hh bf16 [3d,d] itself for
 * the K-split cluster kernel (ark_gru_persist_bwd_ksplit(d, bt0) != 0: d >= 512, d % 256 == 0 — it consumes the weights
 * untransposed, so callers may pass WhhT_b = NULL there).  dy_mask (u8 [N,d] keep mask of the forward dropout of this
 * layer's output, NULL = none) with p_drop: dy is multiplied by keep/(1-p) on the fly (replaces ark_dropout_bwd). */
int ark_gru_persist_bwd_ksplit(int64_t d, int64_t bt0);
int ark_gru_persist_bwd(const float* dy, const uint16_t* r, const uint16_t* z, const uint16_t* n, const uint16_t* ghn,
                        const uint16_t* hp_b, const uint16_t* WhhT_b, const int32_t* bt_dev, const int32_t* off_dev,
                        int64_t L, int64_t bt0, int64_t N, int64_t d, uint16_t* dgi_b, uint16_t* dgh_b,
                        float* dh0, int dh0_accumulate, const uint16_t* Whh_b, const uint8_t* dy_mask, float p_drop,
                        int32_t* sync_ws, void* stream);
/* Wavefront GRU STACK (ark_b200/csrc/gru_wave.cu): all nl layers x L steps of nn.GRU (models.py:121-127,141;
 * decoder-only :329-343) in ONE cooperative launch per direction; layer k step t runs as soon as layer k-1 step t
 * and layer k step t-1 are done (L+nl-1 dependent steps instead of nl*L).  The input projections W_ih u_t run
 * inside the recurrence, so there is no gi buffer.  Per-layer tensors are stacked: hp_b / out_b / r / z / n / ghn
 * bf16 [nl,N,d]; mask u8 [nl-1,N,d] (NULL when p_drop == 0); dgi_b / dgh_b bf16 [nl,N,3d].  x_b bf16 [N,d] are the
 * layer-0 input rows; out_b[k] is layer k's output AFTER the inter-layer dropout (k < nl-1; Philox stream of
 * ark_dropout_bf16 with offset + k*ceil(N*d/4)), hp_b[k] holds the undropped h_{t-1} rows (block 0 of every
 * layer pre-filled with bf16(h0)).  h0 f32 [bt0,d] is shared by all layers (NULL = zeros).  Wih_b / Whh_b /
 * b_ih / b_hh (and WhhT_b / WihT_b = transposed [d,3d] bf16 weights; WihT_b[0] unused) are HOST arrays of nl
 * device pointers.  dh0 (optional) receives the SUM over layers of d loss / d h0.  sync_ws int32 [nl*ceil(bt0/128)].
 * ark_gru_wave_supported returns the hidden-slice width (16/32) or 0 when the stack does not fit (resident
 * weights 12*slice*d bytes per CTA, (d/slice)*nl CTAs <= 148 for ONE 128-row batch tile, nl <= 4, d % 64 == 0).
 * Batch tiles are independent; when they do not all fit the 148 SMs at once, groups of tiles run back to back. */
int ark_gru_wave_supported(int64_t d, int64_t bt0, int64_t nl);
int ark_gru_wave_fwd(const uint16_t* x_b, uint16_t* hp_b, uint16_t* out_b, const float* h0,
                     const uint16_t* const* Wih_b, const uint16_t* const* Whh_b, const float* const* b_ih,
                     const float* const* b_hh, const int32_t* bt_dev, const int32_t* off_dev, int64_t L, int64_t bt0,
                     int64_t N, int64_t d, int64_t nl, uint16_t* r, uint16_t* z, uint16_t* n, uint16_t* ghn,
                     uint8_t* mask, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                     int32_t* sync_ws, void* stream);
int ark_gru_wave_bwd(const float* dy_top, const uint16_t* r, const uint16_t* z, const uint16_t* n, const uint16_t* ghn,
                     const uint16_t* hp_b, const uint8_t* mask, float p_drop, const uint16_t* const* WhhT_b,
                     const uint16_t* const* WihT_b, const int32_t* bt_dev, const int32_t* off_dev, int64_t L,
                     int64_t bt0, int64_t N, int64_t d, int64_t nl, uint16_t* dgi_b, uint16_t* dgh_b, float* dh0,
                     int32_t* sync_ws, void* stream);
/* Cluster GRU STACK (ark_b200/csrc/gru_cluster.cu): same contract, tensors and Philox stream as ark_gru_wave_*
 * (models.py:121-127,141) for short batch tiles and long chains: one thread-block cluster of d/32 CTAs per (layer,
 * batch tile) keeps W_hh resident and exchanges the recurrent state (forward: h_t slices; backward: bf16 partial
 * sums of dgh_t W_hh, reduce-scattered) through distributed shared memory instead of global memory; the input
 * projections (forward W_ih u_t, backward dgi^{k+1} W_ih^{k+1}) run in separate pipelined CTAs of the same launch.
 * ark_gru_cluster_supported returns the batch-tile rows NB (16/32/64) or 0 (needs d in {128,256,384,512}, nl <= 4,
 * 2*nl*(d/32)*ceil(bt0/NB) <= 148 co-resident CTAs, L <= 4096; queries the device).  ws: scratch of
 * ark_gru_cluster_workspace_bytes(L, bt0, d, nl) bytes (16-byte aligned; the same buffer may serve fwd and bwd);
 * sync_ws int32 [2*nl*ceil(bt0/NB)*16] (one counter per CTA of every stage). */
int ark_gru_cluster_supported(int64_t d, int64_t bt0, int64_t nl, int64_t L);
/* debugging aid: with ARK_GRU_CLUSTER_DBG=<first iteration> in the environment the cluster kernels record a clock64
 * timeline [2 dir][8 blockIdx.z][3 thread roles][4 iterations][16 points] of CTA (0,0) of every stage */
int ark_gru_cluster_debug_dump(int64_t* out_host, int64_t n_words);
/* same for the per-layer persistent kernels (ARK_GRU_PERSIST_DBG=1): [fwd 4 steps x 8 events | bwd 4 x 8] */
int ark_gru_persist_debug_dump(int64_t* out_host, int64_t n_words);
int64_t ark_gru_cluster_workspace_bytes(int64_t L, int64_t bt0, int64_t d, int64_t nl);
int ark_gru_cluster_fwd(const uint16_t* x_b, uint16_t* hp_b, uint16_t* out_b, const float* h0,
                        const uint16_t* const* Wih_b, const uint16_t* const* Whh_b, const float* const* b_ih,
                        const float* const* b_hh, const int32_t* bt_dev, const int32_t* off_dev, int64_t L, int64_t bt0,
                        int64_t N, int64_t d, int64_t nl, uint16_t* r, uint16_t* z, uint16_t* n, uint16_t* ghn,
                        uint8_t* mask, float p_drop, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                        int32_t* sync_ws, void* ws, int64_t ws_bytes, void* stream);
int ark_gru_cluster_bwd(const float* dy_top, const uint16_t* r, const uint16_t* z, const uint16_t* n, const uint16_t* ghn,
                        const uint16_t* hp_b, const uint8_t* mask, float p_drop, const uint16_t* const* WhhT_b,
                        const uint16_t* const* WihT_b, const int32_t* bt_dev, const int32_t* off_dev, int64_t L,
                        int64_t bt0, int64_t N, int64_t d, int64_t nl, uint16_t* dgi_b, uint16_t* dgh_b, float* dh0,
                        int32_t* sync_ws, void* ws, int64_t ws_bytes, void* stream);
/* out[C,R] = in[R,C]^T (bf16) */
int ark_transpose_bf16(const uint16_t* in, int64_t R, int64_t C, uint16_t* out, void* stream);

/* ---- K8/K9 + glue: t-SAIL (Transformer encoder/decoder, models.py:66-114) over ragged, PAD-free, graph-major
 * packed rows: graph b owns rows cu[b]..cu[b+1] (cu int32 [n_graphs+1]); sq_off int64 [n_graphs+1] = prefix sums
 * of n_b^2; the per-(graph, head) [n_b x n_b] score blocks live at sq_off[b]*H + h*n_b^2 of a packed buffer;
 * tok_graph int32 [n_tok] = graph of every row.  See ark_b200/csrc/attn_ops.cu.
 *
 * ark_attn_bgemm: C_p = alpha * A_p . B_p for every (graph, head) p.  Operand kind 0 (TOK) = the graph's rows of a
 * token matrix [n_tok, ld], columns col0 + h*hd .. +hd; kind 1 (SQ) = the score block; *_trans views the operand
 * transposed; *_f32 selects f32 (else bf16).  mode 0: [n x hd].[hd x n] -> SQ (Q.K^T, dO.V^T); mode 1:
 * [n x n].[n x hd] -> TOK (P.V, P^T.dO, dS.K, dS^T.Q).  causal != 0 skips the never-read upper triangle. */
int ark_attn_bgemm(const void* A, int a_kind, int a_trans, int a_f32, int64_t a_ld, int64_t a_col0,
                   const void* B, int b_kind, int b_trans, int b_f32, int64_t b_ld, int64_t b_col0,
                   void* C, int c_kind, int c_f32, int64_t c_ld, int64_t c_col0, const int32_t* cu,
                   const int64_t* sq_off, int64_t n_graphs, int64_t n_max, int64_t H, int64_t hd, int mode,
                   int causal, float alpha, void* stream);
/* P = softmax over keys (j <= i when causal) of the f32 scores; P_drop = dropout(P) (Philox counter = element
 * index / 4) when p_drop > 0, else NULL.  Both bf16, zeros above the diagonal. */
int ark_attn_softmax_fwd(const float* S, const int32_t* cu, const int64_t* sq_off, const int32_t* tok_graph,
                         int64_t n_tok, int64_t H, int causal, float p_drop, uint64_t seed, uint64_t offset,
                         const uint64_t* offset_dev, uint16_t* P, uint16_t* P_drop, void* stream);
/* fp32 inference path: S <- softmax(S) in place (zeros above the diagonal when causal), no dropout */
int ark_attn_softmax_inplace(float* S, const int32_t* cu, const int64_t* sq_off, const int32_t* tok_graph,
                             int64_t n_tok, int64_t H, int causal, void* stream);
/* dS = alpha * P * (dP - sum_j P dP) with dP = dP_drop * keep/(1-p) (keep <=> P_drop != 0) */
int ark_attn_softmax_bwd(const uint16_t* P, const uint16_t* P_drop, const float* dP, const int32_t* cu,
                         const int64_t* sq_off, const int32_t* tok_graph, int64_t n_tok, int64_t H, int causal,
                         float p_drop, float alpha, uint16_t* dS, void* stream);
/* Post-LN residual block (nn.TransformerEncoderLayer/DecoderLayer, norm_first=False): s = res + dropout(branch),
 * y = LayerNorm(s)*gamma+beta.  `branch` f32 [n,D] is OVERWRITTEN with s (kept for the backward); mask u8 [n,D]
 * when p_drop > 0; mean/rstd f32 [n]. */
int ark_add_layernorm_fwd(float* branch, const float* res, const float* gamma, const float* beta, int64_t n,
                          int64_t D, float eps, float p_drop, uint64_t seed, uint64_t offset,
                          const uint64_t* offset_dev, uint8_t* mask, float* y, uint16_t* y_bf16, float* mean,
                          float* rstd, void* stream);
/* d_res = ds (f32); d_branch = ds*keep/(1-p) as f32 and/or bf16; dgamma/dbeta are zeroed then accumulated. */
int ark_add_layernorm_bwd(const float* dy, const float* s, const float* mean, const float* rstd,
                          const float* gamma, int64_t n, int64_t D, float p_drop, const uint8_t* mask,
                          float* d_res, float* d_branch_f32, uint16_t* d_branch_bf16, float* dgamma,
                          float* dbeta, void* stream);
/* X[r, slot*d..] = (slot==1 ? R : E)[idx[r,slot]] for real triples (idx int32 [n,3]); models.py:80-83 */
int ark_triple_embed_fwd(const int32_t* idx, const float* E, const float* R, int64_t n, int64_t d, float* X,
                         uint16_t* X_bf16, void* stream);
int ark_triple_embed_bwd(const int32_t* idx, const float* dX, int64_t n, int64_t d, float* dE, float* dR,
                         void* stream);
/* X[r] = W[tok[r]] + P[pos[r]] (f32 masters; f32 + bf16 outputs); models.py:109-110 */
int ark_embed_sum_fwd(const float* W, const float* P, const int32_t* tok, const int32_t* pos, int64_t n,
                      int64_t d, float* X, uint16_t* X_bf16, void* stream);
/* out[b] = (mean ? 1/n_b : 1) * sum_{r in graph b} w[r, head(col)] * X[r]   (wts f32 [n_tok,H] or NULL) */
int ark_seg_reduce(const float* X, const int32_t* cu, int64_t n_graphs, int64_t D, int mean, const float* wts,
                   int64_t H, float* out, uint16_t* out_bf16, void* stream);
/* out[r] = (mean ? 1/n_b : 1) * w[r, head(col)] * src[graph(r)] */
int ark_seg_broadcast(const float* src, const int32_t* cu, const int32_t* tok_graph, int64_t n, int64_t D,
                      int mean, const float* wts, int64_t H, float* out, uint16_t* out_bf16, void* stream);
/* wts[i] = Binomial(n_keys, 1-p) / ((1-p) n_keys): attention dropout of the collapsed cross-attention */
int ark_xattn_weights(int64_t n_items, int64_t n_keys, float p_drop, uint64_t seed, uint64_t offset,
                      const uint64_t* offset_dev, float* wts, void* stream);
/* dpre = (out > 0) ? d*scale : 0  (ReLU backward; folds the in-place FFN dropout backward) */
int ark_relu_bwd(const float* d, const uint16_t* out, int64_t n, float scale, uint16_t* dpre, void* stream);

/* ---- elementwise helpers ----
 * dpre = dact * gelu'(pre) (erf form, models.py:37) -> f32 and/or bf16 */
int ark_gelu_bwd(const float* dact, const float* pre, int64_t n, float* dpre, uint16_t* dpre_bf16, void* stream);
/* dpre = dh * (1 - h^2) */
int ark_tanh_bwd(const float* dh, const float* h, int64_t n, float* dpre, uint16_t* dpre_bf16, void* stream);
/* out[N] (+)= column sums of X[M, ld] (f32 or bf16) — bias gradients.  accumulate: 0 = overwrite, 1 = add to out,
 * 2 = overwrite with a fixed summation order (bitwise reproducible: one CTA per column strip, no cross-CTA atomics) */
int ark_colsum(const void* X, int dtype, int64_t M, int64_t N, int64_t ld, float* out, int accumulate, void* stream);
/* y = a + b (f32), optional bf16 copy; y may alias a */
int ark_add_f32(const float* a, const float* b, int64_t n, float* y, uint16_t* y_bf16, void* stream);
int ark_cast_f32_to_bf16(const float* x, int64_t n, uint16_t* y, void* stream);
/* inter-layer dropout (nn.GRU dropout=dec_dropout, train mode): y = x * keep/(1-p), Philox4x32-10 keyed
 * by (seed, offset + *offset_dev + element); offset_dev (device uint64, may be NULL) lets a captured CUDA graph
 * draw fresh masks on every replay; mask uint8 [n] saved for backward. */
int ark_dropout_fwd(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset,
                    float* y, uint16_t* y_bf16, uint8_t* mask, const uint64_t* offset_dev, void* stream);
/* same Philox stream on a bf16 tensor (the GRU layer outputs of the training path); y may alias x */
int ark_dropout_bf16(const uint16_t* x, int64_t n, float p, uint64_t seed, uint64_t offset,
                     uint16_t* y, uint8_t* mask, const uint64_t* offset_dev, void* stream);
int ark_dropout_bwd(const float* dy, const uint8_t* mask, int64_t n, float p, float* dx, void* stream);

/* ---- K10: dense Adam over a flat parameter buffer (torch.optim.Adam defaults, ablation_study.py:571) ----
 * p,g,m,v f32 [n]; shadow bf16 [n] may be NULL; step = 1-based count of this update; grad_scale
 * multiplies g first (1 for summed DP gradients). */
int ark_adam_flat(float* p, const float* g, float* m, float* v, uint16_t* shadow, int64_t n,
                  float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream);
/* Same update with the step-dependent scalars in DEVICE memory: hyper[0] = lr/(1-beta1^step),
 * hyper[1] = 1/sqrt(1-beta2^step) — so the launch can be replayed from a CUDA graph. */
int ark_adam_flat_dyn(float* p, const float* g, float* m, float* v, uint16_t* shadow, int64_t n,
                      const float* hyper, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---- K11: data-parallel gradient reduce-scatter + SHARDED Adam + parameter broadcast, one kernel over NVSwitch multicast ----
 * Replaces, under data parallelism, `all-reduce(grad); optimizer.step()` (the reference has no multi-GPU path; the
 * single-GPU statement is ablation_study.py:596-600).  The flat buffers of every rank live in one symmetric allocation that is
 * also mapped as a multicast object: grad_mc / param_mc / shadow_mc are MULTICAST addresses of the f32 gradient, f32
 * parameter and bf16 shadow buffers; param / m / v are this rank's own (unicast) buffers.  For each span [begin, end)
 * (elements, multiples of 4; host arrays) rank r owns the r-th 1/world of it: it reads the SUM of all ranks' gradients
 * with multimem.ld_reduce, applies Adam there (m, v are only maintained for owned slices) and multicasts the new
 * parameters + bf16 copies to every rank.  mode 0: gradient all-reduce only (the sum is multicast back into grad).
 * peer_flags: host array [world] of every rank's 512-byte zero-initialised flag block (peer-mapped addresses); the
 * kernel exchanges "ready"/"done" epochs through them, so every rank must issue the same sequence of calls.
 * hyper != NULL: step-dependent scalars in device memory as in ark_adam_flat_dyn.  ctas <= 0: default (32). */
int ark_dp_reduce_adam(float* grad_mc, float* param_mc, uint16_t* shadow_mc, const float* param, float* m, float* v,
                       uint32_t* const* peer_flags, int rank, int world, const int64_t* span_begin,
                       const int64_t* span_end, int n_spans, int mode, float lr, float beta1, float beta2, float eps,
                       int64_t step, const float* hyper, int ctas, void* stream);

/* All-gather by multicast store (same symmetric-memory protocol and flag blocks as ark_dp_reduce_adam; the calls of both
 * kinds must be issued in the same order on every rank, on one stream): item i = nbytes[i] bytes at src[i] (this rank's
 * rows) written to dst_mc[i] = the MULTICAST address of this rank's slot in the gathered buffer, i.e. into every rank's
 * copy.  Used for the [B, 3d] factors of the encoder-MLP weight gradients (dW = dY_all^T X_all).  src / dst_mc / nbytes:
 * host arrays [n_items], n_items <= 8; 16-byte granularity.  ctas <= 0: default (16). */
int ark_dp_allgather_mc(const void* const* src, void* const* dst_mc, const int64_t* nbytes, int n_items,
                        uint32_t* const* peer_flags, int rank, int world, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARKB200_H_ */
