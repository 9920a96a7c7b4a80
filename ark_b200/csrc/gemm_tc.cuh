// Internal interface of the tcgen05 GEMM (gemm_tc.cu) for the other translation units (GRU sequence driver).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ark {

struct EpiParams {
  void* C;
  float* aux;
  const float* bias;
  int64_t ldc;
  int c_bf16;
  int epilogue;
  int accumulate;
  int swap_raster;   // set by the launcher
  int kb_per_split;  // set by the launcher: k-blocks per blockIdx.y slice (split-K: partial sums meet through f32 atomics)
  int atomic;        // set by the launcher: C += via red.global.add (split-K)
  int tma_store;     // set by the launcher: full bf16 tiles leave through a TMA store (persistent kernel)
};

int tc_pick_bn(int64_t M, int64_t N);
// tile_rows = 128 for the A operand, BN for the B operand (only used for K-major operands)
int tc_make_operand_map(CUtensorMap* tm, const uint16_t* P, int major, int64_t rows, int64_t K, int64_t ld, int tile_rows);
int tc_check_operands(const char* who, const void* A, int a_major, int64_t lda, const void* B, int b_major, int64_t ldb,
                      int64_t M, int64_t N, int64_t K);
// a_row0 / b_row0 are added to the M / N coordinate of every TMA load (sub-matrix of a mapped tensor)
int tc_enqueue(const CUtensorMap& tmA, const CUtensorMap& tmB, int a_major, int b_major, int BN, const EpiParams& ep,
               int M, int N, int K, int a_row0, int b_row0, cudaStream_t s);

}  // namespace ark
