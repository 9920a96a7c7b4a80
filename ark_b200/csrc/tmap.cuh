// Host-side TMA descriptor construction.  cuTensorMapEncodeTiled is resolved from the driver at run time
// (cudaGetDriverEntryPoint) so libarkb200.so does not link against libcuda and loads on a GPU-less box.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ark {

// 2-D bf16 tensor, row-major with `ld` elements between rows: dims {inner, outer}; box {box_inner, box_outer};
// 128-byte swizzle (box_inner must be 64 elements); out-of-bounds elements read as zero.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer);

// The same [rows, K] bf16 matrix seen as K/64 chunks of 64 elements: dims {64, rows, K/64}; one TMA op with box
// {64, box_rows, box_chunks} lands box_chunks consecutive 128B-swizzled [box_rows x 128 B] slabs in shared memory
// (the K-major UMMA operand layout), so a whole K panel costs ONE bulk-tensor instruction and ONE mbarrier.
int make_tmap_kchunked_bf16(CUtensorMap* out, const void* base, uint64_t K, uint64_t rows, uint64_t ld,
                            uint32_t box_rows, uint32_t box_chunks);

// 2-D bf16 tensor WITHOUT swizzle: a box {8, box_outer} lands as box_outer rows of 16 bytes = the 8x16-byte core
// matrices of the interleaved (no-swizzle) K-major UMMA layout stacked along M/N.
int make_tmap_2d_bf16_nosw(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                           uint32_t box_inner, uint32_t box_outer);

}  // namespace ark
