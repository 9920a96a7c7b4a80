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

}  // namespace ark
