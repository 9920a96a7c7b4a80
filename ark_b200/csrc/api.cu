// Library-level state of libarkb200: thread-local error text, launch counter, TMA descriptor encoder lookup.
#include "common.cuh"
#include "tmap.cuh"
#include <mutex>

extern "C" char** environ;
#include <string.h>
namespace ark {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
bool under_profiler() {
  static int v = -1;
  if (v < 0) {
    v = 0;
    for (char** e = environ; e && *e; ++e)
      if (!strncmp(*e, "NV_NSIGHT", 9) || !strncmp(*e, "NV_COMPUTE_PROFILER", 19) || !strncmp(*e, "CUDA_INJECTION64_PATH", 21) ||
          !strncmp(*e, "NVTX_INJECTION64_PATH", 21))
        v = 1;
  }
  return v != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(ARK_E_NODRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstride[1] = {ld * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ARK_E_SHAPE, "cuTensorMapEncodeTiled failed (CUresult %d) dims=(%llu,%llu) ld=%llu box=(%u,%u)", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
  return 0;
}

int make_tmap_2d_bf16_nosw(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                           uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(ARK_E_NODRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstride[1] = {ld * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ARK_E_SHAPE, "cuTensorMapEncodeTiled (no swizzle) failed (CUresult %d) dims=(%llu,%llu) ld=%llu box=(%u,%u)",
                (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
  return 0;
}

int make_tmap_kchunked_bf16(CUtensorMap* out, const void* base, uint64_t K, uint64_t rows, uint64_t ld,
                            uint32_t box_rows, uint32_t box_chunks) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(ARK_E_NODRIVER, "cuTensorMapEncodeTiled not available from the CUDA driver");
  if (K % 64 != 0 || box_rows == 0 || box_rows > 256 || box_chunks == 0 || box_chunks > 256)
    return fail(ARK_E_SHAPE, "k-chunked tensor map: K=%llu rows-box=%u chunk-box=%u", (unsigned long long)K, box_rows, box_chunks);
  const cuuint64_t gdim[3] = {64, rows, K / 64};
  const cuuint64_t gstride[2] = {ld * 2, 128};
  const cuuint32_t box[3] = {64, box_rows, box_chunks};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ARK_E_SHAPE, "cuTensorMapEncodeTiled (k-chunked) failed (CUresult %d) K=%llu rows=%llu ld=%llu box=(%u,%u)",
                (int)r, (unsigned long long)K, (unsigned long long)rows, (unsigned long long)ld, box_rows, box_chunks);
  return 0;
}

}  // namespace ark

extern "C" int ark_abi_version(void) { return ARK_ABI_VERSION; }
extern "C" const char* ark_last_error(void) { return ark::g_err; }
extern "C" int64_t ark_launch_count(void) { return ark::g_launches; }
extern "C" void ark_launch_count_reset(void) { ark::g_launches = 0; }
