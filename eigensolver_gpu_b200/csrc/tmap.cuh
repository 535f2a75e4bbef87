// eigb200 -- host-side construction of a 2-D TMA tensor map over a column-major FP64 matrix (driver entry point
// fetched through the runtime, so libcuda is not linked).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eigb200 {

// Describes a column-major matrix of `rows_d` doubles per column (complex: 2 per element), `cols` columns, column
// stride `ld_bytes`; the box is 16 doubles (128 bytes, one swizzle span) x 64 columns, SWIZZLE_128B, out-of-bounds
// elements read as zero.  Returns 0 on success.
inline int make_tmap_f64_box16x64(CUtensorMap* out, const void* base, uint64_t rows_d, uint64_t cols, uint64_t ld_bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  if (!fn) return -1;
  if (((uintptr_t)base & 15) || (ld_bytes & 15) || rows_d == 0 || cols == 0) return -1;
  cuuint64_t gdim[2] = {rows_d, cols};
  cuuint64_t gstr[1] = {ld_bytes};
  cuuint32_t box[2] = {16, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace eigb200
