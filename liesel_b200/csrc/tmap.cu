// Host-side TMA descriptor builders shared by the tensor-core dense path (dense_tc.cu) and the tools
// library's tcgen05 probe: cuTensorMapEncodeTiled is fetched through the runtime's driver entry point
// so the library does not link libcuda.
#include <cuda.h>

#include "common.cuh"

namespace lslb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D fp32 tensor map over a row-major [rows x cols] matrix (row stride ld floats), box = {32 cols, box_rows},
// 128-byte swizzle
int make_tmap_f32(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld,
                  int box_rows, int kc) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      return LSLB_ERR_CUDA;
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LSLB_OK : LSLB_ERR_CUDA;
}

// 2-D bf16 tensor map over a row-major [rows x cols] matrix (row stride ld elements), box = {16 cols, box_rows},
// 32-byte swizzle
int make_tmap_bf16_k16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                       int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      return LSLB_ERR_CUDA;
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LSLB_OK : LSLB_ERR_CUDA;
}

int make_tmap_f32_k32(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld,
                      int box_rows) {
  return make_tmap_f32(map, base, rows, cols, ld, box_rows, 32);
}

}  // namespace lslb
