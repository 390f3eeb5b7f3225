// Host helpers shared by every kernel translation unit: the cuTensorMapEncodeTiled entry point (resolved through the
// runtime, so the library does not link libcuda) and the programmatic-dependent-launch switch.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace loco {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
}  // namespace

int tensormap_init() {
    if (g_encode != nullptr) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return (int)e;
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return (int)cudaErrorNotSupported;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// 2-D bf16 tensor map, box = [64 elements (128 B), box_rows], SWIZZLE_128B.
int make_tensor_map_bf16_sw128(void* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                               uint32_t box_rows) {
    if (g_encode == nullptr) return (int)cudaErrorNotReady;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {64u, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

bool pdl_enabled() {
    static const bool on = []() {
        const char* e = getenv("LOCO_PDL");      // opt-in: measured 1.7 % SLOWER on the SLURP-shaped bench (24.15 -> 24.58 ms/step)
        return e && e[0] == '1';
    }();
    return on;
}

}  // namespace loco
