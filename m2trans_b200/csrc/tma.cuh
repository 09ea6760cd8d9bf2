// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled) without linking libcuda: the driver
// entry point is fetched through the runtime.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace m2t {

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

// dims / strides_bytes / box: innermost dimension first; strides_bytes[0] is implied by elem_bytes and ignored.
// swizzle: 0 none, 1 32B, 2 64B, 3 128B.  Out-of-bounds elements are filled with zeros.
inline int make_tensor_map(CUtensorMap* map, const void* gaddr, int elem_bytes, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
    static PFN_tensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        M2T_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled is not available"); return M2T_E_CUDA; }
        encode = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn);
    }
    CUtensorMapDataType dt;
    switch (elem_bytes) {
        case 1: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
        case 2: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; break;
        case 4: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
        default: set_error("tensor map: element size %d", elem_bytes); return M2T_E_ARG;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 1; i < rank; ++i) gstr[i - 1] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = encode(map, dt, (cuuint32_t)rank, const_cast<void*>(gaddr), gdim, gstr, bx, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return M2T_E_CUDA; }
    return M2T_OK;
}

}  // namespace m2t
