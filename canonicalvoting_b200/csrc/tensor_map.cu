// canonicalvoting_b200/csrc/tensor_map.cu -- host-side construction of the TMA descriptors the tcgen05 kernels use.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace cvb200 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the driver entry point is looked up once (libcuda is not linked: the library must load on a box without a driver)
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_map_2d(CUtensorMap *m, const float *base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols,
                uint32_t box_rows, int swizzle_atom_32b) {
    EncodeTiledFn fn = encode_tiled();
    CVB_REQUIRE(fn != nullptr, CVB200_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    const CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle_atom_32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVB_REQUIRE(rc == CUDA_SUCCESS, CVB200_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d; %llu x %llu, stride %llu, box %u x %u)",
                (int)rc, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_cols, box_rows);
    return 0;
}

}  // namespace cvb200
