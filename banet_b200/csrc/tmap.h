// Host-side creation of TMA tensor maps through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace banet {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

// fp32 matrix [rows, cols] row-major (cols contiguous); box = box_cols x box_rows, 128B swizzle with 32-byte atoms
// (the layout tcgen05 kind::tf32 needs for MN-major operands); box_cols*4 must be 128.
inline int make_tmap_f32_2d_sw128_32b(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols)
{
    PFN_encodeTiled enc = get_encode_tiled();
    BANET_REQUIRE(enc, BANET_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point not available");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BANET_REQUIRE(r == CUDA_SUCCESS, BANET_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols);
    return BANET_OK;
}

// fp32 tensor [planes, rows, cols] (cols contiguous, dense): box = box_cols x box_rows x box_planes, same swizzle.
// Used for the basis of a dense pixel grid: cols = K, rows = grid_w (x), planes = nb*grid_h (y); an 8x8 pixel tile is one box.
inline int make_tmap_f32_3d_sw128_32b(CUtensorMap* out, const float* base, uint64_t planes, uint64_t rows, uint64_t cols,
                                      uint32_t box_planes, uint32_t box_rows, uint32_t box_cols)
{
    PFN_encodeTiled enc = get_encode_tiled();
    BANET_REQUIRE(enc, BANET_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point not available");
    cuuint64_t gdim[3] = {cols, rows, planes};
    cuuint64_t gstr[2] = {cols * sizeof(float), rows * cols * sizeof(float)};
    cuuint32_t box[3] = {box_cols, box_rows, box_planes};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BANET_REQUIRE(r == CUDA_SUCCESS, BANET_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
    return BANET_OK;
}

// fp32 NHWC map [nb, h, w, c], no swizzle: box = box_c x box_w x box_h x 1 (staged F2 windows: box_c = 32 channels = 128 B rows).
inline int make_tmap_f32_nhwc(CUtensorMap* out, const float* base, uint64_t nb, uint64_t h, uint64_t w, uint64_t c,
                                       uint32_t box_c, uint32_t box_w, uint32_t box_h)
{
    PFN_encodeTiled enc = get_encode_tiled();
    BANET_REQUIRE(enc, BANET_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point not available");
    cuuint64_t gdim[4] = {c, w, h, nb};
    cuuint64_t gstr[3] = {c * sizeof(float), w * c * sizeof(float), h * w * c * sizeof(float)};
    cuuint32_t box[4] = {box_c, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BANET_REQUIRE(r == CUDA_SUCCESS, BANET_ERR_CUDA, "cuTensorMapEncodeTiled(nhwc) failed (%d)", (int)r);
    return BANET_OK;
}

}  // namespace banet
