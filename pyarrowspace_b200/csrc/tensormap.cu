// tensormap.cu -- TMA descriptors for the item matrix.
//
// The item matrix X is n x fp f64 row-major (fp = features rounded up to a multiple of 4, zero
// padded).  Both tensor-core kernels want their shared-memory operand stage as
//     stage[outer][row][4]      (outer = feature/4, 4 = the inner 32 bytes of a feature quad)
// because then a DMMA fragment load (lane l -> row l/4, k l%4 for the search GEMM; lane l ->
// feature l/4, item l%4 for the Gram) touches 256 contiguous bytes per warp: conflict free without
// padding or swizzle.  TMA produces exactly that layout from ONE descriptor that views X as a 3-D
// tensor (4, n, fp/4) with strides (8 B, fp*8 B, 32 B): the box (4, rows, outer) lands in shared
// memory innermost-first.
#include "common.cuh"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int asp_make_items_tmap(CUtensorMap *out, const double *base, int64_t rows, int32_t fp, int box_rows, int box_outer)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) ASP_FAIL(ASP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if (fp % 4 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0)
        ASP_FAIL(ASP_ERR_ARG, "item matrix must be 16-byte aligned with a pitch that is a multiple of 4 doubles");
    cuuint64_t dims[3] = {4, (cuuint64_t)rows, (cuuint64_t)(fp / 4)};
    cuuint64_t strides[2] = {(cuuint64_t)fp * 8, 32};          // bytes, for dims 1 and 2
    cuuint32_t box[3] = {4, (cuuint32_t)box_rows, (cuuint32_t)box_outer};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ASP_FAIL(ASP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ASP_OK;
}

// fp16 operand of the tcgen05 path: rows x kp (kp a multiple of 64), K-major.  Box = 64 elements (128 B = one
// swizzle atom) x box_rows, SWIZZLE_128B: exactly the canonical K-major layout the UMMA shared-memory descriptor
// (layout type SWIZZLE_128B, SBO = 1024 B) expects.
int asp_make_f16_tmap(CUtensorMap *out, const void *base, int64_t rows, int32_t kp, int box_rows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) ASP_FAIL(ASP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if (kp % 64 != 0) ASP_FAIL(ASP_ERR_ARG, "fp16 operand pitch must be a multiple of 64");
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ASP_FAIL(ASP_ERR_CUDA, "cuTensorMapEncodeTiled (fp16) failed with CUresult %d", (int)r);
    return ASP_OK;
}

// Narrow box over the same operand: `nw` = 16 or 32 elements (32 / 64 bytes) x box_rows, SWIZZLE_32B / SWIZZLE_64B: the
// canonical K-major layout of a 32- / 64-byte-row tile (UMMA layout types 6 / 4).  For the last k block of the tcgen05
// candidate pass, which holds only the 3 rank-1 columns (+ padding) when the feature count is a multiple of 64.
int asp_make_f16_tmap_narrow(CUtensorMap *out, const void *base, int64_t rows, int32_t kp, int box_rows, int nw)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) ASP_FAIL(ASP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if (kp % 64 != 0 || (nw != 16 && nw != 32)) ASP_FAIL(ASP_ERR_ARG, "narrow fp16 box must be 16 or 32 elements wide");
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)nw, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, nw == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ASP_FAIL(ASP_ERR_CUDA, "cuTensorMapEncodeTiled (narrow fp16) failed with CUresult %d", (int)r);
    return ASP_OK;
}
