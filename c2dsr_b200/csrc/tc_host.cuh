// Host-side helpers shared by the tcgen05 kernels: TMA descriptor creation, launch, bf16 split.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"

namespace c2dsr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 matrix [rows, k] with leading dimension ld (elements): box = 64 (k) x box_rows, 128-byte swizzle,
// out-of-bounds elements read as zero (so ragged M, N and K need no padding)
static inline int make_bf16_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int64_t ld,
                                int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return C2DSR_ERR_ARCH;
    }
    if (((uintptr_t)base & 15) || (ld * 2) % 16) {
        set_error("TMA operand must be 16-byte aligned with a leading dimension that is a multiple of 8");
        return C2DSR_ERR_ARG;
    }
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)(rows > 0 ? rows : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return C2DSR_ERR_ARG;
    }
    return C2DSR_OK;
}

// One operand with `rows` M/N indices and inner extent K.  K-major: the matrix is [rows, K] (ld), one box of
// 64 k x box_rows.  MN-major: the matrix is [K, rows] (ld), boxes of 64 mn x 64 k.
static inline int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld,
                                   bool mn_major, int box_rows) {
    if (mn_major) return make_bf16_map(map, base, K, rows, ld, tc::BK);
    return make_bf16_map(map, base, rows, K, ld, box_rows);
}

// maps for A (M rows) and B (N rows), hi and (optionally) lo parts
template <int BN>
static inline int make_maps(tc::Maps* maps, const uint16_t* a_hi, const uint16_t* a_lo, int64_t M, int64_t lda,
                            const uint16_t* b_hi, const uint16_t* b_lo, int64_t N, int64_t ldb, int64_t K,
                            int passes, bool a_mn = false, bool b_mn = false) {
    int rc;
    if ((rc = make_operand_map(&maps->a_hi, a_hi, M, K, lda, a_mn, tc::BM))) return rc;
    if ((rc = make_operand_map(&maps->b_hi, b_hi, N, K, ldb, b_mn, BN))) return rc;
    if ((rc = make_operand_map(&maps->a_lo, passes == 3 ? a_lo : a_hi, M, K, lda, a_mn, tc::BM))) return rc;
    if ((rc = make_operand_map(&maps->b_lo, passes == 3 ? b_lo : b_hi, N, K, ldb, b_mn, BN))) return rc;
    return C2DSR_OK;
}

// The kernel gives every K slab ceil(n_kb / slabs) k-blocks; shrink the slab count so that none is empty.
static inline int effective_splits(int64_t K, int64_t slabs) {
    const int64_t n_kb = ceil_div(K, tc::BK);
    if (slabs < 1) slabs = 1;
    if (slabs > n_kb) slabs = n_kb;
    const int64_t per = ceil_div(n_kb, slabs);
    return (int)ceil_div(n_kb, per);
}

static inline int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN, int STAGES, bool ARES, bool A_MN, bool B_MN, class Epi>
static inline int launch_gemm(const tc::Maps& maps, const tc::Problem& pb, const Epi& epi, cudaStream_t st) {
    using L = tc::SmemLayout<BN, STAGES, ARES>;
    auto kern = tc::gemm_kernel<BN, STAGES, ARES, A_MN, B_MN, Epi>;
    if (ARES && pb.K > tc::ARES_MAX_KB * tc::BK) {
        set_error("resident-A GEMM needs K <= %d", tc::ARES_MAX_KB * tc::BK);
        return C2DSR_ERR_ARG;
    }
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        attr = true;
    }
    const int64_t m_blocks = ceil_div(pb.M, tc::BM), n_blocks = ceil_div(pb.N, BN);
    const int64_t tiles = pb.diag_only ? m_blocks : m_blocks * n_blocks * (pb.k_splits > 1 ? pb.k_splits : 1);
    if (tiles <= 0) return C2DSR_OK;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    kern<<<grid, tc::THREADS, L::TOTAL, st>>>(maps, pb, epi);
    note_launches(1);
    return check_launch("tc_gemm");
}

// fp32 -> bf16 round-to-nearest-even
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);   // inf / nan
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
__device__ __forceinline__ float bf16_to_f32(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }
// x = hi + lo, hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split2(float x, uint16_t& hi, uint16_t& lo) {
    hi = f32_to_bf16_rn(x);
    lo = f32_to_bf16_rn(x - bf16_to_f32(hi));
}

// launchers defined in score_tc.cu
int split_rows(const float* X, int64_t rows, int d, int64_t ld_out, uint16_t* hi, uint16_t* lo, cudaStream_t st);
// XT[c, r] = X[r, c]: hi/lo [d, ld_out] with ld_out >= rows
int split_rows_transposed(const float* X, int64_t rows, int d, int64_t ld_out, uint16_t* hi, uint16_t* lo,
                          cudaStream_t st);

}  // namespace c2dsr
