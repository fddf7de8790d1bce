// K4b on tensor cores: full-catalogue scoring fused with the rank count (trainer.py:168-179 of the
// reference).  scores = Q W^T + b are produced tile by tile in tensor memory by tcgen05 MMAs fed by TMA
// and consumed in the epilogue (compare with the target score, count) -- the [n_q, N] score matrix
// never reaches HBM.  The target score itself comes from the same MMA sequence applied to the gathered
// target rows, so "s_j > s_gt" compares numbers produced by identical arithmetic.
#include "common.cuh"
#include "tc_gemm.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

// fp32 -> (hi, lo) bf16 split: hi = bf16(x), lo = bf16(x - hi).
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);   // inf / nan
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

__global__ void split_bf16_kernel(const float* __restrict__ X, int64_t rows, int d, int64_t ld_out,
                                  uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int64_t total = rows * (int64_t)d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int c = (int)(i % d);
        const float x = X[i];
        const uint16_t h = f32_to_bf16_rn(x);
        const float hf = __uint_as_float((uint32_t)h << 16);
        hi[r * ld_out + c] = h;
        if (lo) lo[r * ld_out + c] = f32_to_bf16_rn(x - hf);
    }
}

// G[i, :] = W[gt[i] - n0, :] for targets inside the shard, zeros otherwise; bias_gt[i] likewise.
__global__ void gather_target_rows_kernel(const uint16_t* __restrict__ W_hi, const uint16_t* __restrict__ W_lo,
                                          const float* __restrict__ bias, const int64_t* __restrict__ gt, int64_t n_q,
                                          int64_t n0, int64_t n1, int d, uint16_t* __restrict__ G_hi,
                                          uint16_t* __restrict__ G_lo, float* __restrict__ bias_gt) {
    const int64_t i = blockIdx.x;
    const int64_t g = gt[i];
    const bool own = g >= n0 && g < n1;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        G_hi[i * d + c] = own ? W_hi[(g - n0) * d + c] : (uint16_t)0;
        if (G_lo) G_lo[i * d + c] = own ? W_lo[(g - n0) * d + c] : (uint16_t)0;
    }
    if (threadIdx.x == 0) bias_gt[i] = own ? bias[g - n0] : 0.f;
}

// ---- epilogues -----------------------------------------------------------------------------------
struct CountEpilogue {
    const float* bias;       // [N] shard-local
    const float* s_gt;       // [M]
    const int64_t* gt;       // [M] domain-local ids
    int32_t* counts;         // [M]
    float* S_debug;          // optional [M, lds]
    int64_t lds, M, N, n0;
    float tgt;
    int64_t g_local;
    int cnt;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t, int64_t row) {
        cnt = 0;
        if (row < M) {
            tgt = s_gt[row];
            g_local = gt[row] - n0;
        }
    }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M) return;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int64_t c = col0 + i;
            if (c < N) {
                const float s = v[i] + __ldg(bias + c);
                cnt += (c != g_local) && (s > tgt);
                if (S_debug) S_debug[row * lds + c] = s;
            }
        }
    }
    __device__ __forceinline__ void tile_end(int64_t row) {
        if (row < M && cnt) atomicAdd(counts + row, cnt);     // integer: order independent
    }
};

struct DiagEpilogue {       // BN == BM: element (row, row) of tile (b, b)
    const float* bias_gt;    // [M]
    float* s_gt;             // [M]
    int64_t M;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t, int64_t) {}
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M || row < col0 || row >= col0 + 32) return;
        const int want = (int)(row - col0);
        float x = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i == want) x = v[i];
        s_gt[row] = x + bias_gt[row];
    }
    __device__ __forceinline__ void tile_end(int64_t) {}
};

// ---- host helpers --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// bf16 matrix [rows, k] with leading dimension ld (elements): box = 64 (k) x box_rows, 128-byte swizzle
int make_bf16_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
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

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

constexpr int kBN = 128, kStages = 3;

template <class Epi>
static int launch_gemm(const tc::Maps& maps, const tc::Problem& pb, const Epi& epi, cudaStream_t st) {
    using L = tc::SmemLayout<kBN, kStages>;
    auto kern = tc::gemm_kernel<kBN, kStages, Epi>;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        attr = true;
    }
    const int64_t m_blocks = ceil_div(pb.M, tc::BM), n_blocks = ceil_div(pb.N, kBN);
    const int64_t tiles = pb.diag_only ? m_blocks : m_blocks * n_blocks;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    kern<<<grid, 256, L::TOTAL, st>>>(maps, pb, epi);
    note_launches(1);
    return check_launch("tc_gemm");
}

}  // namespace c2dsr

using namespace c2dsr;

#define RUN(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_) return rc_;   \
    } while (0)

extern "C" {

int c2dsr_split_bf16(const float* X, int64_t rows, int d, int64_t ld_out, uint16_t* hi, uint16_t* lo,
                     void* stream) {
    if (rows <= 0) return C2DSR_OK;
    int64_t blocks = ceil_div(rows * (int64_t)d, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, rows, d, ld_out, hi, lo);
    note_launches(1);
    return check_launch("split_bf16");
}

int64_t c2dsr_score_tc_workspace_bytes(int64_t n_q, int64_t n_shard, int d) {
    (void)n_shard;
    const int64_t rows = align_up(n_q > 0 ? n_q : 1, tc::BM);
    return 2 * rows * (int64_t)d * 2 + rows * 4 + 1024;      // gathered target rows (hi, lo) + their bias
}

int c2dsr_score_target_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const uint16_t* W_hi, const uint16_t* W_lo,
                          const float* bias, const int64_t* gt, int64_t n_q, int64_t n0, int64_t n1, int d,
                          int passes, float* s_gt, void* workspace, int64_t workspace_bytes, void* stream) {
    if (n_q <= 0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    if (workspace_bytes < c2dsr_score_tc_workspace_bytes(n_q, n1 - n0, d)) {
        set_error("score_target_tc: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = align_up(n_q, tc::BM);
    uint16_t* G_hi = (uint16_t*)workspace;
    uint16_t* G_lo = G_hi + rows * d;
    float* bias_gt = (float*)(G_lo + rows * d);
    gather_target_rows_kernel<<<(unsigned)n_q, 128, 0, st>>>(W_hi, passes == 3 ? W_lo : nullptr, bias, gt, n_q, n0, n1,
                                                             d, G_hi, passes == 3 ? G_lo : nullptr, bias_gt);
    note_launches(1);
    tc::Maps maps;
    RUN(make_bf16_map(&maps.a_hi, Q_hi, n_q, d, d, tc::BM));
    RUN(make_bf16_map(&maps.b_hi, G_hi, n_q, d, d, kBN));
    RUN(make_bf16_map(&maps.a_lo, passes == 3 ? Q_lo : Q_hi, n_q, d, d, tc::BM));
    RUN(make_bf16_map(&maps.b_lo, passes == 3 ? G_lo : G_hi, n_q, d, d, kBN));
    tc::Problem pb{n_q, n_q, d, passes, 1};
    DiagEpilogue epi{bias_gt, s_gt, n_q};
    return launch_gemm(maps, pb, epi, st);
}

int c2dsr_score_count_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const uint16_t* W_hi, const uint16_t* W_lo,
                         const float* bias, const float* s_gt, const int64_t* gt, int64_t n_q, int64_t n0,
                         int64_t n1, int d, int passes, int32_t* counts, float* S_debug, int64_t lds,
                         void* workspace, int64_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    if (n_q <= 0 || n1 <= n0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    const int64_t n = n1 - n0;
    tc::Maps maps;
    RUN(make_bf16_map(&maps.a_hi, Q_hi, n_q, d, d, tc::BM));
    RUN(make_bf16_map(&maps.b_hi, W_hi, n, d, d, kBN));
    RUN(make_bf16_map(&maps.a_lo, passes == 3 ? Q_lo : Q_hi, n_q, d, d, tc::BM));
    RUN(make_bf16_map(&maps.b_lo, passes == 3 ? W_lo : W_hi, n, d, d, kBN));
    tc::Problem pb{n_q, n, d, passes, 0};
    CountEpilogue epi{bias, s_gt, gt, counts, S_debug, lds, n_q, n, n0, 0.f, 0, 0};
    return launch_gemm(maps, pb, epi, (cudaStream_t)stream);
}

}  // extern "C"
