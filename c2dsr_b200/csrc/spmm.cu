// K2: CSR SpMM for the parameter-free GCN  out = alpha * A (.) X + beta * Y + gamma * Z.
//
// One warp per output row.  The row's (col, val) pairs are read once, 32 at a time, by the lanes
// (coalesced) and handed round with shuffles; every lane then streams its float4 slice of each
// gathered X row, so a row of d floats is one or two 512-byte coalesced requests per warp.  Rows
// average ~4.5 non-zeros, but item popularity is heavy-tailed: rows longer than kLongRow non-zeros
// are skipped by the warp-per-row kernel and handled by a second kernel, one 32-warp CTA per long
// row: each warp sums a contiguous slice of the row's non-zeros and the 32 partials are added in warp
// order, so the result has a fixed summation order.  HBM-bound: algorithmic bytes per row are
// nnz_row * (8 + 4d) + 4 (rowptr) + 4d per addend + 4d written.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

constexpr int kLongRow = 256;

struct SpmmArgs {
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* X;
    const float* Y;
    const float* Z;
    float* out;
    int64_t n_rows;
    int d;
    float alpha, beta, gamma;
    int drop_mode;
    Dropout dr;
};

template <int VPL>
__device__ __forceinline__ void accumulate_range(const SpmmArgs& a, int beg, int end, int lane, int nv,
                                                 float4 (&acc)[VPL]) {
    for (int base = beg; base < end; base += 32) {
        const int mine = base + lane;
        const int c_l = mine < end ? a.col[mine] : 0;
        const float v_l = mine < end ? a.val[mine] : 0.f;
        const int cnt = end - base < 32 ? end - base : 32;
#pragma unroll 4
        for (int e = 0; e < cnt; ++e) {
            const int c = __shfl_sync(0xffffffffu, c_l, e);
            const float v = __shfl_sync(0xffffffffu, v_l, e);
            const float4* x4 = reinterpret_cast<const float4*>(a.X + (int64_t)c * a.d);
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int vi = lane + 32 * k;
                if (vi < nv) {
                    float4 x = __ldg(x4 + vi);
                    if (a.drop_mode == 1) drop_scale4(a.dr, (uint64_t)c * a.d + 4 * vi, x);
                    acc[k].x += v * x.x;
                    acc[k].y += v * x.y;
                    acc[k].z += v * x.z;
                    acc[k].w += v * x.w;
                }
            }
        }
    }
}

__device__ __forceinline__ void finish_row(const SpmmArgs& a, int64_t row, int vi, float4 r) {
    if (a.drop_mode == 2) drop_scale4(a.dr, (uint64_t)row * a.d + 4 * vi, r);
    r.x *= a.alpha; r.y *= a.alpha; r.z *= a.alpha; r.w *= a.alpha;
    if (a.Y) {
        const float4 y = reinterpret_cast<const float4*>(a.Y + row * a.d)[vi];
        r.x += a.beta * y.x; r.y += a.beta * y.y; r.z += a.beta * y.z; r.w += a.beta * y.w;
    }
    if (a.Z) {
        const float4 z = reinterpret_cast<const float4*>(a.Z + row * a.d)[vi];
        r.x += a.gamma * z.x; r.y += a.gamma * z.y; r.z += a.gamma * z.z; r.w += a.gamma * z.w;
    }
    reinterpret_cast<float4*>(a.out + row * a.d)[vi] = r;
}

template <int VPL>   // float4 vectors per lane: d <= 128 * VPL
__global__ void __launch_bounds__(256) spmm_kernel(SpmmArgs a, bool skip_long) {
    const int lane = threadIdx.x & 31;
    const int nv = a.d >> 2;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    // persistent warps: rows are short (a few non-zeros), so each warp walks a strided set of rows instead of
    // paying a block launch per 8 rows
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < a.n_rows; row += n_warps) {
        const int beg = a.rowptr[row], end = a.rowptr[row + 1];
        if (skip_long && end - beg > kLongRow) continue;
        float4 acc[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate_range<VPL>(a, beg, end, lane, nv, acc);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int vi = lane + 32 * k;
            if (vi < nv) finish_row(a, row, vi, acc[k]);
        }
    }
}

// one CTA of 32 warps per long row; dynamic smem = 32 * d floats
template <int VPL>
__global__ void __launch_bounds__(1024) spmm_long_kernel(SpmmArgs a, const int32_t* __restrict__ long_rows) {
    extern __shared__ float part[];                 // [32 warps][d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = long_rows[blockIdx.x];
    const int beg = a.rowptr[row], end = a.rowptr[row + 1];
    const int per = (end - beg + 31) / 32;
    const int w_beg = beg + warp * per < end ? beg + warp * per : end;
    const int w_end = w_beg + per < end ? w_beg + per : end;
    const int nv = a.d >> 2;
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    accumulate_range<VPL>(a, w_beg, w_end, lane, nv, acc);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int vi = lane + 32 * k;
        if (vi < nv) reinterpret_cast<float4*>(part + warp * a.d)[vi] = acc[k];
    }
    __syncthreads();
    for (int vi = threadIdx.x; vi < nv; vi += blockDim.x) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < 32; ++w) {
            const float4 p = reinterpret_cast<const float4*>(part + w * a.d)[vi];
            r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w;
        }
        finish_row(a, row, vi, r);
    }
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" int c2dsr_spmm(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* long_rows,
                          int n_long, const float* X, const float* Y, const float* Z, float* out, int64_t n_rows,
                          int d, float alpha, float beta, float gamma, int drop_mode, float p, uint64_t seed,
                          uint64_t tag, void* stream) {
    if (n_rows <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 1024, "d must be a multiple of 4 in (0, 1024]");
    C2DSR_REQUIRE(drop_mode >= 0 && drop_mode <= 2, "drop_mode must be 0, 1 or 2");
    C2DSR_REQUIRE(X != out, "spmm cannot run in place on X");
    SpmmArgs a{rowptr, col, val, X, Y, Z, out, n_rows, d, alpha, beta, gamma, drop_mode, make_dropout(p, seed, tag)};
    if (a.dr.p == 0.f) a.drop_mode = 0;
    const unsigned blocks = (unsigned)ceil_div(n_rows, 8);     // one warp per row (a persistent, strided variant measured 10 % slower)
    cudaStream_t st = (cudaStream_t)stream;
    const bool split = long_rows != nullptr && n_long > 0;
    const int smem = 32 * d * 4;
#define LAUNCH(V)                                                                                        \
    do {                                                                                                 \
        spmm_kernel<V><<<blocks, 256, 0, st>>>(a, split);                                                \
        if (split) {                                                                                     \
            static bool attr = false;                                                                    \
            if (!attr) {                                                                                 \
                cudaFuncSetAttribute(spmm_long_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4); \
                attr = true;                                                                             \
            }                                                                                            \
            spmm_long_kernel<V><<<(unsigned)n_long, 1024, smem, st>>>(a, long_rows);                     \
        }                                                                                                \
    } while (0)
    if (d <= 128) LAUNCH(1);
    else if (d <= 256) LAUNCH(2);
    else if (d <= 512) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    note_launches(split ? 2 : 1);
    return check_launch("spmm");
}

/* rows with more than this many non-zeros should be listed in long_rows */
extern "C" int c2dsr_spmm_long_row_threshold(void) { return kLongRow; }
