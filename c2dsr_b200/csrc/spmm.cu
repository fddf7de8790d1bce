// K2: CSR SpMM for the parameter-free GCN  out = alpha * A (.) X + beta * Y + gamma * Z.
//
// One warp per output row.  The row's (col, val) pairs are read once, 32 at a time, by the lanes
// (coalesced) and handed round with shuffles; every lane then streams its float4 slice of each
// gathered X row, so a row of d floats is one or two 512-byte coalesced requests per warp.  Rows
// average ~4.5 non-zeros with a tail of ~100, so a warp per row keeps the tail bounded while the grid
// (N / 8 blocks of 8 warps) is tens of waves over 148 SMs.  HBM-bound: algorithmic bytes per row are
// nnz_row * (8 + 4d) + 4 (rowptr) + 4d per addend + 4d written.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

template <int VPL>   // float4 vectors per lane: d <= 128 * VPL
__global__ void __launch_bounds__(256)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
            const float* __restrict__ X, const float* Y, const float* Z,
            float* out, int64_t n_rows, int d, float alpha, float beta, float gamma, int drop_mode,
            Dropout dr) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int nv = d >> 2;
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = rowptr[row], end = rowptr[row + 1];
    for (int base = beg; base < end; base += 32) {
        const int mine = base + lane;
        const int c_l = mine < end ? col[mine] : 0;
        const float v_l = mine < end ? val[mine] : 0.f;
        const int cnt = end - base < 32 ? end - base : 32;
        for (int e = 0; e < cnt; ++e) {
            const int c = __shfl_sync(0xffffffffu, c_l, e);
            const float v = __shfl_sync(0xffffffffu, v_l, e);
            const float4* x4 = reinterpret_cast<const float4*>(X + (int64_t)c * d);
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int vi = lane + 32 * k;
                if (vi < nv) {
                    float4 x = __ldg(x4 + vi);
                    if (drop_mode == 1) {
                        const uint64_t b = (uint64_t)c * d + 4 * vi;
                        x.x *= drop_scale(dr, b);
                        x.y *= drop_scale(dr, b + 1);
                        x.z *= drop_scale(dr, b + 2);
                        x.w *= drop_scale(dr, b + 3);
                    }
                    acc[k].x += v * x.x;
                    acc[k].y += v * x.y;
                    acc[k].z += v * x.z;
                    acc[k].w += v * x.w;
                }
            }
        }
    }
    const float4* y4 = Y ? reinterpret_cast<const float4*>(Y + row * d) : nullptr;
    const float4* z4 = Z ? reinterpret_cast<const float4*>(Z + row * d) : nullptr;
    float4* o4 = reinterpret_cast<float4*>(out + row * d);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int vi = lane + 32 * k;
        if (vi < nv) {
            float4 r = acc[k];
            if (drop_mode == 2) {
                const uint64_t b = (uint64_t)row * d + 4 * vi;
                r.x *= drop_scale(dr, b);
                r.y *= drop_scale(dr, b + 1);
                r.z *= drop_scale(dr, b + 2);
                r.w *= drop_scale(dr, b + 3);
            }
            r.x *= alpha; r.y *= alpha; r.z *= alpha; r.w *= alpha;
            if (y4) {
                const float4 y = y4[vi];
                r.x += beta * y.x; r.y += beta * y.y; r.z += beta * y.z; r.w += beta * y.w;
            }
            if (z4) {
                const float4 z = z4[vi];
                r.x += gamma * z.x; r.y += gamma * z.y; r.z += gamma * z.z; r.w += gamma * z.w;
            }
            o4[vi] = r;
        }
    }
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" int c2dsr_spmm(const int32_t* rowptr, const int32_t* col, const float* val, const float* X,
                          const float* Y, const float* Z, float* out, int64_t n_rows, int d, float alpha,
                          float beta, float gamma, int drop_mode, float p, uint64_t seed, uint64_t tag,
                          void* stream) {
    if (n_rows <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 1024, "d must be a multiple of 4 in (0, 1024]");
    C2DSR_REQUIRE(drop_mode >= 0 && drop_mode <= 2, "drop_mode must be 0, 1 or 2");
    Dropout dr = make_dropout(p, seed, tag);
    if (dr.p == 0.f) drop_mode = 0;
    const unsigned blocks = (unsigned)ceil_div(n_rows, 8);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(V) spmm_kernel<V><<<blocks, 256, 0, st>>>(rowptr, col, val, X, Y, Z, out, n_rows, d, alpha, beta, \
                                                         gamma, drop_mode, dr)
    if (d <= 128) LAUNCH(1);
    else if (d <= 256) LAUNCH(2);
    else if (d <= 512) LAUNCH(4);
    else LAUNCH(8);
#undef LAUNCH
    note_launches(1);
    return check_launch("spmm");
}
