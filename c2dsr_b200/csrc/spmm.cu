// K2: CSR SpMM for the parameter-free GCN  out = alpha * A (.) X + beta * Y + gamma * Z.
//
// A warp owns a few consecutive output rows.  Their (col, val) pairs are read once, 32 at a time, by
// the lanes (coalesced) and handed round with shuffles; every lane then streams its float4 slice of each
// gathered X row, so a row of d floats is one or two 512-byte coalesced requests per warp.  Rows
// average ~4.5 non-zeros, but item popularity is heavy-tailed: rows longer than kLongRow non-zeros
// are skipped by the row-group kernel and handled by a second kernel, one 32-warp CTA per long
// row: each warp sums a contiguous slice of the row's non-zeros and the 32 partials are added in warp
// order, so the result has a fixed summation order.  HBM-bound: algorithmic bytes per row are
// nnz_row * (8 + 4d) + 4 (rowptr) + 4d per addend + 4d written.
#include <stdlib.h>
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
    const uint8_t* out_need;   // optional [n_rows]: 0 = this output row is not wanted (left zero)
    const uint8_t* x_nz;       // optional [n_cols]: 0 = this row of X is all zeros (its entries are skipped)
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

// RPW consecutive rows per warp.  The rows' non-zeros are one contiguous CSR range: the lanes read it
// 32 entries at a time (coalesced), and the gathers of U entries are issued together, across row
// boundaries, before any of them is consumed -- so a warp keeps U * VPL 512-byte requests in flight
// and pays the rowptr -> (col, val) -> X dependency chain once per RPW rows instead of once per row
// (rows average 4-5 non-zeros: the per-row chain, not bandwidth, bounded the one-row-per-warp form).
// The accumulator is flushed whenever the flat entry index crosses a row end; every row is still
// summed by one warp in CSR order, so results do not depend on the grouping.
template <int VPL, int RPW, int U>   // float4 vectors per lane: d <= 128 * VPL
__global__ void __launch_bounds__(256) spmm_kernel(SpmmArgs a, bool skip_long) {
    const int lane = threadIdx.x & 31;
    const int nv = a.d >> 2;
    const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    if (r0 >= a.n_rows) return;
    const int nr = a.n_rows - r0 < RPW ? (int)(a.n_rows - r0) : RPW;
    const int rp = lane <= nr ? a.rowptr[r0 + lane] : 0;
    const int lo = __shfl_sync(0xffffffffu, rp, 0), hi = __shfl_sync(0xffffffffu, rp, nr);
    int cur = 0, row_beg = lo, row_end = __shfl_sync(0xffffffffu, rp, 1);
    bool skip = skip_long && row_end - row_beg > kLongRow;
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&]() {
        if (!skip) {
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int vi = lane + 32 * k;
                if (vi < nv) finish_row(a, r0 + cur, vi, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        ++cur;
        row_beg = row_end;
        row_end = __shfl_sync(0xffffffffu, rp, cur + 1 <= nr ? cur + 1 : nr);
        skip = skip_long && row_end - row_beg > kLongRow;
    };
    for (int base = lo; base < hi; base += 32) {
        const int mine = base + lane;
        const int c_l = mine < hi ? a.col[mine] : 0;
        const float v_l = mine < hi ? a.val[mine] : 0.f;
        const int cnt = hi - base < 32 ? hi - base : 32;
        for (int e0 = 0; e0 < cnt; e0 += U) {
            float4 x[U][VPL];
            int c[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                c[u] = __shfl_sync(0xffffffffu, c_l, (e0 + u) & 31);
                if (e0 + u < cnt) {
                    const float4* x4 = reinterpret_cast<const float4*>(a.X + (int64_t)c[u] * a.d);
#pragma unroll
                    for (int k = 0; k < VPL; ++k) {
                        const int vi = lane + 32 * k;
                        if (vi < nv) x[u][k] = __ldg(x4 + vi);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float v = __shfl_sync(0xffffffffu, v_l, (e0 + u) & 31);
                if (e0 + u < cnt) {
                    while (base + e0 + u >= row_end) flush();          // warp-uniform
                    if (!skip) {
#pragma unroll
                        for (int k = 0; k < VPL; ++k) {
                            const int vi = lane + 32 * k;
                            if (vi < nv) {
                                float4 xx = x[u][k];
                                if (a.drop_mode == 1) drop_scale4(a.dr, (uint64_t)c[u] * a.d + 4 * vi, xx);
                                acc[k].x += v * xx.x;
                                acc[k].y += v * xx.y;
                                acc[k].z += v * xx.z;
                                acc[k].w += v * xx.w;
                            }
                        }
                    }
                }
            }
        }
    }
    while (cur < nr) flush();
}

// ---- bulk-copy (TMA) form ---------------------------------------------------------------------------
// Same row grouping, but the gathered X rows are fetched by the copy engine: lane e issues ONE
// cp.async.bulk (d * 4 bytes, global -> shared) for entry e of the current chunk, completion is counted on
// an mbarrier, and the warp accumulates the chunk from shared memory while the next chunk is already in
// flight (two stages per warp).  No registers are tied up by loads in flight and a 1 KB row costs one
// instruction instead of 32 lanes x 2 LDG.128.  The addend rows Y / Z of the group are contiguous: one bulk
// copy each, issued before the first gather, so that finishing a row never waits on a load (in the
// one-row-per-warp form 58 % of the stall samples sat on the Y load of finish_row).
#ifndef C2DSR_SPMM_WARPS
#define C2DSR_SPMM_WARPS 4
#endif
#ifndef C2DSR_SPMM_STAGES
#define C2DSR_SPMM_STAGES 2
#endif
#ifndef C2DSR_SPMM_STAGE_BYTES
#define C2DSR_SPMM_STAGE_BYTES 2048
#endif
#ifndef C2DSR_SPMM_RPW
#define C2DSR_SPMM_RPW 2
#endif
constexpr int kBulkWarps = C2DSR_SPMM_WARPS;              // warps per CTA
constexpr int kBulkStages = C2DSR_SPMM_STAGES;
constexpr int kBulkStageBytes = C2DSR_SPMM_STAGE_BYTES;   // per warp and stage: bytes / (4 d) gathered rows, at least one
// (measured at d = 256 on the FK graph, 4 warps: 2 stages x 2 KB, 2 rows per warp = 105 us; 4 KB stages / 4 rows
// = 139 us; 3 stages or 8 KB stages = 170-195 us: resident warps per SM matter more than depth per warp)
constexpr int kBulkBarBytes = (kBulkWarps * (kBulkStages + 1) * 8 + 127) / 128 * 128;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_addr(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SPMM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SPMM_DONE;\n\t"
        "bra SPMM_WAIT;\n\t"
        "SPMM_DONE:\n\t"
        "}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

template <int VPL, int RPW>
__global__ void __launch_bounds__(32 * kBulkWarps) spmm_bulk_kernel(SpmmArgs a, bool skip_long, int chunk) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + warp * (kBulkStages + 1);
    const uint32_t row_bytes = (uint32_t)a.d * 4u;
    const int n_add = (a.Y ? 1 : 0) + (a.Z ? 1 : 0);
    const size_t stage_bytes = (size_t)chunk * row_bytes;
    const size_t per_warp = (size_t)kBulkStages * stage_bytes + (size_t)n_add * RPW * row_bytes;
    float* stage0 = reinterpret_cast<float*>(smem_raw + kBulkBarBytes + warp * per_warp);
    float* ybuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage0) + kBulkStages * stage_bytes);
    float* zbuf = ybuf + (a.Y ? RPW * a.d : 0);
    const int nv = a.d >> 2;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s <= kBulkStages; ++s) bar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t r0 = ((int64_t)blockIdx.x * kBulkWarps + warp) * RPW;
    if (r0 >= a.n_rows) return;
    const int nr = a.n_rows - r0 < RPW ? (int)(a.n_rows - r0) : RPW;
    // rows of this group that are wanted (all of them without a mask); a group with none ends here: only the
    // rows a batch reads are propagated in a training step, ~6 % of the table at the Food-Kitchen shape
    uint32_t need = (1u << nr) - 1u;
    if (a.out_need) {
        need = 0;
#pragma unroll
        for (int r = 0; r < RPW; ++r)
            if (r < nr && a.out_need[r0 + r]) need |= 1u << r;
        if (need == 0) return;
    }
    const int rp = lane <= nr ? a.rowptr[r0 + lane] : 0;
    if (n_add && lane == 0) {                                  // addend rows of the whole group: contiguous
        bar_expect_tx(bars + kBulkStages, (uint32_t)(n_add * nr) * row_bytes);
        if (a.Y) bulk_copy_g2s(ybuf, a.Y + r0 * a.d, (uint32_t)nr * row_bytes, bars + kBulkStages);
        if (a.Z) bulk_copy_g2s(zbuf, a.Z + r0 * a.d, (uint32_t)nr * row_bytes, bars + kBulkStages);
    }
    bool add_ready = n_add == 0;
    const int lo = __shfl_sync(0xffffffffu, rp, 0), hi = __shfl_sync(0xffffffffu, rp, nr);
    int cur = 0, row_beg = lo, row_end = __shfl_sync(0xffffffffu, rp, 1);
    bool skip = (skip_long && row_end - row_beg > kLongRow) || !(need & 1u);
    float4 acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&]() {
        if (!add_ready) {
            bar_wait(bars + kBulkStages, 0);
            add_ready = true;
        }
        if (!skip) {
            const int64_t row = r0 + cur;
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int vi = lane + 32 * k;
                if (vi < nv) {
                    float4 r = acc[k];
                    if (a.drop_mode == 2) drop_scale4(a.dr, (uint64_t)row * a.d + 4 * vi, r);
                    r.x *= a.alpha; r.y *= a.alpha; r.z *= a.alpha; r.w *= a.alpha;
                    if (a.Y) {
                        const float4 y = reinterpret_cast<const float4*>(ybuf + cur * a.d)[vi];
                        r.x += a.beta * y.x; r.y += a.beta * y.y; r.z += a.beta * y.z; r.w += a.beta * y.w;
                    }
                    if (a.Z) {
                        const float4 z = reinterpret_cast<const float4*>(zbuf + cur * a.d)[vi];
                        r.x += a.gamma * z.x; r.y += a.gamma * z.y; r.z += a.gamma * z.z; r.w += a.gamma * z.w;
                    }
                    reinterpret_cast<float4*>(a.out + row * a.d)[vi] = r;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        ++cur;
        row_beg = row_end;
        row_end = __shfl_sync(0xffffffffu, rp, cur + 1 <= nr ? cur + 1 : nr);
        skip = (skip_long && row_end - row_beg > kLongRow) || !((need >> cur) & 1u);
    };
    const int n_chunks = (hi - lo + chunk - 1) / chunk;
    int c_s[kBulkStages];
    float v_s[kBulkStages];
    uint32_t a_s[kBulkStages];                           // per stage: which entries of the chunk were fetched
    uint32_t phase = 0;                                  // bit s = parity to wait for on stage s
    auto issue = [&](int k, int s) {
        const int beg = lo + k * chunk;
        const int cnt = hi - beg < chunk ? hi - beg : chunk;
        int c = 0;
        float v = 0.f;
        if (lane < cnt) {
            c = a.col[beg + lane];
            v = a.val[beg + lane];
        }
        // the row this lane's entry belongs to; entries of unwanted / long rows and of all-zero X rows are not fetched
        int er = 0;
#pragma unroll
        for (int r = 1; r < RPW; ++r) {
            const int b = __shfl_sync(0xffffffffu, rp, r);
            if (r < nr && beg + lane >= b) er = r;
        }
        const int e_beg = __shfl_sync(0xffffffffu, rp, er), e_end = __shfl_sync(0xffffffffu, rp, er + 1);
        bool act = lane < cnt && ((need >> er) & 1u) && !(skip_long && e_end - e_beg > kLongRow);
        if (act && a.x_nz) act = a.x_nz[c] != 0;
        const uint32_t mask = __ballot_sync(0xffffffffu, act);
        if (lane == 0) bar_expect_tx(bars + s, (uint32_t)__popc(mask) * row_bytes);
        __syncwarp();
        if (act)
            bulk_copy_g2s(reinterpret_cast<unsigned char*>(stage0) + (size_t)s * stage_bytes + (size_t)lane * row_bytes,
                          a.X + (int64_t)c * a.d, row_bytes, bars + s);
#pragma unroll
        for (int t = 0; t < kBulkStages; ++t)
            if (t == s) {
                c_s[t] = c;
                v_s[t] = v;
                a_s[t] = mask;
            }
    };
#pragma unroll
    for (int s = 0; s < kBulkStages; ++s)
        if (s < n_chunks) issue(s, s);
    for (int k = 0; k < n_chunks; ++k) {
        const int s = k % kBulkStages;
        bar_wait(bars + s, (phase >> s) & 1u);
        phase ^= 1u << s;
        const int beg = lo + k * chunk;
        const int cnt = hi - beg < chunk ? hi - beg : chunk;
        int c_l = 0;
        float v_l = 0.f;
        uint32_t act_l = 0;
#pragma unroll
        for (int t = 0; t < kBulkStages; ++t)
            if (t == s) {
                c_l = c_s[t];
                v_l = v_s[t];
                act_l = a_s[t];
            }
        const float4* st4 = reinterpret_cast<const float4*>(reinterpret_cast<unsigned char*>(stage0) + (size_t)s * stage_bytes);
        for (int e = 0; e < cnt; ++e) {
            while (beg + e >= row_end) flush();                      // warp-uniform
            const float v = __shfl_sync(0xffffffffu, v_l, e);
            const int c = __shfl_sync(0xffffffffu, c_l, e);
            if (!skip && ((act_l >> e) & 1u)) {
#pragma unroll
                for (int q = 0; q < VPL; ++q) {
                    const int vi = lane + 32 * q;
                    if (vi < nv) {
                        float4 xx = st4[e * nv + vi];
                        if (a.drop_mode == 1) drop_scale4(a.dr, (uint64_t)c * a.d + 4 * vi, xx);
                        acc[q].x += v * xx.x;
                        acc[q].y += v * xx.y;
                        acc[q].z += v * xx.z;
                        acc[q].w += v * xx.w;
                    }
                }
            }
        }
        __syncwarp();                                                // stage s fully read before it is refilled
        if (k + kBulkStages < n_chunks) issue(k + kBulkStages, s);
    }
    while (cur < nr) flush();
}

// one CTA of 32 warps per long row; dynamic smem = 32 * d floats
template <int VPL>
__global__ void __launch_bounds__(1024) spmm_long_kernel(SpmmArgs a, const int32_t* __restrict__ long_rows) {
    extern __shared__ float part[];                 // [32 warps][d]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = long_rows[blockIdx.x];
    if (a.out_need && !a.out_need[row]) return;               // not wanted: stays zero
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

__global__ void mark_rows_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t n_rows, uint8_t* __restrict__ mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t k = ids[i];
        if (k >= 0 && k < n_rows) mask[k] = 1;              // benign race: every writer stores the same value
    }
}

extern "C" int c2dsr_mark_rows(const int64_t* ids, int64_t n, int64_t n_rows, uint8_t* mask, void* stream) {
    C2DSR_REQUIRE(n_rows > 0 && n >= 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(mask, 0, n_rows, st);
    if (n > 0) {
        mark_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ids, n, n_rows, mask);
        note_launches(1);
    }
    return check_launch("mark_rows");
}

extern "C" int c2dsr_spmm(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* long_rows,
                          int n_long, const float* X, const float* Y, const float* Z, float* out, int64_t n_rows,
                          int d, float alpha, float beta, float gamma, int drop_mode, float p, uint64_t seed,
                          uint64_t tag, const uint8_t* out_row_needed, const uint8_t* x_row_nonzero, void* stream) {
    if (n_rows <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 1024, "d must be a multiple of 4 in (0, 1024]");
    C2DSR_REQUIRE(drop_mode >= 0 && drop_mode <= 2, "drop_mode must be 0, 1 or 2");
    C2DSR_REQUIRE(X != out, "spmm cannot run in place on X");
    SpmmArgs a{rowptr, col, val, X, Y, Z, out, n_rows, d, alpha, beta, gamma, drop_mode, make_dropout(p, seed, tag),
               out_row_needed, x_row_nonzero};
    C2DSR_REQUIRE(!out_row_needed || (out != Y && out != Z), "a row mask needs out distinct from the addends");
    if (out_row_needed) cudaMemsetAsync(out, 0, (size_t)n_rows * d * 4, (cudaStream_t)stream);   // unwanted rows = 0
    if (a.dr.p == 0.f) a.drop_mode = 0;
    constexpr int kRowsPerWarp = C2DSR_SPMM_RPW;
    const unsigned blocks = (unsigned)ceil_div(n_rows, 8 * kRowsPerWarp);
    cudaStream_t st = (cudaStream_t)stream;
    const bool split = long_rows != nullptr && n_long > 0;
    const int smem = 32 * d * 4;
    const int n_add = (Y ? 1 : 0) + (Z ? 1 : 0);
    int chunk = kBulkStageBytes / (4 * d) < 32 ? kBulkStageBytes / (4 * d) : 32;
    if (chunk < 1) chunk = 1;
    const int bulk_smem = kBulkBarBytes + kBulkWarps * (kBulkStages * chunk * d * 4 + n_add * kRowsPerWarp * d * 4);
    static const char* force = getenv("C2DSR_SPMM");            // "warp" / "bulk": benchmarking override
    const bool bulk = chunk >= 1 && (d % 4) == 0 && bulk_smem <= 200 * 1024 &&
                      ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(Z)) & 15) == 0 &&
                      !(force && force[0] == 'w');
    const unsigned bulk_blocks = (unsigned)ceil_div(n_rows, kBulkWarps * kRowsPerWarp);
#define LAUNCH(V)                                                                                        \
    do {                                                                                                 \
        if (bulk) {                                                                                      \
            static bool battr = false;                                                                   \
            if (!battr) {                                                                                \
                cudaFuncSetAttribute(spmm_bulk_kernel<V, kRowsPerWarp>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
                battr = true;                                                                            \
            }                                                                                            \
            spmm_bulk_kernel<V, kRowsPerWarp><<<bulk_blocks, 32 * kBulkWarps, bulk_smem, st>>>(a, split, chunk); \
        } else                                                                                           \
            spmm_kernel<V, kRowsPerWarp, (V <= 2 ? 8 / V : 2)><<<blocks, 256, 0, st>>>(a, split);        \
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
