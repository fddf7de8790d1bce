// K4a on tensor cores: classifier logits + cross-entropy for training (trainer.py:131-152 of the
// reference) without ever materialising the fp32 logits.
//
// forward   Z = H W^T + b tile by tile in tensor memory (tcgen05, bf16 hi/lo split, fp32 accumulate);
//           the epilogue keeps, per row and tile, an online (max, sum exp) pair and picks the target
//           logit; a small kernel combines the per-tile pairs (+ the pad logit) into lse and loss.
// backward  the same GEMM is recomputed; its epilogue turns each tile into dZ = (softmax - onehot) * coef
//           and stores it as bf16 hi/lo in both orientations ([M, N] and [N, M]); two more tcgen05 GEMMs
//           give dH = dZ W (K = N) and dW += dZ^T H (K = M); db is a row sum of dZ^T.
// All reductions have a fixed order (per-tile partials, ordered combines), so results are deterministic.
#include <cuda_bf16.h>
#include "tc_host.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

constexpr float kLog2e = 1.4426950408889634f;

struct LseEpilogue {
    const float* bias;        // [N]
    const int64_t* gt;        // [M], N = ignored
    float* pmax;              // [M, n_blocks]
    float* psum;              // [M, n_blocks]
    float* zgt;               // [M] target logit (written by the tile that holds it)
    int64_t M, N, n_blocks;
    float m_run, s_run;
    int64_t g, nb;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t n_blk, int64_t row, int, int part) {
        m_run = -INFINITY;
        s_run = 0.f;
        nb = n_blk * tc::EPI_PARTS + part;                   // one (max, sum) pair per row, tile and column part
        g = row < M ? gt[row] : -1;
    }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M) return;
        float z[32];
        float cm = -INFINITY;
        if (col0 + 32 <= N && (reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0) {
            const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                z[4 * j] = v[4 * j] + b.x; z[4 * j + 1] = v[4 * j + 1] + b.y;
                z[4 * j + 2] = v[4 * j + 2] + b.z; z[4 * j + 3] = v[4 * j + 3] + b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = col0 + i < N ? v[i] + __ldg(bias + col0 + i) : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) cm = fmaxf(cm, z[i]);
        if (g >= col0 && g < col0 + 32) {                    // target logit lives in this chunk
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i == g) zgt[row] = z[i];
        }
        if (cm == -INFINITY) return;                         // chunk entirely past N
        const float new_m = fmaxf(m_run, cm);
        const float off = new_m * kLog2e;
        float s0 = s_run * tc::ex2_approx((m_run - new_m) * kLog2e), s1 = 0.f, s2 = 0.f, s3 = 0.f;   // first chunk: 0 * 0
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            s0 += tc::ex2_approx(fmaf(z[i], kLog2e, -off));
            s1 += tc::ex2_approx(fmaf(z[i + 1], kLog2e, -off));
            s2 += tc::ex2_approx(fmaf(z[i + 2], kLog2e, -off));
            s3 += tc::ex2_approx(fmaf(z[i + 3], kLog2e, -off));
        }
        m_run = new_m;
        s_run = (s0 + s1) + (s2 + s3);
    }
    __device__ __forceinline__ void tile_end(int64_t row) {
        if (row < M) {
            pmax[row * n_blocks + nb] = m_run;
            psum[row * n_blocks + nb] = s_run;
        }
    }
};

// one warp per row: lse over the per-tile pairs and the pad logit, then the row loss
__global__ void lse_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ psum,
                                   const float* __restrict__ zgt, const float* __restrict__ zpad,
                                   const int64_t* __restrict__ gt, int64_t M, int64_t N, int64_t n_blocks,
                                   float* __restrict__ lse, float* __restrict__ loss_row) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float zp = zpad[row];
    float m = zp;
    for (int64_t b = lane; b < n_blocks; b += 32) m = fmaxf(m, pmax[row * n_blocks + b]);
    m = warp_max(m);
    float s = 0.f;
    for (int64_t b = lane; b < n_blocks; b += 32)
        s += psum[row * n_blocks + b] * exp2f((pmax[row * n_blocks + b] - m) * kLog2e);
    s = warp_sum(s);
    if (lane == 0) {
        s += exp2f((zp - m) * kLog2e);
        const float l = m + logf(s);
        lse[row] = l;
        const int64_t g = gt[row];
        loss_row[row] = (g >= 0 && g < N) ? l - zgt[row] : 0.f;
    }
}

struct GradEpilogue {
    const float* bias;        // [N]
    const int64_t* gt;        // [M]
    const float* lse;         // [M]
    const float* coef;        // [M]
    uint16_t *dz_hi, *dz_lo;      // [M, ldn]
    uint16_t *dzt_hi, *dzt_lo;    // [N, ldm]
    int64_t M, N, ldn, ldm;
    float l2, cf;             // lse * log2(e), row coefficient
    int64_t g;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t, int64_t row, int, int) {
        if (row < M) {
            g = gt[row];
            l2 = lse[row] * kLog2e;
            cf = (g >= 0 && g < N) ? coef[row] : 0.f;
        }
    }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M || col0 >= ldn) return;
        float dz[32];
        if (col0 + 32 <= N && (reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0) {
            const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                dz[4 * j] = v[4 * j] + b.x; dz[4 * j + 1] = v[4 * j + 1] + b.y;
                dz[4 * j + 2] = v[4 * j + 2] + b.z; dz[4 * j + 3] = v[4 * j + 3] + b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) dz[i] = col0 + i < N ? v[i] + __ldg(bias + col0 + i) : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) dz[i] = tc::ex2_approx(fmaf(dz[i], kLog2e, -l2)) * cf;     // softmax * coef; 0 past N
        if (g >= col0 && g < col0 + 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i == g) dz[i] -= cf;                                              // minus onehot * coef
        }
        // bf16 hi / lo pairs, packed two columns per register
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(dz[2 * j], dz[2 * j + 1]);
            hi[j] = *reinterpret_cast<const uint32_t*>(&h);
            const float r0 = dz[2 * j] - __uint_as_float(hi[j] << 16);
            const float r1 = dz[2 * j + 1] - __uint_as_float(hi[j] & 0xffff0000u);
            const __nv_bfloat162 q = __floats2bfloat162_rn(r0, r1);
            lo[j] = *reinterpret_cast<const uint32_t*>(&q);
        }
        // transposed copy dZ^T[c, row]: the 32 lanes of the warp write 32 consecutive rows of one column
        const int n_valid = N - col0 < 32 ? (int)(N - col0) : 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i < n_valid) {
                const int64_t o = (col0 + i) * ldm + row;
                dzt_hi[o] = (uint16_t)((i & 1) ? (hi[i >> 1] >> 16) : (hi[i >> 1] & 0xffffu));
                if (dzt_lo) dzt_lo[o] = (uint16_t)((i & 1) ? (lo[i >> 1] >> 16) : (lo[i >> 1] & 0xffffu));
            }
        }
        // row-major copy: 8 bf16 per 16-byte store (ldn is a multiple of 8, col0 of 32)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (col0 + 8 * j < ldn) {
                *reinterpret_cast<uint4*>(dz_hi + row * ldn + col0 + 8 * j) =
                    make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                if (dz_lo)
                    *reinterpret_cast<uint4*>(dz_lo + row * ldn + col0 + 8 * j) =
                        make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
        }
    }
    __device__ __forceinline__ void tile_end(int64_t) {}
};

struct StoreEpilogue {       // slab s of a split-K GEMM goes to C + s * slab_stride (C[row, c] = acc, or += acc)
    float* C;
    int64_t ldc, M, N;
    int accumulate;
    int64_t slab_stride;
    float* base;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t, int64_t, int slab, int) { base = C + slab * slab_stride; }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M) return;
        float* c = base + row * ldc + col0;
        if (col0 + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (accumulate) {
                    const float4 old = reinterpret_cast<float4*>(c)[j];
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                reinterpret_cast<float4*>(c)[j] = o;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < N) c[i] = accumulate ? c[i] + v[i] : v[i];
        }
    }
    __device__ __forceinline__ void tile_end(int64_t) {}
};

// out (+)= sum over slabs, in slab order (fp32 round-to-nearest adds between the tensor-core partial sums)
__global__ void slab_reduce_kernel(const float* __restrict__ part, int slabs, int64_t n, float* out, int accumulate) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < slabs; ++k) s += part[(int64_t)k * n + i];
        out[i] = accumulate ? out[i] + s : s;
    }
}

// db[n] += sum_m dZ^T[n, m] (hi + lo), one warp per row, fixed order
__global__ void rowsum_bf16_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int64_t rows,
                                   int64_t cols, int64_t ld, float* out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float s = 0.f;
    for (int64_t c = lane; c < cols; c += 32) {
        float x = bf16_to_f32(hi[r * ld + c]);
        if (lo) x += bf16_to_f32(lo[r * ld + c]);
        s += x;
    }
    s = warp_sum(s);
    if (lane == 0) out[r] += s;
}

__global__ void dzpad_kernel(const float* __restrict__ zpad, const float* __restrict__ lse,
                             const float* __restrict__ coef, const int64_t* __restrict__ gt, int64_t M, int64_t N,
                             float* __restrict__ dzpad) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t g = gt[m];
    dzpad[m] = (g >= 0 && g < N) ? expf(zpad[m] - lse[m]) * coef[m] : 0.f;
}

struct CeLayout {       // carve-up of the caller's workspace
    uint16_t *h_hi, *h_lo, *w_hi, *w_lo;            // [M, d], [N, d]
    uint16_t *ht_hi, *ht_lo, *wt_hi, *wt_lo;        // [d, ldm], [d, ldn]
    uint16_t *dz_hi, *dz_lo, *dzt_hi, *dzt_lo;      // [M, ldn], [N, ldm]
    float *pmax, *psum, *zgt;
    float* slabs;                                    // split-K partial sums of the gradient GEMMs
    int64_t ldn, ldm, n_blocks, bytes;
    int ks_dh, ks_dw;
};

static CeLayout ce_layout(void* ws, int64_t M, int64_t N, int d, int BN, bool backward) {
    CeLayout L;
    L.ldn = align_up(N, 8);
    L.ldm = align_up(M, 8);
    L.n_blocks = ceil_div(N, BN) * tc::EPI_PARTS;          // (max, sum) pairs per row: one per tile and column part
    char* p = (char*)ws;
    auto take = [&](int64_t bytes) {
        char* q = p;
        p += align_up(bytes, 256);
        return q;
    };
    L.h_hi = (uint16_t*)take(M * d * 2); L.h_lo = (uint16_t*)take(M * d * 2);
    L.w_hi = (uint16_t*)take(N * d * 2); L.w_lo = (uint16_t*)take(N * d * 2);
    L.pmax = (float*)take(M * L.n_blocks * 4); L.psum = (float*)take(M * L.n_blocks * 4);
    L.zgt = (float*)take(M * 4);
    L.ht_hi = L.ht_lo = L.wt_hi = L.wt_lo = L.dz_hi = L.dz_lo = L.dzt_hi = L.dzt_lo = nullptr;
    L.slabs = nullptr;
    // K slabs: enough tiles to fill the SMs, and short accumulation chains in tensor memory
    const int64_t tiles_dh = ceil_div(M, tc::BM) * ceil_div(d, 256), tiles_dw = ceil_div(N, tc::BM) * ceil_div(d, 256);
    auto pick = [](int64_t tiles, int64_t K) {
        int64_t want = ceil_div(2 * 148, tiles > 0 ? tiles : 1);
        int64_t by_k = ceil_div(K, 64 * 48);                  // at most ~48 k-blocks per slab
        int64_t s = want > by_k ? want : by_k;
        const int64_t max_s = ceil_div(K, 64 * 4);            // at least 4 k-blocks per slab
        if (s > max_s) s = max_s;
        if (s > 16) s = 16;
        return (int)(s < 1 ? 1 : s);
    };
    L.ks_dh = pick(tiles_dh, N);
    L.ks_dw = pick(tiles_dw, M);
    if (backward) {
        const int64_t slab_floats = (int64_t)L.ks_dh * M * d > (int64_t)L.ks_dw * N * d ? (int64_t)L.ks_dh * M * d
                                                                                        : (int64_t)L.ks_dw * N * d;
        L.slabs = (float*)take(slab_floats * 4);
        L.ht_hi = (uint16_t*)take(d * L.ldm * 2); L.ht_lo = (uint16_t*)take(d * L.ldm * 2);
        L.wt_hi = (uint16_t*)take(d * L.ldn * 2); L.wt_lo = (uint16_t*)take(d * L.ldn * 2);
        L.dz_hi = (uint16_t*)take(M * L.ldn * 2); L.dz_lo = (uint16_t*)take(M * L.ldn * 2);
        L.dzt_hi = (uint16_t*)take(N * L.ldm * 2); L.dzt_lo = (uint16_t*)take(N * L.ldm * 2);
    }
    L.bytes = p - (char*)ws;
    return L;
}

constexpr int kBN1 = 128, kStages1 = 3;     // logits GEMM (epilogue-heavy)
constexpr int kBN2 = 256, kStages2 = 2;     // gradient GEMMs: one tile spans d = 256 so dZ is read once

}  // namespace c2dsr

using namespace c2dsr;

#define RUN(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_) return rc_;   \
    } while (0)

extern "C" {

int64_t c2dsr_score_ce_tc_workspace_bytes(int64_t M, int64_t N, int d, int backward) {
    return ce_layout(nullptr, M, N, d, kBN1, backward != 0).bytes + 1024;
}

int c2dsr_score_ce_fwd_tc(const float* H, const float* W, const float* bias, const float* zpad, const int64_t* gt,
                          int64_t M, int64_t N, int d, int passes, float* lse, float* loss_row, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    if (M <= 0 || N <= 0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    if (workspace_bytes < c2dsr_score_ce_tc_workspace_bytes(M, N, d, 0)) {
        set_error("score_ce_fwd_tc: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const CeLayout L = ce_layout(workspace, M, N, d, kBN1, false);
    const bool split = passes == 3;
    RUN(split_rows(H, M, d, d, L.h_hi, split ? L.h_lo : nullptr, st));
    RUN(split_rows(W, N, d, d, L.w_hi, split ? L.w_lo : nullptr, st));
    tc::Maps maps;
    RUN(make_maps<kBN1>(&maps, L.h_hi, L.h_lo, M, d, L.w_hi, L.w_lo, N, d, d, passes));
    tc::Problem pb{M, N, d, passes, 0, 1};
    LseEpilogue epi{bias, gt, L.pmax, L.psum, L.zgt, M, N, L.n_blocks, 0.f, 0.f, 0, 0};
    if (d <= tc::ARES_MAX_KB * tc::BK) {     // the H row block stays resident in shared memory
        RUN((launch_gemm<kBN1, kStages1, true>(maps, pb, epi, st)));
    } else {
        RUN((launch_gemm<kBN1, kStages1, false>(maps, pb, epi, st)));
    }
    lse_combine_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(L.pmax, L.psum, L.zgt, zpad, gt, M, N, L.n_blocks, lse,
                                                                 loss_row);
    note_launches(1);
    return check_launch("score_ce_fwd_tc");
}

int c2dsr_score_ce_bwd_tc(const float* H, const float* W, const float* bias, const float* zpad, const int64_t* gt,
                          const float* lse, const float* coef, int64_t M, int64_t N, int d, int passes, float* dH,
                          float* dW, float* dbias, float* dzpad, void* workspace, int64_t workspace_bytes,
                          void* stream) {
    if (M <= 0 || N <= 0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    if (workspace_bytes < c2dsr_score_ce_tc_workspace_bytes(M, N, d, 1)) {
        set_error("score_ce_bwd_tc: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const CeLayout L = ce_layout(workspace, M, N, d, kBN1, true);
    const bool split = passes == 3;
    RUN(split_rows(H, M, d, d, L.h_hi, split ? L.h_lo : nullptr, st));
    RUN(split_rows(W, N, d, d, L.w_hi, split ? L.w_lo : nullptr, st));
    RUN(split_rows_transposed(H, M, d, L.ldm, L.ht_hi, split ? L.ht_lo : nullptr, st));
    RUN(split_rows_transposed(W, N, d, L.ldn, L.wt_hi, split ? L.wt_lo : nullptr, st));
    // 1. recompute the logits, emit dZ (both orientations)
    {
        tc::Maps maps;
        RUN(make_maps<kBN1>(&maps, L.h_hi, L.h_lo, M, d, L.w_hi, L.w_lo, N, d, d, passes));
        tc::Problem pb{M, N, d, passes, 0, 1};
        GradEpilogue epi{bias, gt, lse, coef, L.dz_hi, split ? L.dz_lo : nullptr, L.dzt_hi, split ? L.dzt_lo : nullptr,
                         M, N, L.ldn, L.ldm, 0.f, 0.f, 0};
        if (d <= tc::ARES_MAX_KB * tc::BK) {
            RUN((launch_gemm<kBN1, kStages1, true>(maps, pb, epi, st)));
        } else {
            RUN((launch_gemm<kBN1, kStages1, false>(maps, pb, epi, st)));
        }
    }
    // 2. dH[M, d] = dZ[M, N] * W[N, d]      (A = dZ, B = W^T, K = N)
    {
        tc::Maps maps;
        RUN(make_maps<kBN2>(&maps, L.dz_hi, L.dz_lo, M, L.ldn, L.wt_hi, L.wt_lo, d, L.ldn, N, passes));
        tc::Problem pb{M, d, (int)N, passes, 0, L.ks_dh};
        StoreEpilogue epi{L.slabs, d, M, d, 0, M * (int64_t)d, nullptr};
        RUN((launch_gemm<kBN2, kStages2, false>(maps, pb, epi, st)));
        const int64_t n = M * (int64_t)d;
        slab_reduce_kernel<<<(unsigned)(ceil_div(n, 256) < 2368 ? ceil_div(n, 256) : 2368), 256, 0, st>>>(
            L.slabs, L.ks_dh, n, dH, 0);
        note_launches(1);
    }
    // 3. dW[N, d] += dZ^T[N, M] * H[M, d]   (A = dZ^T, B = H^T, K = M)
    {
        tc::Maps maps;
        RUN(make_maps<kBN2>(&maps, L.dzt_hi, L.dzt_lo, N, L.ldm, L.ht_hi, L.ht_lo, d, L.ldm, M, passes));
        tc::Problem pb{N, d, (int)M, passes, 0, L.ks_dw};
        StoreEpilogue epi{L.slabs, d, N, d, 0, N * (int64_t)d, nullptr};
        RUN((launch_gemm<kBN2, kStages2, false>(maps, pb, epi, st)));
        const int64_t n = N * (int64_t)d;
        slab_reduce_kernel<<<(unsigned)(ceil_div(n, 256) < 2368 ? ceil_div(n, 256) : 2368), 256, 0, st>>>(
            L.slabs, L.ks_dw, n, dW, 1);
        note_launches(1);
    }
    rowsum_bf16_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(L.dzt_hi, split ? L.dzt_lo : nullptr, N, M, L.ldm, dbias);
    dzpad_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(zpad, lse, coef, gt, M, N, dzpad);
    note_launches(2);
    return check_launch("score_ce_bwd_tc");
}

}  // extern "C"
