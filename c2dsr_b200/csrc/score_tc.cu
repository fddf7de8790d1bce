// K4b on tensor cores: full-catalogue scoring fused with the rank count (trainer.py:168-179 of the
// reference).  scores = Q W^T + b are produced tile by tile in tensor memory by tcgen05 MMAs fed by TMA
// and consumed in the epilogue (compare with the target score, count) -- the [n_q, N] score matrix
// never reaches HBM.  The target score itself comes from the same MMA sequence applied to the gathered
// target rows, so "s_j > s_gt" compares numbers produced by identical arithmetic.
#include "tc_host.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

// 16-byte loads, 8-byte stores (d, ld_out multiples of 4, aligned pointers)
__global__ void split_bf16_vec_kernel(const float* __restrict__ X, int64_t rows, int d, int64_t ld_out,
                                      uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int q = d >> 2;
    const int64_t total = rows * (int64_t)q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / q;
        const int c = (int)(i - r * q) << 2;
        const float4 x = *reinterpret_cast<const float4*>(X + r * d + c);
        uint16_t h[4], l[4];
        split2(x.x, h[0], l[0]);
        split2(x.y, h[1], l[1]);
        split2(x.z, h[2], l[2]);
        split2(x.w, h[3], l[3]);
        uint2 ph, pl;
        ph.x = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
        ph.y = (uint32_t)h[2] | ((uint32_t)h[3] << 16);
        pl.x = (uint32_t)l[0] | ((uint32_t)l[1] << 16);
        pl.y = (uint32_t)l[2] | ((uint32_t)l[3] << 16);
        *reinterpret_cast<uint2*>(hi + r * ld_out + c) = ph;
        if (lo) *reinterpret_cast<uint2*>(lo + r * ld_out + c) = pl;
    }
}

__global__ void split_bf16_kernel(const float* __restrict__ X, int64_t rows, int d, int64_t ld_out,
                                  uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int64_t total = rows * (int64_t)d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int c = (int)(i % d);
        uint16_t h, l;
        split2(X[i], h, l);
        hi[r * ld_out + c] = h;
        if (lo) lo[r * ld_out + c] = l;
    }
}

// 32 x 32 tiles through shared memory: coalesced reads of X rows, coalesced writes of XT rows
__global__ void split_bf16_T_kernel(const float* __restrict__ X, int64_t rows, int d, int64_t ld_out,
                                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int64_t r = r0 + j;
        const int c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < d) ? X[r * d + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c = c0 + j;
        const int64_t r = r0 + threadIdx.x;
        if (c < d && r < rows) {
            uint16_t h, l;
            split2(tile[threadIdx.x][j], h, l);
            hi[(int64_t)c * ld_out + r] = h;
            if (lo) lo[(int64_t)c * ld_out + r] = l;
        }
    }
}

int split_rows(const float* X, int64_t rows, int d, int64_t ld_out, uint16_t* hi, uint16_t* lo, cudaStream_t st) {
    if (rows <= 0) return C2DSR_OK;
    const bool vec = (d & 3) == 0 && (ld_out & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
                     ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0;
    int64_t blocks = ceil_div(rows * (int64_t)d, vec ? 1024 : 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (vec) split_bf16_vec_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, rows, d, ld_out, hi, lo);
    else split_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, rows, d, ld_out, hi, lo);
    note_launches(1);
    return check_launch("split_bf16");
}

int split_rows_transposed(const float* X, int64_t rows, int d, int64_t ld_out, uint16_t* hi, uint16_t* lo,
                          cudaStream_t st) {
    if (rows <= 0) return C2DSR_OK;
    split_bf16_T_kernel<<<dim3((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(d, 32)), dim3(32, 8), 0, st>>>(
        X, rows, d, ld_out, hi, lo);
    note_launches(1);
    return check_launch("split_bf16_T");
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

// The same from the fp32 classifier itself (every rank holds all of it): G = split(W[gt]), for EVERY query, whatever
// shard owns the target -- bit-identical to gathering from the shard's split (same split2), and the target scores
// then need no exchange between the catalogue shards.
__global__ void gather_target_rows_f32_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                              const int64_t* __restrict__ gt, int64_t n_items, int d,
                                              uint16_t* __restrict__ G_hi, uint16_t* __restrict__ G_lo,
                                              float* __restrict__ bias_gt) {
    const int64_t i = blockIdx.x;
    const int64_t g = gt[i];
    const bool ok = g >= 0 && g < n_items;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        uint16_t h = 0, l = 0;
        if (ok) split2(W[g * d + c], h, l);
        G_hi[i * d + c] = h;
        if (G_lo) G_lo[i * d + c] = l;
    }
    if (threadIdx.x == 0) bias_gt[i] = ok ? bias[g] : 0.f;
}

// ---- epilogues -----------------------------------------------------------------------------------
constexpr int kBNc = 128;        // tile width of the kernels these epilogues run in
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
    __device__ __forceinline__ void tile_begin(int64_t, int64_t n_blk, int64_t row, int, int part) {
        cnt = 0;
        const int64_t c0 = n_blk * kBNc + part * (kBNc / tc::EPI_PARTS);
        if (c0 < N) tc::prefetch_l1(bias + c0);
        if (c0 + 32 < N) tc::prefetch_l1(bias + c0 + 32);
        if (row < M) {
            tgt = s_gt[row];
            g_local = gt[row] - n0;
        }
    }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M) return;
        if (col0 + 32 <= N && !S_debug && (g_local < col0 || g_local >= col0 + 32) &&
            (reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0) {
            // interior chunk without the target: add the bias, compare, count
            const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                cnt += (v[4 * j] + b.x > tgt) + (v[4 * j + 1] + b.y > tgt) + (v[4 * j + 2] + b.z > tgt) +
                       (v[4 * j + 3] + b.w > tgt);
            }
            return;
        }
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
    __device__ __forceinline__ void tile_begin(int64_t, int64_t, int64_t, int, int) {}
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

constexpr int kBN = 128, kStages = 3;

// ---- evaluation batch on the device: stable partition of the queries by domain -------------------------------
// one block: slot[i] = number of earlier queries of the same domain (block-wide exclusive scan over chunks of 1024)
__global__ void __launch_bounds__(1024) eval_slots_kernel(const int64_t* __restrict__ dom, int64_t B,
                                                          int32_t* __restrict__ slot, int32_t* __restrict__ n_ab) {
    __shared__ int warp_tot[32];
    __shared__ int base_b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_b = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < B; c0 += 1024) {
        const int64_t i = c0 + threadIdx.x;
        const int is_b = (i < B && dom[i] != 0) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, is_b);
        const int before = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
        const int nb_before = base_b + wbase + before;              // B-domain queries among [0, i)
        if (i < B) slot[i] = is_b ? nb_before : (int)(i - nb_before);
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += warp_tot[w];
            base_b += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        n_ab[0] = (int)(B - base_b);
        n_ab[1] = base_b;
    }
}

// block per query: its row goes to slot[i] of its domain's buffer as bf16 hi / lo
__global__ void eval_scatter_kernel(const float* __restrict__ q, const int64_t* __restrict__ dom,
                                    const int64_t* __restrict__ gt, const int32_t* __restrict__ slot, int d,
                                    uint16_t* __restrict__ QA_hi, uint16_t* __restrict__ QA_lo,
                                    uint16_t* __restrict__ QB_hi, uint16_t* __restrict__ QB_lo,
                                    int64_t* __restrict__ gtA, int64_t* __restrict__ gtB) {
    const int64_t i = blockIdx.x;
    const bool is_b = dom[i] != 0;
    const int64_t r = slot[i];
    uint16_t* hi = (is_b ? QB_hi : QA_hi) + r * d;
    uint16_t* lo = is_b ? QB_lo : QA_lo;
    if (lo) lo += r * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        uint16_t h, l;
        split2(q[i * d + c], h, l);
        hi[c] = h;
        if (lo) lo[c] = l;
    }
    if (threadIdx.x == 0) (is_b ? gtB : gtA)[r] = gt[i];
}

__global__ void eval_ranks_kernel(const int32_t* __restrict__ cA, const int32_t* __restrict__ cB,
                                  const int32_t* __restrict__ slot, const int64_t* __restrict__ dom, int64_t B,
                                  int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const bool is_b = dom[i] != 0;
    out[i] = 1 + (is_b ? cB : cA)[slot[i]];
    out[B + i] = is_b ? 1 : 0;
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
    return split_rows(X, rows, d, ld_out, hi, lo, (cudaStream_t)stream);
}

int c2dsr_eval_partition(const float* q, const int64_t* dom, const int64_t* gt, int64_t B, int d, uint16_t* QA_hi,
                         uint16_t* QA_lo, uint16_t* QB_hi, uint16_t* QB_lo, int64_t* gtA, int64_t* gtB, int32_t* slot,
                         int32_t* n_ab, void* stream) {
    if (B <= 0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    cudaStream_t st = (cudaStream_t)stream;
    eval_slots_kernel<<<1, 1024, 0, st>>>(dom, B, slot, n_ab);
    eval_scatter_kernel<<<(unsigned)B, 128, 0, st>>>(q, dom, gt, slot, d, QA_hi, QA_lo, QB_hi, QB_lo, gtA, gtB);
    note_launches(2);
    return check_launch("eval_partition");
}

int c2dsr_eval_ranks(const int32_t* countsA, const int32_t* countsB, const int32_t* slot, const int64_t* dom, int64_t B,
                     int32_t* out, void* stream) {
    if (B <= 0) return C2DSR_OK;
    eval_ranks_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(countsA, countsB, slot, dom, B, out);
    note_launches(1);
    return check_launch("eval_ranks");
}

int64_t c2dsr_score_tc_workspace_bytes(int64_t n_q, int64_t n_shard, int d) {
    (void)n_shard;
    const int64_t rows = align_up(n_q > 0 ? n_q : 1, tc::BM);
    return 2 * rows * (int64_t)d * 2 + rows * 4 + 1024;      // gathered target rows (hi, lo) + their bias
}

int c2dsr_score_target_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const uint16_t* W_hi, const uint16_t* W_lo,
                          const float* bias, const int64_t* gt, int64_t n_q, int64_t n0, int64_t n1, int d,
                          int passes, const int* n_q_limit, float* s_gt, void* workspace, int64_t workspace_bytes,
                          void* stream) {
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
    RUN(make_maps<kBN>(&maps, Q_hi, Q_lo, n_q, d, G_hi, G_lo, n_q, d, d, passes));
    tc::Problem pb{n_q, n_q, d, passes, 1, 1, Q_hi, Q_lo, d, n_q_limit};
    DiagEpilogue epi{bias_gt, s_gt, n_q};
    if (d <= tc::ARES_MAX_KB * tc::BK) return launch_gemm<kBN, tc::ARES_STAGES, true, false, false>(maps, pb, epi, st);
    return launch_gemm<kBN, kStages, false, false, false>(maps, pb, epi, st);
}

int c2dsr_score_target_full_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const float* W, const float* bias,
                               const int64_t* gt, int64_t n_q, int64_t n_items, int d, int passes, const int* n_q_limit,
                               float* s_gt, void* workspace, int64_t workspace_bytes, void* stream) {
    if (n_q <= 0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    if (workspace_bytes < c2dsr_score_tc_workspace_bytes(n_q, n_items, d)) {
        set_error("score_target_full_tc: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = align_up(n_q, tc::BM);
    uint16_t* G_hi = (uint16_t*)workspace;
    uint16_t* G_lo = G_hi + rows * d;
    float* bias_gt = (float*)(G_lo + rows * d);
    gather_target_rows_f32_kernel<<<(unsigned)n_q, 128, 0, st>>>(W, bias, gt, n_items, d, G_hi, passes == 3 ? G_lo : nullptr,
                                                                 bias_gt);
    note_launches(1);
    tc::Maps maps;
    RUN(make_maps<kBN>(&maps, Q_hi, Q_lo, n_q, d, G_hi, G_lo, n_q, d, d, passes));
    tc::Problem pb{n_q, n_q, d, passes, 1, 1, Q_hi, Q_lo, d, n_q_limit};
    DiagEpilogue epi{bias_gt, s_gt, n_q};
    if (d <= tc::ARES_MAX_KB * tc::BK) return launch_gemm<kBN, tc::ARES_STAGES, true, false, false>(maps, pb, epi, st);
    return launch_gemm<kBN, kStages, false, false, false>(maps, pb, epi, st);
}

int c2dsr_score_count_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const uint16_t* W_hi, const uint16_t* W_lo,
                         const float* bias, const float* s_gt, const int64_t* gt, int64_t n_q, int64_t n0,
                         int64_t n1, int d, int passes, const int* n_q_limit, int32_t* counts, float* S_debug,
                         int64_t lds, void* workspace, int64_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    if (n_q <= 0 || n1 <= n0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    const int64_t n = n1 - n0;
    tc::Maps maps;
    RUN(make_maps<kBN>(&maps, Q_hi, Q_lo, n_q, d, W_hi, W_lo, n, d, d, passes));
    tc::Problem pb{n_q, n, d, passes, 0, 1, Q_hi, Q_lo, d, n_q_limit};
    CountEpilogue epi{bias, s_gt, gt, counts, S_debug, lds, n_q, n, n0, 0.f, 0, 0};
    // the queries' row block stays resident in shared memory when it fits (d <= 256); same MMA sequence either way
    if (d <= tc::ARES_MAX_KB * tc::BK) return launch_gemm<kBN, tc::ARES_STAGES, true, false, false>(maps, pb, epi, (cudaStream_t)stream);
    return launch_gemm<kBN, kStages, false, false, false>(maps, pb, epi, (cudaStream_t)stream);
}

}  // extern "C"
