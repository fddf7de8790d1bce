// K4a on tensor cores: classifier logits + cross-entropy for training (trainer.py:131-152 of the
// reference) without ever materialising the fp32 logits.
//
// forward   Z = H W^T + b tile by tile in tensor memory (tcgen05, bf16 hi/lo split, fp32 accumulate);
//           the epilogue keeps, per row and tile part, an online (max, sum exp) pair and picks the target
//           logit; a small kernel combines the pairs (+ the pad logit) into lse and loss.
// backward  the same GEMM is recomputed; its epilogue turns each tile into dZ = (softmax - onehot) * coef
//           and stores it once, row-major [M, N], as bf16 hi/lo.  Two more tcgen05 GEMMs read it in both
//           orientations without a transposed copy: dH = dZ W (dZ K-major, W MN-major, K = N) and
//           dW += dZ^T H (dZ and H MN-major, K = M).  db is a column sum of dZ.
// All reductions have a fixed order (per-tile partials, K slabs added in order), so results are deterministic.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "tc_host.cuh"
#include "ce_bwd_fused.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void add_bias32(const float (&v)[32], const float* __restrict__ bias, int64_t col0,
                                           int64_t N, float (&z)[32]) {
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
        for (int i = 0; i < 32; ++i) z[i] = col0 + i < N ? v[i] + __ldg(bias + col0 + i) : -INFINITY;   // past N: -inf
    }
}

struct LseEpilogue {
    const float* bias;        // [N]
    const int64_t* gt;        // [M], N = ignored
    float* pmax;              // [M, n_pairs]
    float* psum;              // [M, n_pairs]
    float* zgt;               // [M] target logit (written by the thread that holds it)
    int64_t M, N, n_pairs;
    float m_run, s_run;
    int64_t g, slot;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t n_blk, int64_t row, int, int part) {
        const int64_t c0 = n_blk * 128 + part * (128 / tc::EPI_PARTS);
        if (c0 < N) tc::prefetch_l1(bias + c0);
        if (c0 + 32 < N) tc::prefetch_l1(bias + c0 + 32);
        m_run = -INFINITY;
        s_run = 0.f;
        slot = n_blk * tc::EPI_PARTS + part;                 // one (max, sum) pair per row, tile and column part
        g = row < M ? gt[row] : -1;
    }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M) return;
        float z[32];
        add_bias32(v, bias, col0, N, z);
        float cm = -INFINITY;
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
            pmax[row * n_pairs + slot] = m_run;
            psum[row * n_pairs + slot] = s_run;
        }
    }
};

// one warp per row: lse over the per-tile pairs and the pad logit, then the row loss
__global__ void lse_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ psum,
                                   const float* __restrict__ zgt, const float* __restrict__ zpad,
                                   const int64_t* __restrict__ gt, int64_t M, int64_t N, int64_t n_pairs,
                                   float* __restrict__ lse, float* __restrict__ loss_row) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float zp = zpad[row];
    float m = zp;
    for (int64_t b = lane; b < n_pairs; b += 32) m = fmaxf(m, pmax[row * n_pairs + b]);
    m = warp_max(m);
    float s = 0.f;
    for (int64_t b = lane; b < n_pairs; b += 32)
        s += psum[row * n_pairs + b] * exp2f((pmax[row * n_pairs + b] - m) * kLog2e);   // empty part: 0 * exp2(-inf) = 0
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
    uint16_t *dz_hi, *dz_lo;  // [M, ldn]
    int64_t M, N, ldn;
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
        add_bias32(v, bias, col0, N, dz);
#pragma unroll
        for (int i = 0; i < 32; ++i) dz[i] = tc::ex2_approx(fmaf(dz[i], kLog2e, -l2)) * cf;     // softmax * coef; 0 past N
        if (g >= col0 && g < col0 + 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i == g) dz[i] -= cf;                                              // minus onehot * coef
        }
        // bf16 hi / lo, two columns per register, 8 columns per 16-byte store (ldn % 8 == 0, col0 % 32 == 0)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (col0 + 8 * j >= ldn) break;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = dz[8 * j + 2 * k], b = dz[8 * j + 2 * k + 1];
                const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                hi[k] = *reinterpret_cast<const uint32_t*>(&h);
                const __nv_bfloat162 q = __floats2bfloat162_rn(a - __uint_as_float(hi[k] << 16),
                                                               b - __uint_as_float(hi[k] & 0xffff0000u));
                lo[k] = *reinterpret_cast<const uint32_t*>(&q);
            }
            *reinterpret_cast<uint4*>(dz_hi + row * ldn + col0 + 8 * j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (dz_lo) *reinterpret_cast<uint4*>(dz_lo + row * ldn + col0 + 8 * j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    __device__ __forceinline__ void tile_end(int64_t) {}
};

struct SlabEpilogue {        // slab s of a split-K GEMM goes to C + s * slab_stride, compact [M, N]
    float* C;
    int64_t M, N, slab_stride;
    float* base;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t, int64_t, int slab, int) { base = C + slab * slab_stride; }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M) return;
        float* c = base + row * N + col0;
        if (col0 + 32 <= N && (N & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(c)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < N) c[i] = v[i];
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

// db: column sums of dZ (hi + lo) in two phases: row chunks of 64 -> partial[chunk][col], then chunk order
constexpr int kDbChunk = 64;
__global__ void dz_colsum_partial_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int64_t M,
                                         int64_t N, int64_t ld, float* __restrict__ partial) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int64_t r0 = (int64_t)blockIdx.y * kDbChunk, r1 = r0 + kDbChunk < M ? r0 + kDbChunk : M;
    float s = 0.f;
#pragma unroll 8
    for (int64_t r = r0; r < r1; ++r) {
        float x = bf16_to_f32(hi[r * ld + c]);
        if (lo) x += bf16_to_f32(lo[r * ld + c]);
        s += x;
    }
    partial[(int64_t)blockIdx.y * N + c] = s;
}
__global__ void dz_colsum_final_kernel(const float* __restrict__ partial, int64_t n_chunks, int64_t N, float* out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    float s = 0.f;
    for (int64_t k = 0; k < n_chunks; ++k) s += partial[k * N + c];
    out[c] += s;
}

__global__ void dzpad_kernel(const float* __restrict__ zpad, const float* __restrict__ lse,
                             const float* __restrict__ coef, const int64_t* __restrict__ gt, int64_t M, int64_t N,
                             float* __restrict__ dzpad) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t g = gt[m];
    dzpad[m] = (g >= 0 && g < N) ? expf(zpad[m] - lse[m]) * coef[m] : 0.f;
}

constexpr int kBN1 = 128, kStages1 = 3;     // logits GEMM (epilogue-heavy)
constexpr int kBN2 = 256, kStages2 = 2;     // gradient GEMMs: one tile spans d = 256 so dZ is read once

struct CeLayout {       // carve-up of the caller's workspace
    uint16_t *h_hi, *h_lo, *w_hi, *w_lo;    // [M, d], [N, d]
    uint16_t *dz_hi, *dz_lo;                // [M, ldn]
    float *pmax, *psum, *zgt;
    float* slabs;                           // split-K partial sums of the gradient GEMMs
    float* db_part;                         // chunk partials of the dZ column sums (own buffer: runs beside the GEMMs)
    int64_t ldn, n_pairs, bytes;
    int ks_dh, ks_dw;
};

static CeLayout ce_layout(void* ws, int64_t M, int64_t N, int d, bool backward) {
    CeLayout L;
    L.ldn = align_up(N, 8);
    L.n_pairs = ceil_div(N, kBN1) * tc::EPI_PARTS;
    char* p = (char*)ws;
    auto take = [&](int64_t bytes) {
        char* q = p;
        p += align_up(bytes, 256);
        return q;
    };
    L.h_hi = (uint16_t*)take(M * d * 2); L.h_lo = (uint16_t*)take(M * d * 2);
    L.w_hi = (uint16_t*)take(N * d * 2); L.w_lo = (uint16_t*)take(N * d * 2);
    L.pmax = (float*)take(M * L.n_pairs * 4); L.psum = (float*)take(M * L.n_pairs * 4);
    L.zgt = (float*)take(M * 4);
    L.dz_hi = L.dz_lo = nullptr;
    L.slabs = L.db_part = nullptr;
    // K slabs: enough tiles to fill the SMs, and short accumulation chains in tensor memory
    const int64_t tiles_dh = ceil_div(M, tc::BM) * ceil_div(d, kBN2), tiles_dw = ceil_div(N, tc::BM) * ceil_div(d, kBN2);
    auto pick = [](int64_t tiles, int64_t K) {
        const int64_t want = ceil_div(2 * 148, tiles > 0 ? tiles : 1);
        const int64_t by_k = ceil_div(K, 64 * 48);            // at most ~48 k-blocks per slab
        int64_t s = want > by_k ? want : by_k;
        const int64_t max_s = ceil_div(K, 64 * 4);            // at least 4 k-blocks per slab
        if (s > max_s) s = max_s;
        if (s > 16) s = 16;
        return effective_splits(K, s);
    };
    L.ks_dh = pick(tiles_dh, N);
    L.ks_dw = pick(tiles_dw, M);
    if (backward) {
        L.dz_hi = (uint16_t*)take(M * L.ldn * 2); L.dz_lo = (uint16_t*)take(M * L.ldn * 2);
        int64_t f = (int64_t)L.ks_dh * M * d;
        if ((int64_t)L.ks_dw * N * d > f) f = (int64_t)L.ks_dw * N * d;
        L.slabs = (float*)take(f * 4);
        L.db_part = (float*)take(ceil_div(M, kDbChunk) * N * 4);
    }
    L.bytes = p - (char*)ws;
    return L;
}

// ---- fused backward (ce_bwd_fused.cuh): parameter preparation, partial reduction ---------------------------------
// l2s / cfs / g32 [M_pad] and bl [N_pad], padded with values that make dZ exactly 0 (see CeBwdProblem)
__global__ void ce_prep_kernel(const float* __restrict__ lse, const float* __restrict__ coef,
                               const int64_t* __restrict__ gt, const float* __restrict__ bias, int64_t M, int64_t N,
                               int64_t M_pad, int64_t N_pad, float* __restrict__ l2s, float* __restrict__ cfs,
                               int* __restrict__ g32, float* __restrict__ bl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M_pad) {
        bool ok = false;
        if (i < M) {
            const int64_t g = gt[i];
            ok = g >= 0 && g < N;
            g32[i] = ok ? (int)g : -1;
        } else {
            g32[i] = -1;
        }
        l2s[i] = ok ? lse[i] * kLog2e : 1e30f;
        cfs[i] = ok ? coef[i] : 0.f;
    }
    if (i < N_pad) bl[i] = i < N ? bias[i] * kLog2e : -1e30f;
}

// out[x, :] (+)= sum over the partial slabs of x's block, in slab order
__global__ void ce_part_reduce_kernel(const float* __restrict__ part, int64_t rows, int d, int64_t y_tiles,
                                      int64_t n_tiles, int64_t grid, int max_slots, float* __restrict__ out,
                                      int accumulate) {
    const int64_t total = rows * (int64_t)(d >> 2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t x = i / (d >> 2);
        const int c = (int)(i - x * (d >> 2)) << 2;
        const int64_t xb = x / tc::BM;
        const int64_t c0 = tc::fb_cta_of_tile(xb * y_tiles, n_tiles, grid), c1 = tc::fb_cta_of_tile(xb * y_tiles + y_tiles - 1, n_tiles, grid);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t sl = 0; sl <= c1 - c0; ++sl) {
            const float4 v = *reinterpret_cast<const float4*>(part + ((xb * max_slots + sl) * tc::BM + (x - xb * tc::BM)) * d + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float4* o = reinterpret_cast<float4*>(out + x * d + c);
        if (accumulate) {
            const float4 p = *o;
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        *o = acc;
    }
}
__global__ void ce_db_reduce_kernel(const float* __restrict__ part, int64_t rows, int64_t y_tiles, int64_t n_tiles,
                                    int64_t grid, int max_slots, float* __restrict__ db) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= rows) return;
    const int64_t xb = x / tc::BM;
    const int64_t c0 = tc::fb_cta_of_tile(xb * y_tiles, n_tiles, grid), c1 = tc::fb_cta_of_tile(xb * y_tiles + y_tiles - 1, n_tiles, grid);
    float s = 0.f;
    for (int64_t sl = 0; sl <= c1 - c0; ++sl) {
        const float* p = part + ((xb * max_slots + sl) * 2) * tc::BM + (x - xb * tc::BM);
        s += p[0] + p[tc::BM];
    }
    db[x] += s;
}

struct FusedLayout {
    float *l2s, *cfs, *bl, *part_h, *part_w, *db_part;
    int* g32;
    int64_t M_pad, N_pad, bytes;
    int64_t xb_h, yt_h, tiles_h, grid_h, xb_w, yt_w, tiles_w, grid_w;
    int slots_h, slots_w;
};
static int fused_slots(int64_t x_blocks, int64_t y_tiles, int64_t grid) {
    const int64_t n_tiles = x_blocks * y_tiles;
    int64_t m = 1;
    for (int64_t xb = 0; xb < x_blocks; ++xb) {
        const int64_t c = tc::fb_cta_of_tile(xb * y_tiles + y_tiles - 1, n_tiles, grid) - tc::fb_cta_of_tile(xb * y_tiles, n_tiles, grid) + 1;
        if (c > m) m = c;
    }
    return (int)m;
}
static FusedLayout fused_layout(char* p0, int64_t M, int64_t N, int d) {
    FusedLayout L;
    char* p = p0;
    auto take = [&](int64_t bytes) {
        char* q = p;
        p += align_up(bytes, 256);
        return q;
    };
    L.M_pad = align_up(M, tc::BM);
    L.N_pad = align_up(N, tc::BM);
    L.l2s = (float*)take(L.M_pad * 4); L.cfs = (float*)take(L.M_pad * 4); L.g32 = (int*)take(L.M_pad * 4);
    L.bl = (float*)take(L.N_pad * 4);
    const int64_t sms = 148;            // (sizes only; the launch uses the real count, which is not larger on B200)
    L.xb_h = L.M_pad / tc::BM; L.yt_h = L.N_pad / tc::BM; L.tiles_h = L.xb_h * L.yt_h;
    L.grid_h = L.tiles_h < sms ? L.tiles_h : sms;
    L.xb_w = L.N_pad / tc::BM; L.yt_w = L.M_pad / tc::BM; L.tiles_w = L.xb_w * L.yt_w;
    L.grid_w = L.tiles_w < sms ? L.tiles_w : sms;
    L.slots_h = fused_slots(L.xb_h, L.yt_h, L.grid_h);
    L.slots_w = fused_slots(L.xb_w, L.yt_w, L.grid_w);
    L.part_h = (float*)take(L.xb_h * L.slots_h * tc::BM * (int64_t)d * 4);
    L.part_w = (float*)take(L.xb_w * L.slots_w * tc::BM * (int64_t)d * 4);
    L.db_part = (float*)take(L.xb_w * L.slots_w * 2 * tc::BM * 4);
    L.bytes = p - p0;
    return L;
}
static bool fused_ok(int d) { return d <= tc::ARES_MAX_KB * tc::BK && d % 8 == 0; }

}  // namespace c2dsr

using namespace c2dsr;

#define RUN(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_) return rc_;   \
    } while (0)

static unsigned ew_grid(int64_t n) {
    const int64_t b = ceil_div(n, 256);
    return (unsigned)(b < 2368 ? (b > 0 ? b : 1) : 2368);
}

// One helper stream + two events per device and host thread, created on first use (never during a stream capture:
// the first call of a process is an eager one).  Only used for work that is forked from and joined back into the
// caller's stream inside a single entry point.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    int device = -1;
    bool ok = false;
};
static SideStream& side_stream() {
    static thread_local SideStream tab[16];
    int dev = 0;
    cudaGetDevice(&dev);
    SideStream& s = tab[dev & 15];
    if (s.device != dev) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        s.ok = false;
        s.device = dev;
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess)
            s.ok = true;
        else
            cudaGetLastError();
        (void)cap;
    }
    return s;
}

extern "C" {

int64_t c2dsr_score_ce_tc_workspace_bytes(int64_t M, int64_t N, int d, int backward) {
    if (backward && fused_ok(d))      // fused backward: no dZ, no K slabs -- the operand splits + the partial slabs
        return ce_layout(nullptr, M, N, d, false).bytes + fused_layout(nullptr, M, N, d).bytes + 1024;
    return ce_layout(nullptr, M, N, d, backward != 0).bytes + 1024;
}

int c2dsr_score_ce_fwd_tc(const float* H, const float* W, const uint16_t* W_hi, const uint16_t* W_lo, const float* bias,
                          const float* zpad, const int64_t* gt, int64_t M, int64_t N, int d, int passes, float* lse,
                          float* loss_row, void* workspace, int64_t workspace_bytes, void* stream) {
    if (M <= 0 || N <= 0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    if (workspace_bytes < c2dsr_score_ce_tc_workspace_bytes(M, N, d, 0)) {
        set_error("score_ce_fwd_tc: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    CeLayout L = ce_layout(workspace, M, N, d, false);
    const bool split = passes == 3;
    RUN(split_rows(H, M, d, d, L.h_hi, split ? L.h_lo : nullptr, st));
    if (W_hi && (W_lo || !split)) {                 // the caller's split of W (shared by forward and backward)
        L.w_hi = const_cast<uint16_t*>(W_hi);
        L.w_lo = const_cast<uint16_t*>(W_lo);
    } else {
        RUN(split_rows(W, N, d, d, L.w_hi, split ? L.w_lo : nullptr, st));
    }
    tc::Maps maps;
    RUN(make_maps<kBN1>(&maps, L.h_hi, L.h_lo, M, d, L.w_hi, L.w_lo, N, d, d, passes));
    tc::Problem pb{M, N, d, passes, 0, 1, L.h_hi, L.h_lo, d};
    LseEpilogue epi{bias, gt, L.pmax, L.psum, L.zgt, M, N, L.n_pairs, 0.f, 0.f, 0, 0};
    if (d <= tc::ARES_MAX_KB * tc::BK) {     // the H row block stays resident in shared memory
        RUN((launch_gemm<kBN1, tc::ARES_STAGES, true, false, false>(maps, pb, epi, st)));
    } else {
        RUN((launch_gemm<kBN1, kStages1, false, false, false>(maps, pb, epi, st)));
    }
    lse_combine_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(L.pmax, L.psum, L.zgt, zpad, gt, M, N, L.n_pairs, lse,
                                                                 loss_row);
    note_launches(1);
    return check_launch("score_ce_fwd_tc");
}

int c2dsr_score_ce_bwd_tc(const float* H, const float* W, const uint16_t* W_hi, const uint16_t* W_lo, const float* bias,
                          const float* zpad, const int64_t* gt, const float* lse, const float* coef, int64_t M,
                          int64_t N, int d, int passes, float* dH, float* dW, float* dbias, float* dzpad,
                          void* workspace, int64_t workspace_bytes, void* stream) {
    if (M <= 0 || N <= 0) return C2DSR_OK;
    RUN(c2dsr_device_check());
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(d % 8 == 0, "d must be a multiple of 8");
    if (workspace_bytes < c2dsr_score_ce_tc_workspace_bytes(M, N, d, 1)) {
        set_error("score_ce_bwd_tc: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool split = passes == 3;
    const bool fused = fused_ok(d) && !getenv("C2DSR_CE_BWD_UNFUSED");
    CeLayout L = ce_layout(workspace, M, N, d, !fused);
    uint16_t* dz_lo = split ? L.dz_lo : nullptr;
    RUN(split_rows(H, M, d, d, L.h_hi, split ? L.h_lo : nullptr, st));
    if (W_hi && (W_lo || !split)) {
        L.w_hi = const_cast<uint16_t*>(W_hi);
        L.w_lo = const_cast<uint16_t*>(W_lo);
    } else {
        RUN(split_rows(W, N, d, d, L.w_hi, split ? L.w_lo : nullptr, st));
    }
    if (fused) {
        // dZ never reaches HBM: two launches of ce_bwd_kernel recompute the logits tile by tile, turn each tile into
        // dZ inside tensor memory and feed it straight back to the tensor cores (see ce_bwd_fused.cuh)
        FusedLayout F = fused_layout((char*)workspace + align_up(L.bytes, 256), M, N, d);
        const int64_t np = F.M_pad > F.N_pad ? F.M_pad : F.N_pad;
        ce_prep_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, st>>>(lse, coef, gt, bias, M, N, F.M_pad, F.N_pad, F.l2s, F.cfs,
                                                                    F.g32, F.bl);
        dzpad_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(zpad, lse, coef, gt, M, N, dzpad);
        note_launches(2);
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(tc::ce_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::FB_SMEM);
            cudaFuncSetAttribute(tc::ce_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::FB_SMEM);
            attr = true;
        }
        {   // dH: X = H (stationary), Y = W (streamed)
            tc::Maps maps;
            RUN(make_maps<tc::BM>(&maps, L.h_hi, L.h_lo, M, d, L.w_hi, L.w_lo, N, d, d, passes));
            tc::CeBwdProblem pb{M, N, d, passes, F.l2s, F.cfs, F.g32, F.bl, F.part_h, nullptr, F.slots_h};
            tc::ce_bwd_kernel<false><<<(unsigned)F.grid_h, tc::THREADS, tc::FB_SMEM, st>>>(maps, pb);
            ce_part_reduce_kernel<<<ew_grid(M * (int64_t)(d / 4)), 256, 0, st>>>(F.part_h, M, d, F.yt_h, F.tiles_h, F.grid_h,
                                                                              F.slots_h, dH, 0);
        }
        {   // dW, db: X = W (stationary), Y = H (streamed)
            tc::Maps maps;
            RUN(make_maps<tc::BM>(&maps, L.w_hi, L.w_lo, N, d, L.h_hi, L.h_lo, M, d, d, passes));
            tc::CeBwdProblem pb{N, M, d, passes, F.l2s, F.cfs, F.g32, F.bl, F.part_w, F.db_part, F.slots_w};
            tc::ce_bwd_kernel<true><<<(unsigned)F.grid_w, tc::THREADS, tc::FB_SMEM, st>>>(maps, pb);
            ce_part_reduce_kernel<<<ew_grid(N * (int64_t)(d / 4)), 256, 0, st>>>(F.part_w, N, d, F.yt_w, F.tiles_w, F.grid_w,
                                                                              F.slots_w, dW, 1);
            ce_db_reduce_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(F.db_part, N, F.yt_w, F.tiles_w, F.grid_w,
                                                                         F.slots_w, dbias);
        }
        note_launches(5);
        return check_launch("score_ce_bwd_tc (fused)");
    }
    // 1. recompute the logits, emit dZ [M, N] as bf16 hi / lo
    {
        tc::Maps maps;
        RUN(make_maps<kBN1>(&maps, L.h_hi, L.h_lo, M, d, L.w_hi, L.w_lo, N, d, d, passes));
        tc::Problem pb{M, N, d, passes, 0, 1, L.h_hi, L.h_lo, d};
        GradEpilogue epi{bias, gt, lse, coef, L.dz_hi, dz_lo, M, N, L.ldn, 0.f, 0.f, 0};
        if (d <= tc::ARES_MAX_KB * tc::BK) {
            RUN((launch_gemm<kBN1, tc::ARES_STAGES, true, false, false>(maps, pb, epi, st)));
        } else {
            RUN((launch_gemm<kBN1, kStages1, false, false, false>(maps, pb, epi, st)));
        }
    }
    // db[N] += column sums of dZ, and dzpad: HBM-bound reads of dZ that depend on nothing but step 1.  They run
    // on a side stream next to the two tensor-core GEMMs below (one persistent GEMM CTA per SM leaves room for
    // them) and are joined before returning; the fork / join are event edges, so they are captured like the rest.
    SideStream& side = side_stream();
    if (side.ok) {
        cudaEventRecord(side.fork, st);
        cudaStreamWaitEvent(side.stream, side.fork, 0);
    }
    {
        cudaStream_t s2 = side.ok ? side.stream : st;
        const int64_t n_chunks = ceil_div(M, kDbChunk);
        dz_colsum_partial_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)n_chunks), 256, 0, s2>>>(
            L.dz_hi, dz_lo, M, N, L.ldn, L.db_part);
        dz_colsum_final_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, s2>>>(L.db_part, n_chunks, N, dbias);
        dzpad_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, s2>>>(zpad, lse, coef, gt, M, N, dzpad);
        note_launches(3);
        if (side.ok) cudaEventRecord(side.join, side.stream);
    }
    // 2. dH[M, d] = dZ[M, N] W[N, d]: A = dZ (K-major, K = N), B = W read MN-major from its [N, d] storage
    {
        tc::Maps maps;
        RUN(make_maps<kBN2>(&maps, L.dz_hi, L.dz_lo, M, L.ldn, L.w_hi, L.w_lo, d, d, N, passes, false, true));
        tc::Problem pb{M, d, (int)N, passes, 0, L.ks_dh};
        SlabEpilogue epi{L.slabs, M, d, M * (int64_t)d, nullptr};
        RUN((launch_gemm<kBN2, kStages2, false, false, true>(maps, pb, epi, st)));
        slab_reduce_kernel<<<ew_grid(M * (int64_t)d), 256, 0, st>>>(L.slabs, L.ks_dh, M * (int64_t)d, dH, 0);
        note_launches(1);
    }
    // 3. dW[N, d] += dZ^T H: A = dZ read MN-major ([M, N] storage, K = M), B = H read MN-major ([M, d] storage)
    {
        tc::Maps maps;
        RUN(make_maps<kBN2>(&maps, L.dz_hi, L.dz_lo, N, L.ldn, L.h_hi, L.h_lo, d, d, M, passes, true, true));
        tc::Problem pb{N, d, (int)M, passes, 0, L.ks_dw};
        SlabEpilogue epi{L.slabs, N, d, N * (int64_t)d, nullptr};
        RUN((launch_gemm<kBN2, kStages2, false, true, true>(maps, pb, epi, st)));
        slab_reduce_kernel<<<ew_grid(N * (int64_t)d), 256, 0, st>>>(L.slabs, L.ks_dw, N * (int64_t)d, dW, 1);
        note_launches(1);
    }
    if (side.ok) cudaStreamWaitEvent(st, side.join, 0);
    return check_launch("score_ce_bwd_tc");
}

}  // extern "C"
