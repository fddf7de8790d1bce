// General dense product on tensor cores for the small layers of the path (encoder projections, their
// input / weight gradients):   C[M,N] = drop(act(op(A) op(B) + bias)) + beta * C,   beta in {0, 1}
// with the same operand conventions as c2dsr_gemm (ta / tb).  fp32 operands are split into bf16 hi / lo
// once per call; a transposed operand (ta = 1 or tb = 0) is consumed MN-major straight from its row-major
// storage, so no transposed copies are made.  Short output grids with a long K are cut into K slabs whose
// partial sums are added in slab order by a second kernel that also applies the epilogue.
#include "tc_host.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

struct LinearEpilogue {
    float* C;                 // output, or slab buffer when k_splits > 1
    int64_t ldc, M, N;
    const float* bias;
    int act;
    float beta;
    Dropout dr;
    int64_t slab_stride;      // 0: single pass, apply the epilogue here
    float* base;
    __device__ __forceinline__ void tile_begin(int64_t, int64_t, int64_t, int slab, int) {
        base = C + slab * slab_stride;
    }
    __device__ __forceinline__ void chunk(int64_t row, int64_t col0, const float (&v)[32]) {
        if (row >= M) return;
        if (slab_stride) {                      // raw partial sums, compact [M, N]
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
            return;
        }
        float* c = base + row * ldc + col0;
        if (col0 + 32 <= N && dr.p == 0.f && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(c) & 15) == 0 &&
            (!bias || (reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0)) {
            // interior chunk: 8 x 16-byte loads / stores per thread
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 x = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0) + j);
                    x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
                }
                if (act == 1) {
                    x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
                }
                if (beta != 0.f) {
                    const float4 o = reinterpret_cast<const float4*>(c)[j];
                    x.x += beta * o.x; x.y += beta * o.y; x.z += beta * o.z; x.w += beta * o.w;
                }
                reinterpret_cast<float4*>(c)[j] = x;
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int64_t n = col0 + i;
            if (n < N) {
                float x = v[i];
                if (bias) x += __ldg(bias + n);
                if (act == 1) x = fmaxf(x, 0.f);
                if (dr.p != 0.f) x *= drop_scale(dr, (uint64_t)row * N + n);
                if (beta != 0.f) x += beta * c[i];
                c[i] = x;
            }
        }
    }
    __device__ __forceinline__ void tile_end(int64_t) {}
};

__global__ void linear_slab_reduce_kernel(const float* __restrict__ part, int slabs, int64_t M, int64_t N, float* C,
                                          int64_t ldc, const float* __restrict__ bias, int act, float beta, Dropout dr) {
    const int64_t total = M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float x = 0.f;
        for (int k = 0; k < slabs; ++k) x += part[(int64_t)k * total + i];
        const int64_t m = i / N, n = i % N;
        if (bias) x += bias[n];
        if (act == 1) x = fmaxf(x, 0.f);
        if (dr.p != 0.f) x *= drop_scale(dr, (uint64_t)i);
        float* c = C + m * ldc + n;
        if (beta != 0.f) x += beta * *c;
        *c = x;
    }
}

// fp32 [rows, cols] with leading dimension ld_in -> bf16 hi / lo [rows, ld_out].  Both operands of a product
// are split by ONE launch (the first blocks_a blocks take job a, the rest job b); rows whose length and
// strides are multiples of 4 go through 16-byte loads and 8-byte stores.
struct SplitJob {
    const float* X;
    int64_t rows, cols, ld_in, ld_out;
    uint16_t* hi;
    uint16_t* lo;
};

__device__ __forceinline__ void split_job(const SplitJob& j, int64_t bid, int64_t nblk) {
    const bool vec = ((j.cols | j.ld_in | j.ld_out) & 3) == 0 && (reinterpret_cast<uintptr_t>(j.X) & 15) == 0 &&
                     ((reinterpret_cast<uintptr_t>(j.hi) | reinterpret_cast<uintptr_t>(j.lo)) & 7) == 0;
    if (vec) {
        const int64_t c4n = j.cols >> 2, total = j.rows * c4n;
        for (int64_t i = bid * blockDim.x + threadIdx.x; i < total; i += nblk * blockDim.x) {
            const int64_t r = i / c4n, c = (i - r * c4n) << 2;
            const float4 x = *reinterpret_cast<const float4*>(j.X + r * j.ld_in + c);
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
            *reinterpret_cast<uint2*>(j.hi + r * j.ld_out + c) = ph;
            if (j.lo) *reinterpret_cast<uint2*>(j.lo + r * j.ld_out + c) = pl;
        }
        return;
    }
    const int64_t total = j.rows * j.cols;
    for (int64_t i = bid * blockDim.x + threadIdx.x; i < total; i += nblk * blockDim.x) {
        const int64_t r = i / j.cols, c = i % j.cols;
        uint16_t h, l;
        split2(j.X[r * j.ld_in + c], h, l);
        j.hi[r * j.ld_out + c] = h;
        if (j.lo) j.lo[r * j.ld_out + c] = l;
    }
}

__global__ void split_pair_kernel(SplitJob a, SplitJob b, int blocks_a) {
    if ((int)blockIdx.x < blocks_a) split_job(a, blockIdx.x, blocks_a);
    else split_job(b, blockIdx.x - blocks_a, gridDim.x - blocks_a);
}

static int split_blocks(const SplitJob& j) {
    int64_t blocks = ceil_div(j.rows * j.cols, 256 * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    return blocks < 1 ? 1 : (int)blocks;
}

static int split_pair(const SplitJob& a, const SplitJob& b, cudaStream_t st) {
    const int na = split_blocks(a), nb = split_blocks(b);
    split_pair_kernel<<<(unsigned)(na + nb), 256, 0, st>>>(a, b, na);
    note_launches(1);
    return check_launch("split_pair");
}

constexpr int kLBN = 128, kLStages = 3;

static int pick_splits(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ceil_div(M, tc::BM) * ceil_div(N, kLBN);
    if (tiles >= 100 || K < 512) return 1;
    int64_t s = ceil_div(2 * 148, tiles);
    const int64_t max_s = K / 256 > 1 ? K / 256 : 1;       // at least 4 k-blocks per slab
    if (s > max_s) s = max_s;
    if (s > 32) s = 32;
    return effective_splits(K, s);
}

int64_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    const int64_t ldk = align_up(K, 8), ldm = align_up(M, 8), ldn = align_up(N, 8);
    const int64_t a = (M * ldk > K * ldm ? M * ldk : K * ldm), b = (N * ldk > K * ldn ? N * ldk : K * ldn);
    return align_up(a * 4, 256) + align_up(b * 4, 256) + align_up((int64_t)pick_splits(M, N, K) * M * N * 4, 256) + 1024;
}

// Shared by the C entry point and the encoder composite.
int gemm_tc_dispatch(int ta, int tb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act, Dropout dr, int passes,
                     void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    if (M <= 0 || N <= 0) return C2DSR_OK;
    if (workspace_bytes < gemm_tc_workspace_bytes(M, N, K)) {
        set_error("gemm_tc: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    const bool split = passes == 3;
    const int64_t ldk = align_up(K, 8), ldm = align_up(M, 8), ldn = align_up(N, 8);
    const int64_t a_elems = ta ? K * ldm : M * ldk, b_elems = tb ? N * ldk : K * ldn;
    char* p = (char*)workspace;
    uint16_t* a_hi = (uint16_t*)p;
    uint16_t* a_lo = a_hi + a_elems;
    p += align_up((M * ldk > K * ldm ? M * ldk : K * ldm) * 4, 256);
    uint16_t* b_hi = (uint16_t*)p;
    uint16_t* b_lo = b_hi + b_elems;
    p += align_up((N * ldk > K * ldn ? N * ldk : K * ldn) * 4, 256);
    float* slabs = (float*)p;
    int rc;
    // ta = 0: A is [M, K] (K-major operand); ta = 1: A is stored [K, M] (MN-major operand)
    // tb = 1: B is stored [N, K] (K-major operand); tb = 0: B is [K, N] (MN-major operand)
    const SplitJob ja = ta ? SplitJob{A, K, M, lda, ldm, a_hi, split ? a_lo : nullptr}
                           : SplitJob{A, M, K, lda, ldk, a_hi, split ? a_lo : nullptr};
    const SplitJob jb = tb ? SplitJob{B, N, K, ldb, ldk, b_hi, split ? b_lo : nullptr}
                           : SplitJob{B, K, N, ldb, ldn, b_hi, split ? b_lo : nullptr};
    if ((rc = split_pair(ja, jb, st))) return rc;
    const bool a_mn = ta != 0, b_mn = tb == 0;
    tc::Maps maps;
    if ((rc = make_maps<kLBN>(&maps, a_hi, a_lo, M, a_mn ? ldm : ldk, b_hi, b_lo, N, b_mn ? ldn : ldk, K, passes, a_mn,
                              b_mn))) return rc;
    const int ks = pick_splits(M, N, K);
    tc::Problem pb{M, N, (int)K, passes, 0, ks};
    LinearEpilogue epi{ks > 1 ? slabs : C, ldc, M, N, bias, act, beta, dr, ks > 1 ? M * N : 0, nullptr};
    if (!a_mn && !b_mn) rc = launch_gemm<kLBN, kLStages, false, false, false>(maps, pb, epi, st);
    else if (!a_mn && b_mn) rc = launch_gemm<kLBN, kLStages, false, false, true>(maps, pb, epi, st);
    else if (a_mn && b_mn) rc = launch_gemm<kLBN, kLStages, false, true, true>(maps, pb, epi, st);
    else rc = launch_gemm<kLBN, kLStages, false, true, false>(maps, pb, epi, st);
    if (rc) return rc;
    if (ks > 1) {
        int64_t blocks = ceil_div(M * N, 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        linear_slab_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(slabs, ks, M, N, C, ldc, bias, act, beta, dr);
        note_launches(1);
    }
    return check_launch("gemm_tc");
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int64_t c2dsr_gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K) { return gemm_tc_workspace_bytes(M, N, K); }

int c2dsr_gemm_tc(int ta, int tb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                  int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act, float p, uint64_t seed,
                  uint64_t tag, int passes, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = c2dsr_device_check();
    if (rc) return rc;
    C2DSR_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
    C2DSR_REQUIRE(beta == 0.f || beta == 1.f, "beta must be 0 or 1");
    C2DSR_REQUIRE(K > 0 && K < (1ll << 31), "bad K");
    return gemm_tc_dispatch(ta, tb, M, N, K, A, lda, B, ldb, beta, C, ldc, bias, act, make_dropout(p, seed, tag),
                            passes, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
