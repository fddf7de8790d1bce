// fp32 SGEMM on the FFMA pipe (exact fp32 products, fp32 accumulate):
//   C[M,N] = drop(act(alpha * op(A) op(B) + bias)) + beta * C
// Used for the small dense layers of the encoder / discriminator and as the un-fused comparison path
// of the score GEMMs (the tcgen05 kernels in score_tc.cu are the fast path for those).
//
// Tiling: BM x BN x 8 per CTA (128x128 or 64x64), 256 threads, (BM/16) x (BN/16) outputs per thread,
// double-buffered shared tiles stored k-major so fragments are float4 LDS.  All four transpose
// combinations share the kernel; loads are float4 when the operand is 16-byte aligned, guarded scalar
// otherwise.  Small grids with a long K get a deterministic split-K (partials summed in split order).
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

constexpr int BK = 8;

struct Epilogue {
    float alpha, beta;
    const float* bias;
    int act;
    Dropout dr;
};

__device__ __forceinline__ float4 load4_guarded(const float* __restrict__ base, int64_t r, int64_t c, int64_t ld,
                                                int64_t nr, int64_t nc, bool vec_ok) {
    // element (r, c..c+3) of a row-major [nr, nc] matrix with leading dimension ld
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= nr) return v;
    const float* p = base + r * ld + c;
    if (vec_ok && c + 3 < nc) return __ldg(reinterpret_cast<const float4*>(p));
    if (c < nc) v.x = __ldg(p);
    if (c + 1 < nc) v.y = __ldg(p + 1);
    if (c + 2 < nc) v.z = __ldg(p + 2);
    if (c + 3 < nc) v.w = __ldg(p + 3);
    return v;
}

__device__ __forceinline__ float apply_epilogue(float acc, const Epilogue& e, int64_t m, int64_t n, int64_t N,
                                                float c_old) {
    float v = e.alpha * acc;
    if (e.bias) v += e.bias[n];
    if (e.act == 1) v = fmaxf(v, 0.f);
    if (e.dr.p != 0.f) v *= drop_scale(e.dr, (uint64_t)m * N + n);
    if (e.beta != 0.f) v += e.beta * c_old;
    return v;
}

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda,
             const float* __restrict__ B, int64_t ldb, float* C, int64_t ldc, Epilogue ep, int64_t k_per_split,
             float* splitk_ws, bool a_vec, bool b_vec) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int GM = TM / 4, GN = TN / 4;          // groups of 4 consecutive rows / cols per thread
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t k_beg = (int64_t)blockIdx.z * k_per_split;
    const int64_t k_end = k_beg + k_per_split < K ? k_beg + k_per_split : K;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // per-thread load slots (one float4 of A and of B per k-tile when BM = BN = 128)
    const bool a_active = tid < BM * 2, b_active = tid < BN * 2;
    float4 a_reg = make_float4(0.f, 0.f, 0.f, 0.f), b_reg = a_reg;

    auto fetch = [&](int64_t k0) {
        if (a_active) {
            if (!TA) {      // A[M,K]: float4 along k
                const int r = tid >> 1, kq = (tid & 1) * 4;
                a_reg = load4_guarded(A, m0 + r, k0 + kq, lda, M, k_end, a_vec);
            } else {        // A stored [K,M]: float4 along m
                const int k = tid / (BM / 4), mq = (tid % (BM / 4)) * 4;
                a_reg = load4_guarded(A, k0 + k, m0 + mq, lda, k_end, M, a_vec);
            }
        }
        if (b_active) {
            if (TB) {       // B stored [N,K]: float4 along k
                const int r = tid >> 1, kq = (tid & 1) * 4;
                b_reg = load4_guarded(B, n0 + r, k0 + kq, ldb, N, k_end, b_vec);
            } else {        // B[K,N]: float4 along n
                const int k = tid / (BN / 4), nq = (tid % (BN / 4)) * 4;
                b_reg = load4_guarded(B, k0 + k, n0 + nq, ldb, k_end, N, b_vec);
            }
        }
    };
    auto stash = [&](int buf) {
        if (a_active) {
            if (!TA) {
                const int r = tid >> 1, kq = (tid & 1) * 4;
                As[buf][kq][r] = a_reg.x; As[buf][kq + 1][r] = a_reg.y;
                As[buf][kq + 2][r] = a_reg.z; As[buf][kq + 3][r] = a_reg.w;
            } else {
                const int k = tid / (BM / 4), mq = (tid % (BM / 4)) * 4;
                *reinterpret_cast<float4*>(&As[buf][k][mq]) = a_reg;
            }
        }
        if (b_active) {
            if (TB) {
                const int r = tid >> 1, kq = (tid & 1) * 4;
                Bs[buf][kq][r] = b_reg.x; Bs[buf][kq + 1][r] = b_reg.y;
                Bs[buf][kq + 2][r] = b_reg.z; Bs[buf][kq + 3][r] = b_reg.w;
            } else {
                const int k = tid / (BN / 4), nq = (tid % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][k][nq]) = b_reg;
            }
        }
    };

    const int64_t n_tiles = k_end > k_beg ? (k_end - k_beg + BK - 1) / BK : 0;
    if (n_tiles > 0) {
        fetch(k_beg);
        stash(0);
    }
    __syncthreads();
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int cur = (int)(t & 1);
        if (t + 1 < n_tiles) fetch(k_beg + (t + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a_frag[TM], b_frag[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][g * 64 + ty * 4]);
                a_frag[4 * g] = v.x; a_frag[4 * g + 1] = v.y; a_frag[4 * g + 2] = v.z; a_frag[4 * g + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[cur][kk][g * 64 + tx * 4]);
                b_frag[4 * g] = v.x; b_frag[4 * g + 1] = v.y; b_frag[4 * g + 2] = v.z; b_frag[4 * g + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a_frag[i], b_frag[j], acc[i][j]);
        }
        if (t + 1 < n_tiles) stash(cur ^ 1);
        __syncthreads();
    }

    // epilogue (or raw partial store for split-K)
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t n = n0 + (j / 4) * 64 + tx * 4 + (j % 4);
            if (n >= N) continue;
            if (splitk_ws) {
                splitk_ws[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
            } else {
                float* c = C + m * ldc + n;
                *c = apply_epilogue(acc[i][j], ep, m, n, N, ep.beta != 0.f ? *c : 0.f);
            }
        }
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t M, int64_t N, float* C,
                                     int64_t ldc, Epilogue ep) {
    const int64_t total = M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * total + i];
        const int64_t m = i / N, n = i % N;
        float* c = C + m * ldc + n;
        *c = apply_epilogue(s, ep, m, n, N, ep.beta != 0.f ? *c : 0.f);
    }
}

static bool aligned16(const void* p, int64_t ld) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0; }

template <int BM, int BN>
static void launch_tile(int ta, int tb, dim3 grid, cudaStream_t st, int64_t M, int64_t N, int64_t K, const float* A,
                        int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, const Epilogue& ep,
                        int64_t kps, float* ws) {
    const bool av = aligned16(A, lda), bv = aligned16(B, ldb);
    if (!ta && !tb) sgemm_kernel<BM, BN, false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, kps, ws, av, bv);
    else if (!ta && tb) sgemm_kernel<BM, BN, false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, kps, ws, av, bv);
    else if (ta && !tb) sgemm_kernel<BM, BN, true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, kps, ws, av, bv);
    else sgemm_kernel<BM, BN, true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, kps, ws, av, bv);
}

// Host-side dispatch shared by the C entry point and the composite kernels (encoder, infomax, score).
int gemm_dispatch(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                  const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act,
                  Dropout dr, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    if (M <= 0 || N <= 0) return C2DSR_OK;
    Epilogue ep{alpha, beta, bias, act, dr};
    const bool small = (M <= 64 || N <= 64 || ceil_div(M, 128) * ceil_div(N, 128) < 74);
    const int bm = small ? 64 : 128;
    const int64_t tiles = ceil_div(M, bm) * ceil_div(N, bm);
    int splits = 1;
    // few output tiles: cut K into slabs (added in slab order by a second kernel) so that more than a handful of
    // SMs work -- also for the short K = 256 products of the infomax block (16 tiles of 64 x 64 otherwise)
    if (((tiles < 120 && K >= 512) || (tiles < 40 && K >= 256)) && workspace) {
        splits = (int)ceil_div(296, tiles);
        const int64_t max_by_k = K >= 512 ? K / 128 : K / 64;
        if (splits > max_by_k) splits = (int)max_by_k;
        if (splits > 64) splits = 64;
        while (splits > 1 && (int64_t)splits * M * N * 4 > workspace_bytes) --splits;
        if (splits < 1) splits = 1;
    }
    int64_t kps = K;
    if (splits > 1) {
        kps = align_up(ceil_div(K, splits), BK);
        splits = (int)ceil_div(K, kps);
    }
    float* ws = splits > 1 ? (float*)workspace : nullptr;
    dim3 grid((unsigned)ceil_div(N, bm), (unsigned)ceil_div(M, bm), (unsigned)splits);
    if (small) launch_tile<64, 64>(ta, tb, grid, st, M, N, K, A, lda, B, ldb, C, ldc, ep, kps, ws);
    else launch_tile<128, 128>(ta, tb, grid, st, M, N, K, A, lda, B, ldb, C, ldc, ep, kps, ws);
    note_launches(splits > 1 ? 2 : 1);
    if (splits > 1) {
        int64_t blocks = ceil_div(M * N, 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(ws, splits, M, N, C, ldc, ep);
    }
    return check_launch("gemm");
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int64_t c2dsr_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    (void)K;
    const int64_t tiles = ceil_div(M, 64) * ceil_div(N, 64);
    if (tiles >= 120) return 256;
    int64_t splits = ceil_div(296, tiles > 0 ? tiles : 1);
    if (splits > 64) splits = 64;
    return splits * M * N * 4 + 256;
}

int c2dsr_gemm(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
               const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act,
               float p, uint64_t seed, uint64_t tag, void* workspace, int64_t workspace_bytes, void* stream) {
    C2DSR_REQUIRE(K >= 0 && lda > 0 && ldb > 0 && ldc > 0, "bad dimensions");
    C2DSR_REQUIRE(act == 0 || act == 1, "act must be 0 or 1");
    return gemm_dispatch(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act,
                         make_dropout(p, seed, tag), workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
