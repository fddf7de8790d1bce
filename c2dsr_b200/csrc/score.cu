// K4 (un-fused fp32 path): classifier logits + cross-entropy for training (trainer.py:131-152) and
// scoring + rank counting for evaluation (trainer.py:168-179).  The score matrix is materialised by
// the FFMA GEMM; score_tc.cu holds the tcgen05 path that keeps it out of HBM.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

int gemm_dispatch(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                  const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act,
                  Dropout dr, void* workspace, int64_t workspace_bytes, cudaStream_t st);

__device__ __forceinline__ float block_reduce_max(float v, float* sm) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : -INFINITY;
        r = warp_max(r);
        if (threadIdx.x == 0) sm[0] = r;
    }
    __syncthreads();
    const float out = sm[0];
    __syncthreads();
    return out;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* sm) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        r = warp_sum(r);
        if (threadIdx.x == 0) sm[0] = r;
    }
    __syncthreads();
    const float out = sm[0];
    __syncthreads();
    return out;
}

// one block per row: lse over [Z[m, 0..N) | zpad[m]], loss_row = lse - z[gt] (0 when gt == N: ignored)
__global__ void ce_row_kernel(const float* __restrict__ Z, int64_t ldz, const float* __restrict__ zpad,
                              const int64_t* __restrict__ gt, int64_t N, float* __restrict__ lse,
                              float* __restrict__ loss_row) {
    __shared__ float sm[32];
    const int64_t m = blockIdx.x;
    const float* z = Z + m * ldz;
    const float zp = zpad[m];
    float mx = zp;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, z[n]);
    mx = block_reduce_max(mx, sm);
    float s = 0.f;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) s += expf(z[n] - mx);
    s = block_reduce_sum(s, sm);
    if (threadIdx.x == 0) {
        s += expf(zp - mx);
        const float l = mx + logf(s);
        lse[m] = l;
        const int64_t g = gt[m];
        loss_row[m] = (g >= 0 && g < N) ? l - z[g] : 0.f;
    }
}

// Z <- (softmax - onehot) * coef ;  dzpad = softmax_pad * coef
__global__ void ce_grad_kernel(float* __restrict__ Z, int64_t ldz, const float* __restrict__ zpad,
                               const int64_t* __restrict__ gt, const float* __restrict__ lse,
                               const float* __restrict__ coef, int64_t N, float* __restrict__ dzpad) {
    const int64_t m = blockIdx.x;
    const int64_t g = gt[m];
    const bool valid = g >= 0 && g < N;
    const float c = valid ? coef[m] : 0.f;
    const float l = lse[m];
    float* z = Z + m * ldz;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
        float p = expf(z[n] - l);
        if (n == g) p -= 1.f;
        z[n] = p * c;
    }
    if (threadIdx.x == 0) dzpad[m] = expf(zpad[m] - l) * c;
}

__global__ void pick_target_kernel(const float* __restrict__ S, int64_t lds, const int64_t* __restrict__ gt,
                                   int64_t n_q, int64_t n0, int64_t n1, float* __restrict__ s_gt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_q) return;
    const int64_t g = gt[i];
    s_gt[i] = (g >= n0 && g < n1) ? S[i * lds + (g - n0)] : 0.f;
}

// one block per query; strict '>' in fp32; integer result
__global__ void rank_count_kernel(const float* __restrict__ S, int64_t lds, const float* __restrict__ s_gt,
                                  const int64_t* __restrict__ gt, const int64_t* __restrict__ neg, int64_t n_neg,
                                  int64_t n0, int64_t n1, int32_t* __restrict__ counts) {
    __shared__ int sm[32];
    const int64_t i = blockIdx.x;
    const float* s = S + i * lds;
    const float t = s_gt[i];
    const int64_t g = gt[i];
    int c = 0;
    if (neg) {
        const int64_t* nl = neg + i * n_neg;
        for (int64_t k = threadIdx.x; k < n_neg; k += blockDim.x) {
            const int64_t id = nl[k];
            if (id >= n0 && id < n1 && id != g) c += s[id - n0] > t;
        }
    } else {
        const int64_t n = n1 - n0;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) c += (j + n0 != g) && (s[j] > t);
    }
    c = warp_sum_int(c);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        int r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0;
        r = warp_sum_int(r);
        if (threadIdx.x == 0) counts[i] += r;
    }
}

// Stable partition of the row indices: rows whose target is not the ignore class first (original order),
// ignored rows after them.  One block; each thread owns a contiguous slice, block-wide exclusive scan of counts.
__global__ void compact_rows_kernel(const int64_t* __restrict__ gt, int64_t M, int64_t ignore, int64_t* __restrict__ perm) {
    __shared__ int counts[1024];
    const int tid = threadIdx.x;
    const int64_t per = (M + blockDim.x - 1) / blockDim.x;
    const int64_t lo = tid * per, hi = lo + per < M ? lo + per : M;
    int c = 0;
    for (int64_t i = lo; i < hi; ++i) c += gt[i] != ignore;
    counts[tid] = c;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {               // inclusive Hillis-Steele scan
        const int v = tid >= o ? counts[tid - o] : 0;
        __syncthreads();
        counts[tid] += v;
        __syncthreads();
    }
    const int total = counts[blockDim.x - 1];
    int64_t v_pos = counts[tid] - c;                         // valid rows before this slice
    int64_t i_pos = total + (lo < M ? lo : M) - v_pos;       // ignored rows go after all valid ones
    for (int64_t i = lo; i < hi; ++i) {
        if (gt[i] != ignore) perm[v_pos++] = i;
        else perm[i_pos++] = i;
    }
}

}  // namespace c2dsr

using namespace c2dsr;

#define RUN(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_) return rc_;   \
    } while (0)

extern "C" {

int64_t c2dsr_score_ldz(int64_t N) { return align_up(N, 4); }

int c2dsr_compact_rows(const int64_t* gt, int64_t M, int64_t ignore, int64_t* perm, void* stream) {
    if (M <= 0) return C2DSR_OK;
    compact_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(gt, M, ignore, perm);
    note_launches(1);
    return check_launch("compact_rows");
}

int c2dsr_score_ce_fwd(const float* H, const float* W, const float* bias, const float* zpad, const int64_t* gt,
                       int64_t M, int64_t N, int d, float* Z, float* lse, float* loss_row, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    if (M <= 0) return C2DSR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ldz = c2dsr_score_ldz(N);
    RUN(gemm_dispatch(0, 1, M, N, d, 1.f, H, d, W, d, 0.f, Z, ldz, bias, 0, make_dropout(0.f, 0, 0), workspace,
                      workspace_bytes, st));
    ce_row_kernel<<<(unsigned)M, 256, 0, st>>>(Z, ldz, zpad, gt, N, lse, loss_row);
    note_launches(1);
    return check_launch("score_ce_fwd");
}

int c2dsr_score_ce_bwd(const float* H, const float* W, const float* zpad, const int64_t* gt, const float* lse,
                       const float* coef, int64_t M, int64_t N, int d, float* Z, float* dH, float* dW,
                       float* dbias, float* dzpad, void* workspace, int64_t workspace_bytes, void* stream) {
    if (M <= 0) return C2DSR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ldz = c2dsr_score_ldz(N);
    const Dropout none = make_dropout(0.f, 0, 0);
    ce_grad_kernel<<<(unsigned)M, 256, 0, st>>>(Z, ldz, zpad, gt, lse, coef, N, dzpad);
    note_launches(1);
    RUN(gemm_dispatch(0, 0, M, d, N, 1.f, Z, ldz, W, d, 0.f, dH, d, nullptr, 0, none, workspace, workspace_bytes, st));
    RUN(gemm_dispatch(1, 0, N, d, M, 1.f, Z, ldz, H, d, 1.f, dW, d, nullptr, 0, none, workspace, workspace_bytes, st));
    RUN(c2dsr_colsum(Z, ldz, M, N, dbias, 1, workspace, workspace_bytes, st));
    return check_launch("score_ce_bwd");
}

int c2dsr_score_shard(const float* Q, const float* W, const float* bias, int64_t n_q, int64_t n_shard, int d,
                      float* S, int64_t lds, void* workspace, int64_t workspace_bytes, void* stream) {
    if (n_q <= 0 || n_shard <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(lds >= n_shard, "lds must be >= n_shard");
    return gemm_dispatch(0, 1, n_q, n_shard, d, 1.f, Q, d, W, d, 0.f, S, lds, bias, 0, make_dropout(0.f, 0, 0),
                         workspace, workspace_bytes, (cudaStream_t)stream);
}

int c2dsr_pick_target(const float* S, int64_t lds, const int64_t* gt, int64_t n_q, int64_t n0, int64_t n1,
                      float* s_gt, void* stream) {
    if (n_q <= 0) return C2DSR_OK;
    pick_target_kernel<<<(unsigned)ceil_div(n_q, 256), 256, 0, (cudaStream_t)stream>>>(S, lds, gt, n_q, n0, n1, s_gt);
    note_launches(1);
    return check_launch("pick_target");
}

int c2dsr_rank_from_scores(const float* S, int64_t lds, const float* s_gt, const int64_t* gt, const int64_t* neg,
                           int64_t n_neg, int64_t n_q, int64_t n0, int64_t n1, int32_t* counts, void* stream) {
    if (n_q <= 0) return C2DSR_OK;
    rank_count_kernel<<<(unsigned)n_q, 256, 0, (cudaStream_t)stream>>>(S, lds, s_gt, gt, neg, n_neg, n0, n1, counts);
    note_launches(1);
    return check_launch("rank_from_scores");
}

}  // extern "C"
