// K5: infomax contrastive discriminator (trainer.py:85-119, models/C2DSR.py:46-55 of the reference).
//   w_a = gt_mask_a / sum_L gt_mask_a,  w_b likewise                      (Trainer.cal_mask)
//   pooled slots: 0 x_mean = hx.w_a   1 y_mean = hy.w_b
//                 2 share.w_b (D_a positive)   3 neg_a.w_a (D_a negative)
//                 4 share.w_a (D_b positive)   5 neg_b.w_b (D_b negative)   -- the crossed masks of Q6
//   U[k] = pooled[2+k] W^T (W = D_a for k = 0,1; D_b for k = 2,3);  sim[k] = <x1, U[k]> + bias
//   loss = sum_k mean_b BCEwithLogits(sim[k], label_k), labels (1, 0, 1, 0)
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

int gemm_dispatch(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                  const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act,
                  Dropout dr, void* workspace, int64_t workspace_bytes, cudaStream_t st);

struct PoolSrc {
    const float* h[6];
    const int64_t* mask[6];
};

// grid (B, 6), block 128: pooled[slot][b][:] = sum_l h[b,l,:] * mask[b,l] / sum_l mask[b,l]
__global__ void pool_fwd_kernel(PoolSrc src, int L, int d, int64_t B, float* __restrict__ pooled) {
    const int64_t b = blockIdx.x;
    const int slot = blockIdx.y;
    const int64_t* mk = src.mask[slot] + b * L;
    float tot = 0.f;
    for (int l = 0; l < L; ++l) tot += (float)mk[l];
    const float* h = src.h[slot] + b * L * (int64_t)d;
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < L; ++l) {
            const float wgt = (float)mk[l] / tot;
            s += h[(int64_t)l * d + f] * wgt;
        }
        pooled[((int64_t)slot * B + b) * d + f] = s;
    }
}

// one warp per (k, b): sims[k][b] = <x1, U[k][b]> + bias
__global__ void sims_kernel(const float* __restrict__ pooled, const float* __restrict__ U,
                            const float* __restrict__ bias_a, const float* __restrict__ bias_b, int64_t B, int d,
                            float* __restrict__ sims) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= 4 * B) return;
    const int k = (int)(w / B);
    const int64_t b = w % B;
    const float* x1 = pooled + ((int64_t)(k < 2 ? 0 : 1) * B + b) * d;
    const float* u = U + ((int64_t)k * B + b) * d;
    float s = 0.f;
    for (int f = lane; f < d; f += 32) s += x1[f] * u[f];
    s = warp_sum(s);
    if (lane == 0) {
        const float* bias = k < 2 ? bias_a : bias_b;
        sims[w] = s + (bias ? bias[0] : 0.f);
    }
}

// single block: loss = inv_batch * sum_{k,b} bce(sims[k][b], label_k)
__global__ void bce_loss_kernel(const float* __restrict__ sims, int64_t B, float inv_batch, float* __restrict__ loss) {
    __shared__ float part[1024];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < 4 * B; i += blockDim.x) {
        const float z = sims[i];
        const float y = ((i / B) & 1) ? 0.f : 1.f;
        s += fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = part[0] * inv_batch;
}

// grid (B, 4), block 128: dU[k][b] = dsim * x1;  dpool[slot 0/1][b] = sum_k dsim * U[k][b];  dsim stored
__global__ void sims_bwd_kernel(const float* __restrict__ d_loss, const float* __restrict__ pooled,
                                const float* __restrict__ U, const float* __restrict__ sims, int64_t B, int d,
                                float inv_batch, float* __restrict__ dsim, float* __restrict__ dU,
                                float* __restrict__ dpool) {
    const int64_t b = blockIdx.x;
    const int k = blockIdx.y;
    const float up = d_loss[0] * inv_batch;
    const float z = sims[(int64_t)k * B + b];
    const float y = (k & 1) ? 0.f : 1.f;
    const float g = up * (1.f / (1.f + expf(-z)) - y);
    if (threadIdx.x == 0) dsim[(int64_t)k * B + b] = g;
    const float* x1 = pooled + ((int64_t)(k < 2 ? 0 : 1) * B + b) * d;
    for (int f = threadIdx.x; f < d; f += blockDim.x) dU[((int64_t)k * B + b) * d + f] = g * x1[f];
    if ((k & 1) == 0) {   // blocks k = 0 and k = 2 also produce d x_mean / d y_mean (fixed order k, k+1)
        const float z2 = sims[(int64_t)(k + 1) * B + b];
        const float g2 = up * (1.f / (1.f + expf(-z2)));
        const float* u1 = U + ((int64_t)k * B + b) * d;
        const float* u2 = U + ((int64_t)(k + 1) * B + b) * d;
        float* dp = dpool + ((int64_t)(k / 2) * B + b) * d;
        for (int f = threadIdx.x; f < d; f += blockDim.x) dp[f] = g * u1[f] + g2 * u2[f];
    }
}

struct UnpoolDst {
    float* d_h[5];              // share, hx, hy, neg_a, neg_b
};

// grid (B*L), block 128: d_h[b,l,:] = w[b,l] * dpool[slot][b,:] (+ second slot for h_share)
__global__ void unpool_kernel(UnpoolDst dst, const int64_t* __restrict__ mask_a, const int64_t* __restrict__ mask_b,
                              const float* __restrict__ dpool, int64_t B, int L, int d) {
    const int64_t bl = blockIdx.x;
    const int64_t b = bl / L;
    float ta = 0.f, tb = 0.f;
    for (int l = 0; l < L; ++l) {
        ta += (float)mask_a[b * L + l];
        tb += (float)mask_b[b * L + l];
    }
    const float wa = (float)mask_a[bl] / ta, wb = (float)mask_b[bl] / tb;
    auto P = [&](int slot) { return dpool + ((int64_t)slot * B + b) * d; };
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
        const int64_t o = bl * d + f;
        dst.d_h[0][o] = wb * P(2)[f] + wa * P(4)[f];
        dst.d_h[1][o] = wa * P(0)[f];
        dst.d_h[2][o] = wb * P(1)[f];
        dst.d_h[3][o] = wa * P(3)[f];
        dst.d_h[4][o] = wb * P(5)[f];
    }
}

__global__ void bias_grad_kernel(const float* __restrict__ dsim, int64_t B, float* dbias_a, float* dbias_b) {
    // single thread per discriminator, fixed order
    if (threadIdx.x < 2) {
        float* out = threadIdx.x == 0 ? dbias_a : dbias_b;
        if (!out) return;
        float s = 0.f;
        for (int64_t i = 0; i < 2 * B; ++i) s += dsim[(int64_t)threadIdx.x * 2 * B + i];
        out[0] += s;
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

int64_t c2dsr_infomax_workspace_bytes(int64_t B, int d) {
    return (4 * B + 10 * B * (int64_t)d) * 4 + (8ll << 20) + 1024;
}

int c2dsr_infomax_fwd(const float* h_share, const float* hx, const float* hy, const float* h_neg_a,
                      const float* h_neg_b, const int64_t* gt_mask_a, const int64_t* gt_mask_b, const float* W_a,
                      const float* W_b, const float* bias_a, const float* bias_b, int64_t B, int L, int d,
                      float inv_batch, float* pooled, float* U, float* sims, float* loss, void* workspace,
                      int64_t workspace_bytes, void* stream) {
    if (B <= 0) return C2DSR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    PoolSrc src;
    const float* hs[6] = {hx, hy, h_share, h_neg_a, h_share, h_neg_b};
    const int64_t* ms[6] = {gt_mask_a, gt_mask_b, gt_mask_b, gt_mask_a, gt_mask_a, gt_mask_b};
    for (int i = 0; i < 6; ++i) {
        src.h[i] = hs[i];
        src.mask[i] = ms[i];
    }
    pool_fwd_kernel<<<dim3((unsigned)B, 6), 128, 0, st>>>(src, L, d, B, pooled);
    const Dropout none = make_dropout(0.f, 0, 0);
    const int64_t Bd = B * (int64_t)d;
    RUN(gemm_dispatch(0, 1, 2 * B, d, d, 1.f, pooled + 2 * Bd, d, W_a, d, 0.f, U, d, nullptr, 0, none, workspace,
                      workspace_bytes, st));
    RUN(gemm_dispatch(0, 1, 2 * B, d, d, 1.f, pooled + 4 * Bd, d, W_b, d, 0.f, U + 2 * Bd, d, nullptr, 0, none,
                      workspace, workspace_bytes, st));
    sims_kernel<<<(unsigned)ceil_div(4 * B, 8), 256, 0, st>>>(pooled, U, bias_a, bias_b, B, d, sims);
    bce_loss_kernel<<<1, 1024, 0, st>>>(sims, B, inv_batch, loss);
    note_launches(3);
    return check_launch("infomax_fwd");
}

int c2dsr_infomax_bwd(const float* d_loss, const float* pooled, const float* U, const float* sims,
                      const int64_t* gt_mask_a, const int64_t* gt_mask_b, const float* W_a, const float* W_b,
                      int64_t B, int L, int d, float inv_batch, float* d_h_share, float* d_hx, float* d_hy,
                      float* d_h_neg_a, float* d_h_neg_b, float* dW_a, float* dW_b, float* dbias_a, float* dbias_b,
                      void* workspace, int64_t workspace_bytes, void* stream) {
    if (B <= 0) return C2DSR_OK;
    if (workspace_bytes < c2dsr_infomax_workspace_bytes(B, d)) {
        set_error("infomax_bwd: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Bd = B * (int64_t)d;
    float* dsim = (float*)workspace;
    float* dU = dsim + 4 * B;
    float* dpool = dU + 4 * Bd;
    void* gws = dpool + 6 * Bd;
    const int64_t gws_bytes = 8ll << 20;
    const Dropout none = make_dropout(0.f, 0, 0);
    sims_bwd_kernel<<<dim3((unsigned)B, 4), 128, 0, st>>>(d_loss, pooled, U, sims, B, d, inv_batch, dsim, dU, dpool);
    // d pooled[2..5] = dU W ;  dW += dU^T pooled[2..5]
    RUN(gemm_dispatch(0, 0, 2 * B, d, d, 1.f, dU, d, W_a, d, 0.f, dpool + 2 * Bd, d, nullptr, 0, none, gws, gws_bytes, st));
    RUN(gemm_dispatch(0, 0, 2 * B, d, d, 1.f, dU + 2 * Bd, d, W_b, d, 0.f, dpool + 4 * Bd, d, nullptr, 0, none, gws, gws_bytes, st));
    RUN(gemm_dispatch(1, 0, d, d, 2 * B, 1.f, dU, d, pooled + 2 * Bd, d, 1.f, dW_a, d, nullptr, 0, none, gws, gws_bytes, st));
    RUN(gemm_dispatch(1, 0, d, d, 2 * B, 1.f, dU + 2 * Bd, d, pooled + 4 * Bd, d, 1.f, dW_b, d, nullptr, 0, none, gws, gws_bytes, st));
    if (dbias_a || dbias_b) bias_grad_kernel<<<1, 32, 0, st>>>(dsim, B, dbias_a, dbias_b);
    UnpoolDst dst;
    dst.d_h[0] = d_h_share; dst.d_h[1] = d_hx; dst.d_h[2] = d_hy; dst.d_h[3] = d_h_neg_a; dst.d_h[4] = d_h_neg_b;
    unpool_kernel<<<(unsigned)(B * L), 128, 0, st>>>(dst, gt_mask_a, gt_mask_b, dpool, B, L, d);
    note_launches((dbias_a || dbias_b) ? 3 : 2);
    return check_launch("infomax_bwd");
}

}  // extern "C"
