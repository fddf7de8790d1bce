// Row assembly of the recommendation loss of ONE domain (trainer.py:122-152 of the reference):
//   virtual row v in [0, BR):     share term   H = h_share[b, l]              Hpad = h_share[b, l]   gt = gt_share[b, l]
//   virtual row v in [BR, 2BR):   domain term  H = h_share[b, l] + h_dom[b, l] Hpad = h_dom[b, l]     gt = gt_dom[b, l]
// with (b, l) = (j / R, L - R + j % R), j = v mod BR (the last R positions of every sequence).
// The reference builds these with slices, cats, an add and (here) a gather of the rows that carry a target --
// about twenty small launches per domain forward and as many backward.  Here: one stable partition of the
// virtual rows (valid targets first), one kernel that emits the first M rows (H, target, weight and the pad
// logit Hpad . w_pad + b_pad), and one backward kernel that writes d h_share and d h_dom for every token
// through the inverse permutation (each virtual row lands in at most one output row: no atomics).
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

struct RowGeom {
    int64_t B;
    int L, R, d;
    int64_t BR;
};

__device__ __forceinline__ int64_t token_of(const RowGeom& g, int64_t j) {       // j in [0, BR) -> b * L + l
    const int64_t b = j / g.R;
    return b * g.L + (g.L - g.R) + (j - b * g.R);
}

// One block: stable partition of the 2BR virtual rows; perm[pos] = v, inv[v] = pos.
__global__ void loss_rows_perm_kernel(const int64_t* __restrict__ gt_share, const int64_t* __restrict__ gt_dom, RowGeom g,
                                      int64_t ignore, int64_t* __restrict__ perm, int32_t* __restrict__ inv) {
    __shared__ int counts[1024];
    const int tid = threadIdx.x;
    const int64_t n = 2 * g.BR;
    const int64_t per = (n + blockDim.x - 1) / blockDim.x;
    const int64_t lo = tid * per, hi = lo + per < n ? lo + per : n;
    auto target = [&](int64_t v) {
        const int64_t t = token_of(g, v < g.BR ? v : v - g.BR);
        return v < g.BR ? gt_share[t] : gt_dom[t];
    };
    int c = 0;
    for (int64_t v = lo; v < hi; ++v) c += target(v) != ignore;
    counts[tid] = c;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
        const int x = tid >= o ? counts[tid - o] : 0;
        __syncthreads();
        counts[tid] += x;
        __syncthreads();
    }
    const int total = counts[blockDim.x - 1];
    int64_t v_pos = counts[tid] - c;
    int64_t i_pos = total + (lo < n ? lo : n) - v_pos;
    for (int64_t v = lo; v < hi; ++v) {
        const int64_t pos = target(v) != ignore ? v_pos++ : i_pos++;
        perm[pos] = v;
        inv[v] = (int32_t)pos;
    }
}

// warp per emitted row
__global__ void loss_rows_fwd_kernel(const float* __restrict__ h_share, const float* __restrict__ h_dom,
                                     const int64_t* __restrict__ gt_share, const int64_t* __restrict__ gt_dom,
                                     const int64_t* __restrict__ perm, RowGeom g, int64_t M,
                                     const float* __restrict__ w_share, const float* __restrict__ n_dom,
                                     const float* __restrict__ wpad, const float* __restrict__ bpad,
                                     float* __restrict__ H, int64_t* __restrict__ gt, float* __restrict__ w,
                                     float* __restrict__ zpad) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const int64_t v = perm[r];
    const bool share = v < g.BR;
    const int64_t t = token_of(g, share ? v : v - g.BR);
    const float* hs = h_share + t * g.d;
    const float* hd = h_dom + t * g.d;
    float dot = 0.f;
    for (int e = lane; e < g.d; e += 32) {
        const float a = hs[e];
        const float b = share ? 0.f : hd[e];
        H[r * g.d + e] = share ? a : a + b;
        dot += (share ? a : b) * wpad[e];
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        zpad[r] = dot + bpad[0];
        gt[r] = share ? gt_share[t] : gt_dom[t];
        const float nd = n_dom[0];
        w[r] = share ? w_share[0] : (nd > 0.f ? 1.f / nd : 0.f);
    }
}

// warp per token (b, l): gradients of h_share and h_dom from dH [M, d] and the pad-logit gradient dzpad [M]
__global__ void loss_rows_bwd_kernel(const float* __restrict__ dH, const float* __restrict__ dzpad,
                                     const float* __restrict__ wpad, const int32_t* __restrict__ inv, RowGeom g,
                                     int64_t M, float* __restrict__ d_share, float* __restrict__ d_dom) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= g.B * g.L) return;
    const int64_t b = t / g.L;
    const int l = (int)(t - b * g.L);
    float* ds = d_share + t * g.d;
    float* dd = d_dom + t * g.d;
    if (l < g.L - g.R) {
        for (int e = lane; e < g.d; e += 32) ds[e] = dd[e] = 0.f;
        return;
    }
    const int64_t j = b * g.R + (l - (g.L - g.R));
    const int64_t r1 = inv[j], r2 = inv[g.BR + j];
    const bool s1 = r1 < M, s2 = r2 < M;
    const float z1 = s1 ? dzpad[r1] : 0.f, z2 = s2 ? dzpad[r2] : 0.f;
    for (int e = lane; e < g.d; e += 32) {
        const float wp = wpad[e];
        const float a = s1 ? dH[r1 * g.d + e] + z1 * wp : 0.f;      // share row: H = Hpad = h_share
        const float c = s2 ? dH[r2 * g.d + e] : 0.f;                 // domain row: H = h_share + h_dom
        ds[e] = a + c;
        dd[e] = s2 ? c + z2 * wp : 0.f;                              //             Hpad = h_dom
    }
}

// d w_pad = sum_r dzpad[r] * Hpad[r], d b_pad = sum_r dzpad[r]: chunk partials (fixed order), then a final sum
constexpr int kPadChunk = 16;     // short serial chains (each row costs a dependent perm -> row load)
__global__ void loss_rows_wpad_partial_kernel(const float* __restrict__ h_share, const float* __restrict__ h_dom,
                                              const float* __restrict__ dzpad, const int64_t* __restrict__ perm, RowGeom g,
                                              int64_t M, float* __restrict__ partial) {
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * kPadChunk, r1 = r0 + kPadChunk < M ? r0 + kPadChunk : M;
    float s = 0.f, sb = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t v = perm[r];
        const bool share = v < g.BR;
        const int64_t t = token_of(g, share ? v : v - g.BR);
        const float z = dzpad[r];
        if (e < g.d) s += z * (share ? h_share : h_dom)[t * g.d + e];
        sb += z;
    }
    if (e < g.d) partial[(int64_t)blockIdx.x * (g.d + 1) + e] = s;
    if (e == 0) partial[(int64_t)blockIdx.x * (g.d + 1) + g.d] = sb;
}
__global__ void loss_rows_wpad_final_kernel(const float* __restrict__ partial, int64_t n_chunks, int d,
                                            float* __restrict__ dwpad, float* __restrict__ dbpad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > d) return;
    float s = 0.f;
    for (int64_t k = 0; k < n_chunks; ++k) s += partial[k * (d + 1) + e];
    if (e < d) dwpad[e] = s;
    else dbpad[0] = s;
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int c2dsr_loss_rows_fwd(const float* h_share, const float* h_dom, const int64_t* gt_share, const int64_t* gt_dom,
                        int64_t B, int L, int R, int d, int64_t ignore, int64_t M, const float* w_share,
                        const float* n_dom, const float* wpad, const float* bpad, int64_t* perm, int32_t* inv, float* H,
                        int64_t* gt, float* w, float* zpad, void* stream) {
    C2DSR_REQUIRE(B > 0 && L > 0 && R > 0 && R <= L && d > 0, "bad shape");
    C2DSR_REQUIRE(M >= 0 && M <= 2 * B * R && 2 * B * R < (1ll << 31), "M must be in [0, 2 B R]");
    cudaStream_t st = (cudaStream_t)stream;
    const RowGeom g{B, L, R, d, B * R};
    loss_rows_perm_kernel<<<1, 1024, 0, st>>>(gt_share, gt_dom, g, ignore, perm, inv);
    if (M > 0)
        loss_rows_fwd_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(h_share, h_dom, gt_share, gt_dom, perm, g, M, w_share,
                                                                      n_dom, wpad, bpad, H, gt, w, zpad);
    note_launches(M > 0 ? 2 : 1);
    return check_launch("loss_rows_fwd");
}

int64_t c2dsr_loss_rows_bwd_workspace_bytes(int64_t M, int d) {
    return ceil_div(M > 0 ? M : 1, kPadChunk) * (int64_t)(d + 1) * 4 + 256;
}

int c2dsr_loss_rows_bwd(const float* dH, const float* dzpad, const float* h_share, const float* h_dom, const float* wpad,
                        const int64_t* perm, const int32_t* inv, int64_t B, int L, int R, int d, int64_t M,
                        float* d_h_share, float* d_h_dom, float* dwpad, float* dbpad, void* workspace,
                        int64_t workspace_bytes, void* stream) {
    C2DSR_REQUIRE(B > 0 && L > 0 && R > 0 && R <= L && d > 0 && M >= 0 && M <= 2 * B * R, "bad shape");
    if (workspace_bytes < c2dsr_loss_rows_bwd_workspace_bytes(M, d)) {
        set_error("loss_rows_bwd: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const RowGeom g{B, L, R, d, B * R};
    loss_rows_bwd_kernel<<<(unsigned)ceil_div(B * L, 8), 256, 0, st>>>(dH, dzpad, wpad, inv, g, M, d_h_share, d_h_dom);
    const int64_t n_chunks = ceil_div(M > 0 ? M : 1, kPadChunk);
    float* partial = (float*)workspace;
    if (M > 0)
        loss_rows_wpad_partial_kernel<<<dim3((unsigned)n_chunks, (unsigned)ceil_div(d, 128)), 128, 0, st>>>(
            h_share, h_dom, dzpad, perm, g, M, partial);
    loss_rows_wpad_final_kernel<<<(unsigned)ceil_div(d + 1, 128), 128, 0, st>>>(partial, M > 0 ? n_chunks : 0, d, dwpad,
                                                                                dbpad);
    note_launches(M > 0 ? 3 : 2);
    return check_launch("loss_rows_bwd");
}

}  // extern "C"
