// K3: the SASRec-style encoder (models/encoders.py:23-33 of the reference) -- attention with the
// reference's mask semantics, LayerNorm, residual/dropout glue, and the composite forward/backward
// that chains them with the dense layers (gemm.cu).
//
// Attention semantics: query i attends key j iff j <= i AND seq[j] == PAD.  Rows
// with no allowed key produce 0 and receive/propagate no gradient.  Sequences are short (L <= 200), so
// attention is one warp per (sequence, head, query): lanes split the head dimension, keys are walked
// in order with an online softmax, non-pad keys are skipped without touching memory.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

int gemm_dispatch(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                  const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act,
                  Dropout dr, void* workspace, int64_t workspace_bytes, cudaStream_t st);

int gemm_tc_dispatch(int ta, int tb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act, Dropout dr, int passes,
                     void* workspace, int64_t workspace_bytes, cudaStream_t st);
int64_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);

// dense layer product: tcgen05 path (passes = 1 or 3) or the fp32 FFMA kernel (passes = 0); alpha is always 1
static int dense(int passes, void* tws, int64_t tws_bytes, int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha,
                 const float* A, int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                 const float* bias, int act, Dropout dr, void* gws, int64_t gws_bytes, cudaStream_t st) {
    if (passes == 0) return gemm_dispatch(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, dr, gws, gws_bytes, st);
    return gemm_tc_dispatch(ta, tb, M, N, K, A, lda, B, ldb, beta, C, ldc, bias, act, dr, passes, tws, tws_bytes, st);
}

// scratch of the tcgen05 dense path: the largest of the five products of a layer
static int64_t dense_tc_bytes(int64_t T, int d) {
    int64_t b = gemm_tc_workspace_bytes(T, 3 * d, d);
    const int64_t shapes[4][3] = {{T, d, d}, {T, d, 3 * d}, {3 * d, d, T}, {d, d, T}};
    for (auto& s : shapes) {
        const int64_t x = gemm_tc_workspace_bytes(s[0], s[1], s[2]);
        if (x > b) b = x;
    }
    return b;
}

constexpr int kMaxPerLane = 16;   // features per lane: d (or head dim) <= 512
constexpr int64_t kGemmWsBytes = 16ll << 20;   // split-K partials / column-reduction partials
constexpr int kLnChunk = 128;

int colsum_dispatch(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, int accumulate, void* ws,
                    int64_t ws_bytes, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// LayerNorm family: one warp per token, lane l owns features l, l+32, ...
// ------------------------------------------------------------------------------------------------
__global__ void add_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                  const float* __restrict__ w, const float* __restrict__ b, float* s_out,
                                  float* out, float* __restrict__ stats, int64_t n_tok, int d, int do_ln, float eps,
                                  Dropout dr) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tok) return;
    const int nper = (d + 31) >> 5;
    float v[kMaxPerLane];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int f = lane + 32 * k;
        v[k] = 0.f;
        if (k < nper && f < d) {
            float s = x[t * d + f];
            if (y) s += y[t * d + f] * drop_scale(dr, (uint64_t)t * d + f);
            v[k] = s;
            sum += s;
            if (s_out) s_out[t * d + f] = s;
        }
    }
    if (!do_ln) {
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k) {
            const int f = lane + 32 * k;
            if (k < nper && f < d) out[t * d + f] = v[k];
        }
        return;
    }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int f = lane + 32 * k;
        if (k < nper && f < d) {
            const float c = v[k] - mean;
            sq += c * c;
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)d + eps);
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int f = lane + 32 * k;
        if (k < nper && f < d) out[t * d + f] = (v[k] - mean) * rstd * w[f] + b[f];
    }
    if (lane == 0 && stats) {
        stats[2 * t] = mean;
        stats[2 * t + 1] = rstd;
    }
}

__global__ void ln_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ s,
                              const float* __restrict__ stats, const float* __restrict__ w, float* dx_out,
                              int accumulate, int64_t n_tok, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tok) return;
    const int nper = (d + 31) >> 5;
    const float mean = stats[2 * t], rstd = stats[2 * t + 1];
    float g[kMaxPerLane], xh[kMaxPerLane];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int f = lane + 32 * k;
        g[k] = 0.f;
        xh[k] = 0.f;
        if (k < nper && f < d) {
            xh[k] = (s[t * d + f] - mean) * rstd;
            g[k] = d_out[t * d + f] * w[f];
            c1 += g[k];
            c2 += g[k] * xh[k];
        }
    }
    c1 = warp_sum(c1) / (float)d;
    c2 = warp_sum(c2) / (float)d;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int f = lane + 32 * k;
        if (k < nper && f < d) {
            const float ds = rstd * (g[k] - c1 - xh[k] * c2);
            dx_out[t * d + f] = accumulate ? dx_out[t * d + f] + ds : ds;
        }
    }
}

// d_w[f] += sum_t d_out[t,f] * xhat[t,f];  d_b[f] += sum_t d_out[t,f].  Fixed order (see colsum_kernel).
__global__ void ln_param_grad_kernel(const float* __restrict__ d_out, const float* __restrict__ s,
                                     const float* __restrict__ stats, float* d_w, float* d_b, int64_t n_tok, int d) {
    __shared__ float pw[8][33], pb[8][33];
    const int f = blockIdx.x * 32 + threadIdx.x;
    float sw = 0.f, sb = 0.f;
    if (f < d) {
        for (int64_t t = threadIdx.y; t < n_tok; t += 8) {
            const float go = d_out[t * d + f];
            sw += go * (s[t * d + f] - stats[2 * t]) * stats[2 * t + 1];
            sb += go;
        }
    }
    pw[threadIdx.y][threadIdx.x] = sw;
    pb[threadIdx.y][threadIdx.x] = sb;
    __syncthreads();
    if (threadIdx.y == 0 && f < d) {
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a += pw[k][threadIdx.x];
            c += pb[k][threadIdx.x];
        }
        d_w[f] += a;
        d_b[f] += c;
    }
}

// two-phase form of the above: grid (feature blocks, token chunks) -> partials [chunk][2][d], summed in chunk order
__global__ void ln_param_partial_kernel(const float* __restrict__ d_out, const float* __restrict__ s,
                                        const float* __restrict__ stats, float* __restrict__ part, int64_t n_tok,
                                        int d) {
    __shared__ float pw[8][33], pb[8][33];
    const int f = blockIdx.x * 32 + threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.y * kLnChunk;
    const int64_t t1 = t0 + kLnChunk < n_tok ? t0 + kLnChunk : n_tok;
    float sw = 0.f, sb = 0.f;
    if (f < d) {
#pragma unroll 4
        for (int64_t t = t0 + threadIdx.y; t < t1; t += 8) {
            const float go = d_out[t * d + f];
            sw += go * (s[t * d + f] - stats[2 * t]) * stats[2 * t + 1];
            sb += go;
        }
    }
    pw[threadIdx.y][threadIdx.x] = sw;
    pb[threadIdx.y][threadIdx.x] = sb;
    __syncthreads();
    if (threadIdx.y == 0 && f < d) {
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a += pw[k][threadIdx.x];
            c += pb[k][threadIdx.x];
        }
        part[((int64_t)blockIdx.y * 2) * d + f] = a;
        part[((int64_t)blockIdx.y * 2 + 1) * d + f] = c;
    }
}
// block (32, 8): group y adds chunks y, y + 8, ...; the 8 group sums are added in group order (fixed order)
__global__ void ln_param_final_kernel(const float* __restrict__ part, int64_t n_chunks, int d, float* d_w, float* d_b) {
    __shared__ float pw[8][33], pb[8][33];
    const int f = blockIdx.x * 32 + threadIdx.x;
    float a = 0.f, c = 0.f;
    if (f < d) {
#pragma unroll 4
        for (int64_t k = threadIdx.y; k < n_chunks; k += 8) {
            a += part[(k * 2) * d + f];
            c += part[(k * 2 + 1) * d + f];
        }
    }
    pw[threadIdx.y][threadIdx.x] = a;
    pb[threadIdx.y][threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.y == 0 && f < d) {
        float ta = 0.f, tc = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            ta += pw[g][threadIdx.x];
            tc += pb[g][threadIdx.x];
        }
        d_w[f] += ta;
        d_b[f] += tc;
    }
}

// out = in * dropout mask (same index space as the forward site)
__global__ void drop_mul_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, Dropout dr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i] * drop_scale(dr, (uint64_t)i);
}

// backward of fd = drop(relu(pre)):  d_pre = d_fd * (fd > 0 ? 1/(1-p) : 0), in place on d_fd
__global__ void relu_drop_bwd_kernel(float* __restrict__ d_fd, const float* __restrict__ fd, int64_t n,
                                     float inv_keep) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d_fd[i] = fd[i] > 0.f ? d_fd[i] * inv_keep : 0.f;
}

// ------------------------------------------------------------------------------------------------
// attention
// ------------------------------------------------------------------------------------------------
struct AttnShape {
    int64_t n_seq;
    int L, d, H, dh;
    int64_t pad;
    float scale;
};

__device__ __forceinline__ float lane_dot(const float (&a)[kMaxPerLane], const float* __restrict__ b, int dh, int nper,
                                          int lane) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int e = lane + 32 * k;
        if (k < nper && e < dh) s += a[k] * b[e];
    }
    return warp_sum(s);
}

__global__ void attn_fwd_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ seq, AttnShape sh,
                                Dropout dr, float* __restrict__ o, float* __restrict__ lse) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= sh.n_seq * sh.H * sh.L) return;
    const int i = (int)(w % sh.L);
    const int h = (int)((w / sh.L) % sh.H);
    const int64_t b = w / ((int64_t)sh.L * sh.H);
    const int nper = (sh.dh + 31) >> 5;
    const int64_t ld = 3 * (int64_t)sh.d;
    const int64_t ti = b * sh.L + i;
    float q[kMaxPerLane], acc[kMaxPerLane];
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int e = lane + 32 * k;
        q[k] = (k < nper && e < sh.dh) ? qkv[ti * ld + h * sh.dh + e] : 0.f;
        acc[k] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j <= i; ++j) {
        const int64_t tj = b * sh.L + j;
        if (seq[tj] != sh.pad) continue;                                   // warp-uniform
        const float* kj = qkv + tj * ld + sh.d + h * sh.dh;
        const float* vj = kj + sh.d;
        const float s = lane_dot(q, kj, sh.dh, nper, lane) * sh.scale;
        const float m_new = fmaxf(m, s);
        const float corr = expf(m - m_new);                                // exp(-inf) = 0 on the first key
        const float pe = expf(s - m_new);
        l = l * corr + pe;
        const float keep = drop_scale(dr, (uint64_t)((b * sh.H + h) * sh.L + i) * sh.L + j);
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k) {
            const int e = lane + 32 * k;
            if (k < nper && e < sh.dh) acc[k] = acc[k] * corr + pe * keep * vj[e];
        }
        m = m_new;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int e = lane + 32 * k;
        if (k < nper && e < sh.dh) o[ti * sh.d + h * sh.dh + e] = acc[k] * inv;
    }
    if (lane == 0) lse[ti * sh.H + h] = l > 0.f ? m + logf(l) : INFINITY;  // +inf => p = exp(s - lse) = 0
}

// First half of the warps: dq for one query.  Second half: dk, dv for one key.  No atomics.
__global__ void attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ o,
                                const float* __restrict__ lse, const float* __restrict__ d_o,
                                const int64_t* __restrict__ seq, AttnShape sh, Dropout dr, float* __restrict__ d_qkv) {
    const int lane = threadIdx.x & 31;
    int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_q = sh.n_seq * sh.H * sh.L;
    if (w >= 2 * n_q) return;
    const bool key_role = w >= n_q;
    if (key_role) w -= n_q;
    const int r = (int)(w % sh.L);                        // query index i or key index j
    const int h = (int)((w / sh.L) % sh.H);
    const int64_t b = w / ((int64_t)sh.L * sh.H);
    const int nper = (sh.dh + 31) >> 5;
    const int64_t ld = 3 * (int64_t)sh.d;
    const int64_t tr = b * sh.L + r;
    const int hoff = h * sh.dh;
    float a0[kMaxPerLane], a1[kMaxPerLane], g0[kMaxPerLane], g1[kMaxPerLane];

    if (!key_role) {
        // a0 = q_i, a1 = d_o_i, g0 = dq accumulator
        float dsum = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k) {
            const int e = lane + 32 * k;
            const bool ok = k < nper && e < sh.dh;
            a0[k] = ok ? qkv[tr * ld + hoff + e] : 0.f;
            a1[k] = ok ? d_o[tr * sh.d + hoff + e] : 0.f;
            g0[k] = 0.f;
            if (ok) dsum += a1[k] * o[tr * sh.d + hoff + e];
        }
        const float D = warp_sum(dsum);
        const float lse_i = lse[tr * sh.H + h];
        for (int j = 0; j <= r; ++j) {
            const int64_t tj = b * sh.L + j;
            if (seq[tj] != sh.pad) continue;
            const float* kj = qkv + tj * ld + sh.d + hoff;
            const float* vj = kj + sh.d;
            const float s = lane_dot(a0, kj, sh.dh, nper, lane) * sh.scale;
            const float p = expf(s - lse_i);
            const float keep = drop_scale(dr, (uint64_t)((b * sh.H + h) * sh.L + r) * sh.L + j);
            const float dP = lane_dot(a1, vj, sh.dh, nper, lane) * keep;
            const float dS = p * (dP - D) * sh.scale;
#pragma unroll
            for (int k = 0; k < kMaxPerLane; ++k) {
                const int e = lane + 32 * k;
                if (k < nper && e < sh.dh) g0[k] += dS * kj[e];
            }
        }
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k) {
            const int e = lane + 32 * k;
            if (k < nper && e < sh.dh) d_qkv[tr * ld + hoff + e] = g0[k];
        }
        return;
    }

    // key role: a0 = k_j, a1 = v_j, g0 = dk, g1 = dv
    const bool allowed = seq[tr] == sh.pad;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int e = lane + 32 * k;
        const bool ok = allowed && k < nper && e < sh.dh;
        a0[k] = ok ? qkv[tr * ld + sh.d + hoff + e] : 0.f;
        a1[k] = ok ? qkv[tr * ld + 2 * sh.d + hoff + e] : 0.f;
        g0[k] = 0.f;
        g1[k] = 0.f;
    }
    if (allowed) {
        for (int i = r; i < sh.L; ++i) {
            const int64_t ti = b * sh.L + i;
            const float* qi = qkv + ti * ld + hoff;
            const float* doi = d_o + ti * sh.d + hoff;
            const float* oi = o + ti * sh.d + hoff;
            const float lse_i = lse[ti * sh.H + h];
            const float s = lane_dot(a0, qi, sh.dh, nper, lane) * sh.scale;
            const float p = expf(s - lse_i);
            const float keep = drop_scale(dr, (uint64_t)((b * sh.H + h) * sh.L + i) * sh.L + r);
            float dd = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxPerLane; ++k) {
                const int e = lane + 32 * k;
                if (k < nper && e < sh.dh) dd += doi[e] * oi[e];
            }
            const float D = warp_sum(dd);
            const float dP = lane_dot(a1, doi, sh.dh, nper, lane) * keep;
            const float dS = p * (dP - D) * sh.scale;
            const float pk = p * keep;
#pragma unroll
            for (int k = 0; k < kMaxPerLane; ++k) {
                const int e = lane + 32 * k;
                if (k < nper && e < sh.dh) {
                    g0[k] += dS * qi[e];
                    g1[k] += pk * doi[e];
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int e = lane + 32 * k;
        if (k < nper && e < sh.dh) {
            d_qkv[tr * ld + sh.d + hoff + e] = g0[k];
            d_qkv[tr * ld + 2 * sh.d + hoff + e] = g1[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers shared by the C entry points and the composite
// ------------------------------------------------------------------------------------------------
static int launch_add_ln(const float* x, const float* y, const float* w, const float* b, float* s_out, float* out,
                         float* stats, int64_t n_tok, int d, int do_ln, float eps, Dropout dr, cudaStream_t st) {
    add_ln_fwd_kernel<<<(unsigned)ceil_div(n_tok, 8), 256, 0, st>>>(x, y, w, b, s_out, out, stats, n_tok, d, do_ln,
                                                                     eps, dr);
    note_launches(1);
    return check_launch("add_ln_fwd");
}
static int launch_ln_bwd(const float* d_out, const float* s, const float* stats, const float* w, float* dx_out,
                         int accumulate, int64_t n_tok, int d, cudaStream_t st) {
    ln_bwd_kernel<<<(unsigned)ceil_div(n_tok, 8), 256, 0, st>>>(d_out, s, stats, w, dx_out, accumulate, n_tok, d);
    note_launches(1);
    return check_launch("ln_bwd");
}
static int launch_ln_param(void* ws, const float* d_out, const float* s, const float* stats, float* d_w, float* d_b,
                           int64_t n_tok, int d, cudaStream_t st) {
    const int64_t n_chunks = ceil_div(n_tok, kLnChunk);
    if (ws && n_chunks > 1 && 2 * n_chunks * d * 4 <= kGemmWsBytes) {
        float* part = (float*)ws;
        ln_param_partial_kernel<<<dim3((unsigned)ceil_div(d, 32), (unsigned)n_chunks), dim3(32, 8), 0, st>>>(
            d_out, s, stats, part, n_tok, d);
        ln_param_final_kernel<<<(unsigned)ceil_div(d, 32), dim3(32, 8), 0, st>>>(part, n_chunks, d, d_w, d_b);
        note_launches(2);
    } else {
        ln_param_grad_kernel<<<(unsigned)ceil_div(d, 32), dim3(32, 8), 0, st>>>(d_out, s, stats, d_w, d_b, n_tok, d);
        note_launches(1);
    }
    return check_launch("ln_param_grad");
}
static AttnShape make_shape(int64_t n_seq, int L, int d, int H, int64_t pad) {
    AttnShape sh{n_seq, L, d, H, d / H, pad, 1.f / sqrtf((float)(d / H))};
    return sh;
}
// ---- one CTA per (sequence, head): Q, K, V (and dO) staged in shared memory ---------------------------------
// Short sequences (the staged tiles fit in kSeqSmemLimit and head_dim % 4 == 0) take this path.  Every global
// element is read once with 16-byte accesses; the work is cut into three block-wide phases so that no
// phase has a serial dependency longer than one dot product:
//   1. all (query, key) scores at once, one warp per pair (16-byte shared loads, one shuffle reduction);
//   2. soft-max per query row (one warp per row) -> probabilities (with the dropout mask folded in);
//   3. outputs as small dense products, one thread per float4 of output, coefficients broadcast.
// Sums over keys / queries run in index order, so results are reproducible.
constexpr int kSeqThreads = 256;
constexpr int kSeqSmemLimit = 160 * 1024;

__device__ __forceinline__ float smem_dot4(const float* a, const float* b, int dh, int lane) {
    float s = 0.f;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int e = lane; e < (dh >> 2); e += 32) {
        const float4 x = a4[e], y = b4[e];
        s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    return warp_sum(s);
}

// copy rows [L, dh] of one head out of a [*, ld] global matrix into contiguous shared memory
__device__ __forceinline__ void stage_rows(float* dst, const float* src, int64_t ld, int L, int dh) {
    const int q = dh >> 2;
    for (int idx = threadIdx.x; idx < L * q; idx += blockDim.x) {
        const int i = idx / q, e = idx - i * q;
        reinterpret_cast<float4*>(dst)[idx] = __ldg(reinterpret_cast<const float4*>(src + i * ld) + e);
    }
}

__global__ void __launch_bounds__(kSeqThreads)
attn_seq_fwd_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ seq, AttnShape sh, Dropout dr,
                    float* __restrict__ o, float* __restrict__ lse) {
    extern __shared__ __align__(16) float sm[];
    const int L = sh.L, dh = sh.dh, LS = L + 1;
    float* Q = sm;
    float* K = Q + L * dh;
    float* V = K + L * dh;
    float* S = V + L * dh;                       // [L][L + 1] scores, then probabilities
    int* allow = reinterpret_cast<int*>(S + L * LS);
    const int64_t b = blockIdx.x / sh.H;
    const int h = blockIdx.x % sh.H;
    const int64_t ld = 3 * (int64_t)sh.d;
    const float* base = qkv + b * L * ld + h * dh;
    stage_rows(Q, base, ld, L, dh);
    stage_rows(K, base + sh.d, ld, L, dh);
    stage_rows(V, base + 2 * sh.d, ld, L, dh);
    for (int j = threadIdx.x; j < L; j += blockDim.x) allow[j] = seq[b * L + j] == sh.pad;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // 1. scores
    for (int p = warp; p < L * L; p += nw) {
        const int i = p / L, j = p - i * L;
        if (j > i || !allow[j]) continue;                                  // warp-uniform
        const float s = smem_dot4(Q + i * dh, K + j * dh, dh, lane) * sh.scale;
        if (lane == 0) S[i * LS + j] = s;
    }
    __syncthreads();
    // 2. soft-max rows: S <- p_ij * keep_ij, 0 for masked pairs; rows with no allowed key are all zero
    for (int i = warp; i < L; i += nw) {
        float m = -INFINITY;
        for (int j = lane; j <= i; j += 32)
            if (allow[j]) m = fmaxf(m, S[i * LS + j]);
        m = warp_max(m);
        float l = 0.f;
        for (int j = lane; j <= i; j += 32)
            if (allow[j]) l += expf(S[i * LS + j] - m);
        l = warp_sum(l);
        const float inv = l > 0.f ? 1.f / l : 0.f;
        for (int j = lane; j < L; j += 32) {
            float pr = 0.f;
            if (j <= i && allow[j])
                pr = expf(S[i * LS + j] - m) * inv * drop_scale(dr, (uint64_t)((b * sh.H + h) * L + i) * L + j);
            S[i * LS + j] = pr;
        }
        if (lane == 0) lse[(b * L + i) * sh.H + h] = l > 0.f ? m + logf(l) : INFINITY;
    }
    __syncthreads();
    // 3. O = P V
    const int q = dh >> 2;
    for (int idx = threadIdx.x; idx < L * q; idx += blockDim.x) {
        const int i = idx / q, e = idx - i * q;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j <= i; ++j) {
            const float pr = S[i * LS + j];
            if (pr != 0.f) {
                const float4 v = reinterpret_cast<const float4*>(V + j * dh)[e];
                acc.x += pr * v.x; acc.y += pr * v.y; acc.z += pr * v.z; acc.w += pr * v.w;
            }
        }
        reinterpret_cast<float4*>(o + (b * L + i) * sh.d + h * dh)[e] = acc;
    }
}

__global__ void __launch_bounds__(kSeqThreads)
attn_seq_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ lse,
                    const float* __restrict__ d_o, const int64_t* __restrict__ seq, AttnShape sh, Dropout dr,
                    float* __restrict__ d_qkv) {
    extern __shared__ __align__(16) float sm[];
    const int L = sh.L, dh = sh.dh, LS = L + 1;
    float* Q = sm;
    float* K = Q + L * dh;
    float* V = K + L * dh;
    float* dO = V + L * dh;
    float* Pk = dO + L * dh;           // [L][L + 1]  p_ij * keep_ij          (for dV)
    float* dS = Pk + L * LS;           // [L][L + 1]  dS_ij * scale           (for dQ, dK)
    float* D = dS + L * LS;            // [L]  dO_i . O_i
    int* allow = reinterpret_cast<int*>(D + L);
    const int64_t b = blockIdx.x / sh.H;
    const int h = blockIdx.x % sh.H;
    const int64_t ld = 3 * (int64_t)sh.d;
    const float* base = qkv + b * L * ld + h * dh;
    stage_rows(Q, base, ld, L, dh);
    stage_rows(K, base + sh.d, ld, L, dh);
    stage_rows(V, base + 2 * sh.d, ld, L, dh);
    stage_rows(dO, d_o + b * L * sh.d + h * dh, sh.d, L, dh);
    for (int j = threadIdx.x; j < L; j += blockDim.x) allow[j] = seq[b * L + j] == sh.pad;
    for (int idx = threadIdx.x; idx < 2 * L * LS; idx += blockDim.x) Pk[idx] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int q = dh >> 2;
    // 0. D_i = dO_i . O_i  (O straight from global memory: it is used nowhere else)
    for (int i = warp; i < L; i += nw) {
        const float4* o4 = reinterpret_cast<const float4*>(o + (b * L + i) * sh.d + h * dh);
        const float4* g4 = reinterpret_cast<const float4*>(dO + i * dh);
        float s = 0.f;
        for (int e = lane; e < q; e += 32) {
            const float4 x = __ldg(o4 + e), y = g4[e];
            s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        }
        s = warp_sum(s);
        if (lane == 0) D[i] = s;
    }
    __syncthreads();
    // 1. per (query, key) coefficients
    for (int p = warp; p < L * L; p += nw) {
        const int i = p / L, j = p - i * L;
        if (j > i || !allow[j]) continue;                                  // warp-uniform
        float s = 0.f, t = 0.f;
        const float4* q4 = reinterpret_cast<const float4*>(Q + i * dh);
        const float4* k4 = reinterpret_cast<const float4*>(K + j * dh);
        const float4* g4 = reinterpret_cast<const float4*>(dO + i * dh);
        const float4* v4 = reinterpret_cast<const float4*>(V + j * dh);
        for (int e = lane; e < q; e += 32) {
            const float4 a = q4[e], c = k4[e], g = g4[e], v = v4[e];
            s += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
            t += g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
        }
        s = warp_sum(s) * sh.scale;
        t = warp_sum(t);
        if (lane == 0) {
            const float pr = expf(s - lse[(b * L + i) * sh.H + h]);
            const float keep = drop_scale(dr, (uint64_t)((b * sh.H + h) * L + i) * L + j);
            Pk[i * LS + j] = pr * keep;
            dS[i * LS + j] = pr * (t * keep - D[i]) * sh.scale;
        }
    }
    __syncthreads();
    // 2. dQ_i = sum_j dS_ij K_j ;  dK_j = sum_i dS_ij Q_i ;  dV_j = sum_i Pk_ij dO_i
    for (int idx = threadIdx.x; idx < 3 * L * q; idx += blockDim.x) {
        const int which = idx / (L * q), rem = idx - which * (L * q);
        const int r = rem / q, e = rem - r * q;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (which == 0) {
            for (int j = 0; j <= r; ++j) {
                const float c = dS[r * LS + j];
                if (c != 0.f) {
                    const float4 x = reinterpret_cast<const float4*>(K + j * dh)[e];
                    acc.x += c * x.x; acc.y += c * x.y; acc.z += c * x.z; acc.w += c * x.w;
                }
            }
        } else {
            const float* coef = which == 1 ? dS : Pk;
            const float* src = which == 1 ? Q : dO;
            for (int i = r; i < L; ++i) {
                const float c = coef[i * LS + r];
                if (c != 0.f) {
                    const float4 x = reinterpret_cast<const float4*>(src + i * dh)[e];
                    acc.x += c * x.x; acc.y += c * x.y; acc.z += c * x.z; acc.w += c * x.w;
                }
            }
        }
        reinterpret_cast<float4*>(d_qkv + (b * L + r) * ld + which * sh.d + h * dh)[e] = acc;
    }
}

// ---- evaluation: one query position per sequence ------------------------------------------------------------
// Only h[b, sel[b]] is read from a branch at evaluation time (trainer.py:169-177).  In the LAST encoder layer every
// token still contributes its key and value, but the query projection is only needed at sel[b], and everything
// after the attention (output projection, residual + LayerNorm, feed-forward, LayerNorms) is per-token work that
// is only needed for that one token: n_seq rows instead of n_seq * L.
// warp per (sequence, head): soft-max attention of the single query sel[b] over the allowed keys j <= sel[b]
__global__ void attn_select_kernel(const float* __restrict__ qkv, const int64_t* __restrict__ seq,
                                   const int64_t* __restrict__ sel, AttnShape sh, float* __restrict__ o) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= sh.n_seq * sh.H) return;
    const int h = (int)(w % sh.H);
    const int64_t b = w / sh.H;
    const int i = (int)sel[b];
    const int nper = (sh.dh + 31) >> 5;
    const int64_t ld = 3 * (int64_t)sh.d;
    const int64_t ti = b * sh.L + i;
    float q[kMaxPerLane], acc[kMaxPerLane];
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int e = lane + 32 * k;
        q[k] = (k < nper && e < sh.dh) ? qkv[ti * ld + h * sh.dh + e] : 0.f;
        acc[k] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j <= i; ++j) {
        const int64_t tj = b * sh.L + j;
        if (seq[tj] != sh.pad) continue;                                   // warp-uniform
        const float* kj = qkv + tj * ld + sh.d + h * sh.dh;
        const float* vj = kj + sh.d;
        const float s = lane_dot(q, kj, sh.dh, nper, lane) * sh.scale;
        const float m_new = fmaxf(m, s);
        const float corr = expf(m - m_new);
        const float pe = expf(s - m_new);
        l = l * corr + pe;
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k) {
            const int e = lane + 32 * k;
            if (k < nper && e < sh.dh) acc[k] = acc[k] * corr + pe * vj[e];
        }
        m = m_new;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int e = lane + 32 * k;
        if (k < nper && e < sh.dh) o[b * sh.d + h * sh.dh + e] = acc[k] * inv;
    }
}

// out[b, :] = x[b, sel[b], :]   (warp per sequence)
__global__ void select_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ sel, int64_t n_seq, int L,
                                   int d, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n_seq) return;
    const float* src = x + (b * L + sel[b]) * d;
    for (int e = lane; e < d; e += 32) out[b * d + e] = src[e];
}

// Pad-key shortcut (evaluation, see c2dsr_encoder_fwd_padkeys): y[b, :] = attention block output of token sel[b]
// = y_pad when some key j <= sel[b] is a pad token, else the output-projection bias alone (no allowed key: the
// attention output is 0).  Warp per sequence.
__global__ void padkey_rows_kernel(const int64_t* __restrict__ seq, const int64_t* __restrict__ sel, int64_t n_seq,
                                   int L, int d, int64_t pad, const float* __restrict__ y_pad,
                                   const float* __restrict__ b_o, float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n_seq) return;
    const int i = (int)sel[b];
    int has = 0;
    for (int j = lane; j <= i; j += 32) has |= (seq[b * L + j] == pad) ? 1 : 0;
    has = __any_sync(0xffffffffu, has);
    const float* src = has ? y_pad : b_o;
    for (int e = lane; e < d; e += 32) y[b * d + e] = src[e];
}

static int seq_smem_bytes(const AttnShape& sh, bool bwd) {
    return ((bwd ? 4 : 3) * sh.L * sh.dh + (bwd ? 2 : 1) * sh.L * (sh.L + 1) + 2 * sh.L + 8) * 4;
}
static bool seq_path_ok(const AttnShape& sh, bool bwd) {
    return seq_smem_bytes(sh, bwd) <= kSeqSmemLimit && sh.dh % 4 == 0 && sh.d % 4 == 0;
}

static int launch_attn_fwd(const float* qkv, const int64_t* seq, AttnShape sh, Dropout dr, float* o, float* lse,
                           cudaStream_t st) {
    if (seq_path_ok(sh, false) && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(attn_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqSmemLimit);
            attr = true;
        }
        attn_seq_fwd_kernel<<<(unsigned)(sh.n_seq * sh.H), kSeqThreads, seq_smem_bytes(sh, false), st>>>(qkv, seq, sh, dr,
                                                                                                        o, lse);
    } else {
        const int64_t warps = sh.n_seq * sh.H * sh.L;
        attn_fwd_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>(qkv, seq, sh, dr, o, lse);
    }
    note_launches(1);
    return check_launch("attention_fwd");
}
static int launch_attn_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, const int64_t* seq,
                           AttnShape sh, Dropout dr, float* d_qkv, cudaStream_t st) {
    if (seq_path_ok(sh, true) && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o) |
                                   reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(d_qkv)) & 15) == 0) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(attn_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeqSmemLimit);
            attr = true;
        }
        attn_seq_bwd_kernel<<<(unsigned)(sh.n_seq * sh.H), kSeqThreads, seq_smem_bytes(sh, true), st>>>(
            qkv, o, lse, d_o, seq, sh, dr, d_qkv);
    } else {
        const int64_t warps = 2 * sh.n_seq * sh.H * sh.L;
        attn_bwd_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, st>>>(qkv, o, lse, d_o, seq, sh, dr, d_qkv);
    }
    note_launches(1);
    return check_launch("attention_bwd");
}
static int launch_colsum_acc(void* ws, const float* X, int64_t M, int64_t N, float* out, cudaStream_t st) {
    return colsum_dispatch(X, N, M, N, out, 1, ws, kGemmWsBytes, st);
}
static unsigned ew_blocks(int64_t n) {
    int64_t b = ceil_div(n, 256);
    return (unsigned)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

// saved-activation layout of one layer (floats)
struct LayerSaved {
    float *xin, *qkv, *lse, *o, *s1, *st1, *x1, *fd, *s2, *st2;
};
// every sub-array starts on a 16-byte boundary (float4 access) whatever the token count
static int64_t pad4(int64_t n) { return (n + 3) & ~(int64_t)3; }
static int64_t layer_floats(int64_t T, int d, int H) { return 9 * pad4(T * d) + pad4(T * H) + 2 * pad4(T * 2); }
static LayerSaved carve(float* base, int64_t T, int d, int H) {
    LayerSaved s;
    float* p = base;
    s.xin = p; p += pad4(T * d);
    s.qkv = p; p += 3 * pad4(T * d);
    s.lse = p; p += pad4(T * H);
    s.o = p; p += pad4(T * d);
    s.s1 = p; p += pad4(T * d);
    s.st1 = p; p += pad4(T * 2);
    s.x1 = p; p += pad4(T * d);
    s.fd = p; p += pad4(T * d);
    s.s2 = p; p += pad4(T * d);
    s.st2 = p; p += pad4(T * 2);
    return s;
}

}  // namespace c2dsr

using namespace c2dsr;

#define RUN(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_) return rc_;   \
    } while (0)

extern "C" {

int c2dsr_add_ln_fwd(const float* x, const float* y, const float* w, const float* b, float* s_out, float* out,
                     float* stats, int64_t n_tok, int d, int do_ln, float eps, float p, uint64_t seed, uint64_t tag,
                     void* stream) {
    if (n_tok <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d <= 32 * kMaxPerLane, "d must be in (0, 512]");
    return launch_add_ln(x, y, w, b, s_out, out, stats, n_tok, d, do_ln, eps, make_dropout(p, seed, tag),
                         (cudaStream_t)stream);
}

int c2dsr_ln_bwd(const float* d_out, const float* s, const float* stats, const float* w, float* dx_out,
                 int accumulate, int64_t n_tok, int d, void* stream) {
    if (n_tok <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d <= 32 * kMaxPerLane, "d must be in (0, 512]");
    return launch_ln_bwd(d_out, s, stats, w, dx_out, accumulate, n_tok, d, (cudaStream_t)stream);
}

int c2dsr_ln_param_grad(const float* d_out, const float* s, const float* stats, float* d_w, float* d_b,
                        int64_t n_tok, int d, void* stream) {
    if (n_tok <= 0) return C2DSR_OK;
    return launch_ln_param(nullptr, d_out, s, stats, d_w, d_b, n_tok, d, (cudaStream_t)stream);
}

int c2dsr_attention_fwd(const float* qkv, const int64_t* seq, int64_t n_seq, int L, int d, int n_head,
                        int64_t pad_idx, float p, uint64_t seed, uint64_t tag, float* o, float* lse, void* stream) {
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(n_head > 0 && d % n_head == 0 && d / n_head <= 32 * kMaxPerLane, "bad head configuration");
    return launch_attn_fwd(qkv, seq, make_shape(n_seq, L, d, n_head, pad_idx), make_dropout(p, seed, tag), o, lse,
                           (cudaStream_t)stream);
}

int c2dsr_attention_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, const int64_t* seq,
                        int64_t n_seq, int L, int d, int n_head, int64_t pad_idx, float p, uint64_t seed,
                        uint64_t tag, float* d_qkv, void* stream) {
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(n_head > 0 && d % n_head == 0 && d / n_head <= 32 * kMaxPerLane, "bad head configuration");
    return launch_attn_bwd(qkv, o, lse, d_o, seq, make_shape(n_seq, L, d, n_head, pad_idx),
                           make_dropout(p, seed, tag), d_qkv, (cudaStream_t)stream);
}

int64_t c2dsr_encoder_saved_floats(int64_t n_tok, int d, int n_head, int n_layers) {
    return n_layers * layer_floats(n_tok, d, n_head) + n_tok * ((int64_t)d + 2);
}

int64_t c2dsr_encoder_workspace_bytes(int64_t n_tok, int d, int n_head, int dense_passes) {
    (void)n_head;
    return 7 * n_tok * (int64_t)d * 4 + kGemmWsBytes + (dense_passes ? dense_tc_bytes(n_tok, d) : 0) + 1024;
}

int c2dsr_encoder_fwd(const c2dsr_layer_weights* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                      const float* x, const int64_t* seq, int64_t n_seq, int L, int d, int n_head, int64_t pad_idx,
                      int norm_first, int dense_passes, float eps, float p, uint64_t seed, uint64_t tag, float* out,
                      float* saved, void* workspace, int64_t workspace_bytes, void* stream) {
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 32 * kMaxPerLane, "d must be a multiple of 4 in (0, 512]");
    C2DSR_REQUIRE(n_head > 0 && d % n_head == 0, "d must be divisible by n_head");
    const int64_t T = n_seq * L;
    if (workspace_bytes < c2dsr_encoder_workspace_bytes(T, d, n_head, dense_passes)) {
        set_error("encoder_fwd: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* ybuf = (float*)workspace;
    void* gws = (char*)workspace + 7 * T * (int64_t)d * 4;
    void* tws = (char*)gws + kGemmWsBytes;
    const int64_t tws_bytes = dense_passes ? dense_tc_bytes(T, d) : 0;
    C2DSR_REQUIRE(dense_passes == 0 || dense_passes == 1 || dense_passes == 3, "dense_passes must be 0, 1 or 3");
    const Dropout none = make_dropout(0.f, 0, 0);
    const AttnShape sh = make_shape(n_seq, L, d, n_head, pad_idx);
    const int64_t lf = layer_floats(T, d, n_head);
    float* xlast = saved + n_layers * lf;
    float* stf = xlast + T * d;
    {
        float* first = n_layers > 0 ? carve(saved, T, d, n_head).xin : xlast;
        cudaMemcpyAsync(first, x, T * (int64_t)d * 4, cudaMemcpyDeviceToDevice, st);
    }
    for (int l = 0; l < n_layers; ++l) {
        const LayerSaved s = carve(saved + l * lf, T, d, n_head);
        const c2dsr_layer_weights& w = layers[l];
        float* next = l + 1 < n_layers ? carve(saved + (l + 1) * lf, T, d, n_head).xin : xlast;
        // per-layer, per-site tags; the "seed is a device pointer" flag (top bit) is carried over unchanged
        const uint64_t tb = ((tag & ~kSeedIndirect) * 1024 + (uint64_t)l * 8) | (tag & kSeedIndirect);
        const float* attn_in = s.xin;
        if (norm_first) {
            RUN(launch_add_ln(s.xin, nullptr, w.ln1_w, w.ln1_b, nullptr, s.s1, s.st1, T, d, 1, eps, none, st));
            attn_in = s.s1;
        }
        RUN(dense(dense_passes, tws, tws_bytes, 0, 1, T, 3 * d, d, 1.f, attn_in, d, w.in_proj_w, d, 0.f, s.qkv, 3 * d, w.in_proj_b, 0, none,
                          gws, kGemmWsBytes, st));
        RUN(launch_attn_fwd(s.qkv, seq, sh, make_dropout(p, seed, tb + 0), s.o, s.lse, st));
        RUN(dense(dense_passes, tws, tws_bytes, 0, 1, T, d, d, 1.f, s.o, d, w.out_proj_w, d, 0.f, ybuf, d, w.out_proj_b, 0, none, gws,
                          kGemmWsBytes, st));
        const float* ffn_in;
        if (norm_first) {
            RUN(launch_add_ln(s.xin, ybuf, nullptr, nullptr, nullptr, s.x1, nullptr, T, d, 0, eps,
                              make_dropout(p, seed, tb + 1), st));
            RUN(launch_add_ln(s.x1, nullptr, w.ln2_w, w.ln2_b, nullptr, s.s2, s.st2, T, d, 1, eps, none, st));
            ffn_in = s.s2;
        } else {
            RUN(launch_add_ln(s.xin, ybuf, w.ln1_w, w.ln1_b, s.s1, s.x1, s.st1, T, d, 1, eps,
                              make_dropout(p, seed, tb + 1), st));
            ffn_in = s.x1;
        }
        RUN(dense(dense_passes, tws, tws_bytes, 0, 1, T, d, d, 1.f, ffn_in, d, w.lin1_w, d, 0.f, s.fd, d, w.lin1_b, 1,
                          make_dropout(p, seed, tb + 2), gws, kGemmWsBytes, st));
        RUN(dense(dense_passes, tws, tws_bytes, 0, 1, T, d, d, 1.f, s.fd, d, w.lin2_w, d, 0.f, ybuf, d, w.lin2_b, 0, none, gws,
                          kGemmWsBytes, st));
        if (norm_first) {
            RUN(launch_add_ln(s.x1, ybuf, nullptr, nullptr, nullptr, next, nullptr, T, d, 0, eps,
                              make_dropout(p, seed, tb + 3), st));
        } else {
            RUN(launch_add_ln(s.x1, ybuf, w.ln2_w, w.ln2_b, s.s2, next, s.st2, T, d, 1, eps,
                              make_dropout(p, seed, tb + 3), st));
        }
    }
    RUN(launch_add_ln(xlast, nullptr, lnf_w, lnf_b, nullptr, out, stf, T, d, 1, eps, none, st));
    return C2DSR_OK;
}

int64_t c2dsr_encoder_select_workspace_bytes(int64_t n_seq, int L, int d, int dense_passes) {
    const int64_t T = n_seq * L;
    return (4 * T + 6 * n_seq) * (int64_t)d * 4 + kGemmWsBytes + (dense_passes ? dense_tc_bytes(T, d) : 0) + 1024;
}

int c2dsr_encoder_fwd_select(const c2dsr_layer_weights* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                             const float* x, const int64_t* seq, const int64_t* sel, int64_t n_seq, int L, int d,
                             int n_head, int64_t pad_idx, int norm_first, int dense_passes, float eps, float* out,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(n_layers == 1, "the single-query forward covers one encoder layer (use c2dsr_encoder_fwd otherwise)");
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 32 * kMaxPerLane, "d must be a multiple of 4 in (0, 512]");
    C2DSR_REQUIRE(n_head > 0 && d % n_head == 0, "d must be divisible by n_head");
    C2DSR_REQUIRE(dense_passes == 0 || dense_passes == 1 || dense_passes == 3, "dense_passes must be 0, 1 or 3");
    if (workspace_bytes < c2dsr_encoder_select_workspace_bytes(n_seq, L, d, dense_passes)) {
        set_error("encoder_fwd_select: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t T = n_seq * L, Td = T * d, Bd = n_seq * d;
    float* s1 = (float*)workspace;              // [T, d]   pre-norm only: LayerNorm1(x)
    float* qkv = s1 + Td;                       // [T, 3d]
    float* x_sel = qkv + 3 * Td;                // [n_seq, d] each from here on
    float* o_sel = x_sel + Bd;
    float* y = o_sel + Bd;
    float* x1 = y + Bd;
    float* t1 = x1 + Bd;
    float* t2 = t1 + Bd;
    void* gws = t2 + Bd;
    void* tws = (char*)gws + kGemmWsBytes;
    const int64_t tws_bytes = dense_passes ? dense_tc_bytes(T, d) : 0;
    const Dropout none = make_dropout(0.f, 0, 0);
    const AttnShape sh = make_shape(n_seq, L, d, n_head, pad_idx);
    const c2dsr_layer_weights& w = layers[0];
    const float* attn_in = x;
    if (norm_first) {
        RUN(launch_add_ln(x, nullptr, w.ln1_w, w.ln1_b, nullptr, s1, nullptr, T, d, 1, eps, none, st));
        attn_in = s1;
    }
    // keys and values of every token (and, for simplicity, every query: one GEMM)
    RUN(dense(dense_passes, tws, tws_bytes, 0, 1, T, 3 * d, d, 1.f, attn_in, d, w.in_proj_w, d, 0.f, qkv, 3 * d,
              w.in_proj_b, 0, none, gws, kGemmWsBytes, st));
    attn_select_kernel<<<(unsigned)ceil_div(n_seq * n_head, 8), 256, 0, st>>>(qkv, seq, sel, sh, o_sel);
    select_rows_kernel<<<(unsigned)ceil_div(n_seq, 8), 256, 0, st>>>(x, sel, n_seq, L, d, x_sel);
    note_launches(2);
    // from here on: n_seq rows
    RUN(dense(dense_passes, tws, tws_bytes, 0, 1, n_seq, d, d, 1.f, o_sel, d, w.out_proj_w, d, 0.f, y, d, w.out_proj_b, 0,
              none, gws, kGemmWsBytes, st));
    const float* ffn_in;
    if (norm_first) {
        RUN(launch_add_ln(x_sel, y, nullptr, nullptr, nullptr, x1, nullptr, n_seq, d, 0, eps, none, st));
        RUN(launch_add_ln(x1, nullptr, w.ln2_w, w.ln2_b, nullptr, t1, nullptr, n_seq, d, 1, eps, none, st));
        ffn_in = t1;
    } else {
        RUN(launch_add_ln(x_sel, y, w.ln1_w, w.ln1_b, nullptr, x1, nullptr, n_seq, d, 1, eps, none, st));
        ffn_in = x1;
    }
    RUN(dense(dense_passes, tws, tws_bytes, 0, 1, n_seq, d, d, 1.f, ffn_in, d, w.lin1_w, d, 0.f, t2, d, w.lin1_b, 1, none, gws,
              kGemmWsBytes, st));
    RUN(dense(dense_passes, tws, tws_bytes, 0, 1, n_seq, d, d, 1.f, t2, d, w.lin2_w, d, 0.f, y, d, w.lin2_b, 0, none, gws,
              kGemmWsBytes, st));
    float* last = t1;
    if (norm_first) {
        RUN(launch_add_ln(x1, y, nullptr, nullptr, nullptr, last, nullptr, n_seq, d, 0, eps, none, st));
    } else {
        RUN(launch_add_ln(x1, y, w.ln2_w, w.ln2_b, nullptr, last, nullptr, n_seq, d, 1, eps, none, st));
    }
    RUN(launch_add_ln(last, nullptr, lnf_w, lnf_b, nullptr, out, nullptr, n_seq, d, 1, eps, none, st));
    return check_launch("encoder_fwd_select");
}

int64_t c2dsr_encoder_padkeys_workspace_bytes(int64_t n_seq, int d, int dense_passes) {
    return (6 * n_seq + 8) * (int64_t)d * 4 + kGemmWsBytes + (dense_passes ? dense_tc_bytes(n_seq, d) : 0) + 1024;
}

// the PAD token's attention-block output y_pad [d] of a one-layer encoder: value projection of x_pad, then the output
// projection (exact fp32).  It depends on the weights and on the propagated PAD row only, not on the batch.
int c2dsr_encoder_padkeys_prepare(const c2dsr_layer_weights* layers, int n_layers, const float* x_pad, int d,
                                  int norm_first, float eps, float* y_pad, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
    C2DSR_REQUIRE(n_layers == 1, "the pad-key forward covers one encoder layer");
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 32 * kMaxPerLane, "d must be a multiple of 4 in (0, 512]");
    if (workspace_bytes < (int64_t)(8 * d) * 4 + kGemmWsBytes) {
        set_error("encoder_padkeys_prepare: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* pad_in = (float*)workspace;
    float* pad_qkv = pad_in + d;
    void* gws = pad_qkv + 4 * d;
    const Dropout none = make_dropout(0.f, 0, 0);
    const c2dsr_layer_weights& w = layers[0];
    const float* attn_in = x_pad;
    if (norm_first) {
        RUN(launch_add_ln(x_pad, nullptr, w.ln1_w, w.ln1_b, nullptr, pad_in, nullptr, 1, d, 1, eps, none, st));
        attn_in = pad_in;
    }
    RUN(dense(0, nullptr, 0, 0, 1, 1, 3 * d, d, 1.f, attn_in, d, w.in_proj_w, d, 0.f, pad_qkv, 3 * d, w.in_proj_b, 0, none,
              gws, kGemmWsBytes, st));
    RUN(dense(0, nullptr, 0, 0, 1, 1, d, d, 1.f, pad_qkv + 2 * d, d, w.out_proj_w, d, 0.f, y_pad, d, w.out_proj_b, 0, none,
              gws, kGemmWsBytes, st));
    return check_launch("encoder_padkeys_prepare");
}

int c2dsr_encoder_fwd_padkeys(const c2dsr_layer_weights* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                              const float* x_sel, const float* y_pad, const int64_t* seq, const int64_t* sel,
                              int64_t n_seq, int L, int d, int n_head, int64_t pad_idx, int norm_first,
                              int dense_passes, float eps, float* out, void* workspace, int64_t workspace_bytes,
                              void* stream) {
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(n_layers == 1, "the pad-key forward covers one encoder layer (use c2dsr_encoder_fwd otherwise)");
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 32 * kMaxPerLane, "d must be a multiple of 4 in (0, 512]");
    C2DSR_REQUIRE(n_head > 0 && d % n_head == 0, "d must be divisible by n_head");
    C2DSR_REQUIRE(dense_passes == 0 || dense_passes == 1 || dense_passes == 3, "dense_passes must be 0, 1 or 3");
    if (workspace_bytes < c2dsr_encoder_padkeys_workspace_bytes(n_seq, d, dense_passes)) {
        set_error("encoder_fwd_padkeys: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Bd = n_seq * d;
    float* y = (float*)workspace;               // [n_seq, d] each
    float* x1 = y + Bd;
    float* t1 = x1 + Bd;
    float* t2 = t1 + Bd;
    void* gws = t2 + 3 * Bd + 8 * d;
    void* tws = (char*)gws + kGemmWsBytes;
    const int64_t tws_bytes = dense_passes ? dense_tc_bytes(n_seq, d) : 0;
    const Dropout none = make_dropout(0.f, 0, 0);
    const c2dsr_layer_weights& w = layers[0];
    padkey_rows_kernel<<<(unsigned)ceil_div(n_seq, 8), 256, 0, st>>>(seq, sel, n_seq, L, d, pad_idx, y_pad, w.out_proj_b,
                                                                    y);
    note_launches(1);
    // from here on exactly c2dsr_encoder_fwd_select: n_seq rows
    const float* ffn_in;
    if (norm_first) {
        RUN(launch_add_ln(x_sel, y, nullptr, nullptr, nullptr, x1, nullptr, n_seq, d, 0, eps, none, st));
        RUN(launch_add_ln(x1, nullptr, w.ln2_w, w.ln2_b, nullptr, t1, nullptr, n_seq, d, 1, eps, none, st));
        ffn_in = t1;
    } else {
        RUN(launch_add_ln(x_sel, y, w.ln1_w, w.ln1_b, nullptr, x1, nullptr, n_seq, d, 1, eps, none, st));
        ffn_in = x1;
    }
    RUN(dense(dense_passes, tws, tws_bytes, 0, 1, n_seq, d, d, 1.f, ffn_in, d, w.lin1_w, d, 0.f, t2, d, w.lin1_b, 1, none, gws,
              kGemmWsBytes, st));
    RUN(dense(dense_passes, tws, tws_bytes, 0, 1, n_seq, d, d, 1.f, t2, d, w.lin2_w, d, 0.f, y, d, w.lin2_b, 0, none, gws,
              kGemmWsBytes, st));
    float* last = t1;
    if (norm_first) {
        RUN(launch_add_ln(x1, y, nullptr, nullptr, nullptr, last, nullptr, n_seq, d, 0, eps, none, st));
    } else {
        RUN(launch_add_ln(x1, y, w.ln2_w, w.ln2_b, nullptr, last, nullptr, n_seq, d, 1, eps, none, st));
    }
    RUN(launch_add_ln(last, nullptr, lnf_w, lnf_b, nullptr, out, nullptr, n_seq, d, 1, eps, none, st));
    return check_launch("encoder_fwd_padkeys");
}

int c2dsr_encoder_bwd(const c2dsr_layer_weights* layers, const c2dsr_layer_grads* grads, int n_layers,
                      const float* lnf_w, float* d_lnf_w, float* d_lnf_b, const float* d_out, const int64_t* seq,
                      int64_t n_seq, int L, int d, int n_head, int64_t pad_idx, int norm_first, int dense_passes,
                      float eps, float p, uint64_t seed, uint64_t tag, const float* saved_c, float* dx,
                      void* workspace, int64_t workspace_bytes, void* stream) {
    (void)eps;
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d % 4 == 0 && d <= 32 * kMaxPerLane, "d must be a multiple of 4 in (0, 512]");
    const int64_t T = n_seq * L;
    if (workspace_bytes < c2dsr_encoder_workspace_bytes(T, d, n_head, dense_passes)) {
        set_error("encoder_bwd: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* saved = const_cast<float*>(saved_c);
    const int64_t Td = T * (int64_t)d;
    float* g = (float*)workspace;        // gradient w.r.t. the current layer output / input
    float* ds = g + Td;
    float* dy = ds + Td;
    float* dfd = dy + Td;
    float* dqkv = dfd + Td;              // [T, 3d]
    void* gws = (char*)workspace + 7 * Td * 4;
    void* tws = (char*)gws + kGemmWsBytes;
    const int64_t tws_bytes = dense_passes ? dense_tc_bytes(T, d) : 0;
    const Dropout none = make_dropout(0.f, 0, 0);
    const AttnShape sh = make_shape(n_seq, L, d, n_head, pad_idx);
    const int64_t lf = layer_floats(T, d, n_head);
    float* xlast = saved + n_layers * lf;
    float* stf = xlast + Td;
    const bool has_drop = p > 0.f && p < 1.f;
    const float inv_keep = has_drop ? 1.f / (1.f - p) : 1.f;

    RUN(launch_ln_param(gws, d_out, xlast, stf, d_lnf_w, d_lnf_b, T, d, st));
    RUN(launch_ln_bwd(d_out, xlast, stf, lnf_w, g, 0, T, d, st));

    for (int l = n_layers - 1; l >= 0; --l) {
        const LayerSaved s = carve(saved + l * lf, T, d, n_head);
        const c2dsr_layer_weights& w = layers[l];
        const c2dsr_layer_grads& gw = grads[l];
        // per-layer, per-site tags; the "seed is a device pointer" flag (top bit) is carried over unchanged
        const uint64_t tb = ((tag & ~kSeedIndirect) * 1024 + (uint64_t)l * 8) | (tag & kSeedIndirect);
        auto masked = [&](const float* src, uint64_t site) -> const float* {
            if (!has_drop) return src;
            drop_mul_kernel<<<ew_blocks(Td), 256, 0, st>>>(src, dy, Td, make_dropout(p, seed, tb + site));
            note_launches(1);
            return dy;
        };
        // ---- feed-forward block ----
        const float* d_s2;          // gradient at the output of the FFN residual sum
        if (norm_first) {
            d_s2 = g;               // x2 = x1 + drop(y2): the sum is the layer output
        } else {
            RUN(launch_ln_param(gws, g, s.s2, s.st2, gw.ln2_w, gw.ln2_b, T, d, st));
            RUN(launch_ln_bwd(g, s.s2, s.st2, w.ln2_w, ds, 0, T, d, st));
            d_s2 = ds;
        }
        const float* d_y2 = masked(d_s2, 3);
        const float* ffn_in = norm_first ? s.s2 : s.x1;
        RUN(dense(dense_passes, tws, tws_bytes, 1, 0, d, d, T, 1.f, d_y2, d, s.fd, d, 1.f, gw.lin2_w, d, nullptr, 0, none, gws, kGemmWsBytes, st));
        RUN(launch_colsum_acc(gws, d_y2, T, d, gw.lin2_b, st));
        RUN(dense(dense_passes, tws, tws_bytes, 0, 0, T, d, d, 1.f, d_y2, d, w.lin2_w, d, 0.f, dfd, d, nullptr, 0, none, gws, kGemmWsBytes, st));
        relu_drop_bwd_kernel<<<ew_blocks(Td), 256, 0, st>>>(dfd, s.fd, Td, inv_keep);
        note_launches(1);
        RUN(dense(dense_passes, tws, tws_bytes, 1, 0, d, d, T, 1.f, dfd, d, ffn_in, d, 1.f, gw.lin1_w, d, nullptr, 0, none, gws, kGemmWsBytes, st));
        RUN(launch_colsum_acc(gws, dfd, T, d, gw.lin1_b, st));
        // gradient w.r.t. x1 (post-norm: d_s2 + d_pre W1; pre-norm: g + LN2^T(d_pre W1))
        float* d_x1;
        if (norm_first) {
            RUN(dense(dense_passes, tws, tws_bytes, 0, 0, T, d, d, 1.f, dfd, d, w.lin1_w, d, 0.f, ds, d, nullptr, 0, none, gws, kGemmWsBytes, st));
            RUN(launch_ln_param(gws, ds, s.x1, s.st2, gw.ln2_w, gw.ln2_b, T, d, st));
            RUN(launch_ln_bwd(ds, s.x1, s.st2, w.ln2_w, g, 1, T, d, st));
            d_x1 = g;
        } else {
            RUN(dense(dense_passes, tws, tws_bytes, 0, 0, T, d, d, 1.f, dfd, d, w.lin1_w, d, 1.f, ds, d, nullptr, 0, none, gws, kGemmWsBytes, st));
            d_x1 = ds;
        }
        // ---- attention block ----
        const float* d_s1;
        if (norm_first) {
            d_s1 = d_x1;            // x1 = xin + drop(y)
        } else {
            RUN(launch_ln_param(gws, d_x1, s.s1, s.st1, gw.ln1_w, gw.ln1_b, T, d, st));
            RUN(launch_ln_bwd(d_x1, s.s1, s.st1, w.ln1_w, g, 0, T, d, st));
            d_s1 = g;
        }
        const float* d_y = masked(d_s1, 1);
        const float* attn_in = norm_first ? s.s1 : s.xin;
        RUN(dense(dense_passes, tws, tws_bytes, 1, 0, d, d, T, 1.f, d_y, d, s.o, d, 1.f, gw.out_proj_w, d, nullptr, 0, none, gws, kGemmWsBytes, st));
        RUN(launch_colsum_acc(gws, d_y, T, d, gw.out_proj_b, st));
        RUN(dense(dense_passes, tws, tws_bytes, 0, 0, T, d, d, 1.f, d_y, d, w.out_proj_w, d, 0.f, dfd, d, nullptr, 0, none, gws, kGemmWsBytes, st));
        RUN(launch_attn_bwd(s.qkv, s.o, s.lse, dfd, seq, sh, make_dropout(p, seed, tb + 0), dqkv, st));
        RUN(dense(dense_passes, tws, tws_bytes, 1, 0, 3 * d, d, T, 1.f, dqkv, 3 * d, attn_in, d, 1.f, gw.in_proj_w, d, nullptr, 0, none, gws, kGemmWsBytes, st));
        RUN(launch_colsum_acc(gws, dqkv, T, 3 * d, gw.in_proj_b, st));
        if (norm_first) {
            RUN(dense(dense_passes, tws, tws_bytes, 0, 0, T, d, 3 * d, 1.f, dqkv, 3 * d, w.in_proj_w, d, 0.f, ds, d, nullptr, 0, none, gws, kGemmWsBytes, st));
            RUN(launch_ln_param(gws, ds, s.xin, s.st1, gw.ln1_w, gw.ln1_b, T, d, st));
            RUN(launch_ln_bwd(ds, s.xin, s.st1, w.ln1_w, g, 1, T, d, st));
        } else {
            RUN(dense(dense_passes, tws, tws_bytes, 0, 0, T, d, 3 * d, 1.f, dqkv, 3 * d, w.in_proj_w, d, 1.f, g, d, nullptr, 0, none, gws, kGemmWsBytes, st));
        }
        // g now holds the gradient w.r.t. this layer's input
    }
    cudaMemcpyAsync(dx, g, Td * 4, cudaMemcpyDeviceToDevice, st);
    return check_launch("encoder_bwd");
}

}  // extern "C"
