// K1: branch-input gather  x = drop(scale * (hi[seq] + E[seq]) + P[pos])  and its deterministic backward.
//
// Forward is one warp per token, float4 lanes across the feature dimension: each token reads two table
// rows and one positional row (3 * d * 4 B) and writes one row (d * 4 B), fully coalesced.
//
// Backward needs a scatter-add of token rows into table rows (the pad row collects roughly half of
// all tokens).  To keep it deterministic there are no float atomics: tokens are ranked by (key, index)
// -- a stable sort by counting -- and summed segment by segment in sorted order.  Sorted positions are
// cut into fixed chunks of 32; a warp sums the runs inside its chunk, runs that cross a chunk boundary
// go through a per-chunk head/tail partial and are stitched together, in chunk order, by one warp.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

constexpr int kChunk = 32;

__global__ void gather_fwd_kernel(const float* __restrict__ hi, const float* __restrict__ E,
                                  const float* __restrict__ P, const int64_t* __restrict__ seq,
                                  const int64_t* __restrict__ pos, float* __restrict__ x, int64_t n_tok, int d,
                                  float scale, Dropout dr) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tok) return;
    const int64_t item = seq[t], ps = pos[t];
    const float4* h4 = reinterpret_cast<const float4*>(hi + item * d);
    const float4* e4 = reinterpret_cast<const float4*>(E + item * d);
    const float4* p4 = reinterpret_cast<const float4*>(P + ps * d);
    float4* o4 = reinterpret_cast<float4*>(x + t * d);
    const int nv = d >> 2;
    for (int v = lane; v < nv; v += 32) {
        float4 a = __ldg(h4 + v), b = __ldg(e4 + v), c = __ldg(p4 + v), r;
        // separately rounded add, multiply, add (no FMA contraction): bit-identical to the reference's
        // (hi[seq] + E[seq]) * sqrt(d) followed by += pos_emb(pos)
        r.x = __fadd_rn(__fmul_rn(__fadd_rn(a.x, b.x), scale), c.x);
        r.y = __fadd_rn(__fmul_rn(__fadd_rn(a.y, b.y), scale), c.y);
        r.z = __fadd_rn(__fmul_rn(__fadd_rn(a.z, b.z), scale), c.z);
        r.w = __fadd_rn(__fmul_rn(__fadd_rn(a.w, b.w), scale), c.w);
        if (dr.p != 0.f) {
            const uint64_t base = (uint64_t)t * d + 4 * v;
            r.x *= drop_scale(dr, base);
            r.y *= drop_scale(dr, base + 1);
            r.z *= drop_scale(dr, base + 2);
            r.w *= drop_scale(dr, base + 3);
        }
        o4[v] = r;
    }
}

// rank[t] = #{u : (key[u], u) < (key[t], t)}  -- stable sort position by counting.
__global__ void rank_kernel(const int64_t* __restrict__ key, int64_t n, int32_t* __restrict__ perm) {
    __shared__ int64_t tile[1024];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t mine = t < n ? key[t] : 0;
    int cnt = 0;
    for (int64_t base = 0; base < n; base += 1024) {
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) tile[i] = base + i < n ? key[base + i] : INT64_MAX;
        __syncthreads();
        const int lim = (int)(n - base < 1024 ? n - base : 1024);
        if (t < n) {
#pragma unroll 8
            for (int i = 0; i < lim; ++i) {
                const int64_t k = tile[i];
                cnt += (k < mine) || (k == mine && base + i < t);
            }
        }
        __syncthreads();
    }
    if (t < n) perm[cnt] = (int32_t)t;     // perm[sorted position] = token
}

// Single-CTA bitonic sort of (key << 32 | token) for n <= 16384: perm[sorted position] = token.
__global__ void bitonic_perm_kernel(const int64_t* __restrict__ key, int n, int n2, int32_t* __restrict__ perm) {
    extern __shared__ unsigned long long skeys[];
    for (int i = threadIdx.x; i < n2; i += blockDim.x)
        skeys[i] = i < n ? (((unsigned long long)key[i] << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = skeys[i], y = skeys[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        skeys[i] = y;
                        skeys[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = (int32_t)(skeys[i] & 0xffffffffull);
}

struct SegArgs {
    const float* rows;      // [n, d] token rows
    const int64_t* key;     // [n]
    const int32_t* perm;    // [n] sorted position -> token
    float* out1;            // [*, d] accumulated
    float* out2;            // optional second destination (same values)
    int64_t skip2;          // key that out2 ignores (padding_idx), or -1
    float* head;            // [n_chunks, d] partial of the run touching the chunk start
    float* tail;            // [n_chunks, d] partial of the run touching the chunk end
    int64_t n;
    int d;
    float scale;
    Dropout dr;
};

__device__ __forceinline__ void seg_flush(const SegArgs& a, int64_t key, const float (&acc)[16], int nper, int lane,
                                          float* dst_partial) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int f = lane + 32 * k;
        if (k < nper && f < a.d) {
            if (dst_partial) {
                dst_partial[f] = acc[k];
            } else {
                a.out1[key * a.d + f] += acc[k];
                if (a.out2 && key != a.skip2) a.out2[key * a.d + f] += acc[k];
            }
        }
    }
}

// One warp per chunk of 32 sorted positions.  Lane l owns features l, l+32, ... (d <= 32*16).
__global__ void seg_chunk_kernel(SegArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t lo = c * kChunk;
    if (lo >= a.n) return;                                  // warp-uniform
    const int cnt = (int)(lo + kChunk < a.n ? kChunk : a.n - lo);
    const int nper = (a.d + 31) >> 5;
    // lane l holds sorted position lo + l: its token and key, broadcast by shuffle below
    const int32_t my_tok = lane < cnt ? a.perm[lo + lane] : 0;
    const int64_t my_key = lane < cnt ? a.key[my_tok] : -1;
    const int64_t prev_key = lo > 0 ? a.key[a.perm[lo - 1]] : -2;
    const int64_t next_key = lo + cnt < a.n ? a.key[a.perm[lo + cnt]] : -2;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    int64_t run_key = __shfl_sync(0xffffffffu, my_key, 0);
    bool run_from_prev = prev_key == run_key;
    for (int r = 0; r < cnt; ++r) {
        const int64_t tok = __shfl_sync(0xffffffffu, my_tok, r);
        const int64_t k = __shfl_sync(0xffffffffu, my_key, r);
        if (k != run_key) {                                 // warp-uniform branch
            seg_flush(a, run_key, acc, nper, lane, run_from_prev ? a.head + c * a.d : nullptr);
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = 0.f;
            run_key = k;
            run_from_prev = false;
        }
        const float* row = a.rows + tok * a.d;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int f = lane + 32 * q;
            if (q < nper && f < a.d) acc[q] += a.scale * __ldg(row + f) * drop_scale(a.dr, (uint64_t)tok * a.d + f);
        }
    }
    float* dst = nullptr;
    if (run_from_prev) dst = a.head + c * a.d;              // also the whole-chunk case
    else if (next_key == run_key) dst = a.tail + c * a.d;
    seg_flush(a, run_key, acc, nper, lane, dst);
}

// One warp per chunk whose last run starts a chunk-crossing segment: add tail[c] + head[c+1] + ... in order.
__global__ void seg_stitch_kernel(SegArgs a, int64_t n_chunks) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_chunks) return;
    const int64_t last = (c + 1) * kChunk - 1;
    if (last + 1 >= a.n) return;                                   // nothing after this chunk
    const int64_t key = a.key[a.perm[last]];
    if (a.key[a.perm[last + 1]] != key) return;                    // the last run ends here
    const int64_t lo = c * kChunk;
    const bool whole = a.key[a.perm[lo]] == key;                   // chunk is a single run
    if (whole && lo > 0 && a.key[a.perm[lo - 1]] == key) return;   // segment started earlier: not the owner
    const int nper = (a.d + 31) >> 5;
    float acc[16];
    // a whole-chunk run that starts its segment is stored in tail (run_from_prev false, run_to_next true)
    const float* first = a.tail + c * a.d;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int f = lane + 32 * k;
        acc[k] = (k < nper && f < a.d) ? first[f] : 0.f;
    }
    for (int64_t cc = c + 1; cc < n_chunks; ++cc) {
        const float* h = a.head + cc * a.d;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int f = lane + 32 * k;
            if (k < nper && f < a.d) acc[k] += h[f];
        }
        const int64_t cl = (cc + 1) * kChunk - 1;                  // does the segment continue past chunk cc?
        if (cl + 1 >= a.n) break;
        if (a.key[a.perm[cl]] != key || a.key[a.perm[cl + 1]] != key) break;
    }
    seg_flush(a, key, acc, nper, lane, nullptr);
}

static int segmented_rowsum(const float* rows, const int64_t* key, float* out1, float* out2, int64_t skip2,
                            int64_t n, int d, float scale, Dropout dr, char* ws, cudaStream_t st) {
    int32_t* perm = reinterpret_cast<int32_t*>(ws);
    const int64_t n_chunks = ceil_div(n, kChunk);
    float* head = reinterpret_cast<float*>(ws + align_up(n * 4, 256));
    float* tail = head + n_chunks * d;
    if (n <= 16384) {
        int n2 = 32;
        while (n2 < n) n2 <<= 1;
        const int smem = n2 * 8;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(bitonic_perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
            attr_set = true;
        }
        bitonic_perm_kernel<<<1, 1024, smem, st>>>(key, (int)n, n2, perm);
    } else {
        rank_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, st>>>(key, n, perm);
    }
    SegArgs a{rows, key, perm, out1, out2, skip2, head, tail, n, d, scale, dr};
    seg_chunk_kernel<<<(unsigned)ceil_div(n_chunks, 4), 128, 0, st>>>(a);
    seg_stitch_kernel<<<(unsigned)ceil_div(n_chunks, 4), 128, 0, st>>>(a, n_chunks);
    note_launches(3);
    return check_launch("segmented_rowsum");
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int c2dsr_gather_fwd(const float* hi, const float* E, const float* P, const int64_t* seq, const int64_t* pos,
                     float* x, int64_t n_tok, int d, float scale, float p, uint64_t seed, uint64_t tag,
                     void* stream) {
    if (n_tok <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d % 4 == 0, "d must be a positive multiple of 4");
    gather_fwd_kernel<<<(unsigned)ceil_div(n_tok, 8), 256, 0, (cudaStream_t)stream>>>(
        hi, E, P, seq, pos, x, n_tok, d, scale, make_dropout(p, seed, tag));
    note_launches(1);
    return check_launch("gather_fwd");
}

int64_t c2dsr_gather_bwd_workspace_bytes(int64_t n_tok, int d) {
    if (n_tok <= 0) return 256;
    return align_up(n_tok * 4, 256) + 2 * ceil_div(n_tok, kChunk) * (int64_t)d * 4 + 256;
}

int c2dsr_gather_bwd(const float* dx, const int64_t* seq, const int64_t* pos, float* d_hi, float* d_E, float* d_P,
                     int64_t n_tok, int d, int64_t pad_idx, float scale, float p, uint64_t seed, uint64_t tag,
                     void* workspace, int64_t workspace_bytes, void* stream) {
    if (n_tok <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d <= 512, "d must be in (0, 512]");
    if (workspace_bytes < c2dsr_gather_bwd_workspace_bytes(n_tok, d)) {
        set_error("gather_bwd: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    Dropout dr = make_dropout(p, seed, tag);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = segmented_rowsum(dx, seq, d_hi, d_E, pad_idx, n_tok, d, scale, dr, (char*)workspace, st);
    if (rc) return rc;
    if (d_P) rc = segmented_rowsum(dx, pos, d_P, nullptr, -1, n_tok, d, 1.f, dr, (char*)workspace, st);
    return rc;
}

}  // extern "C"
