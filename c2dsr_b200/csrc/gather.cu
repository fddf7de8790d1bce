// K1: branch-input gather  x = drop(scale * (hi[seq] + E[seq]) + P[pos])  and its deterministic backward.
//
// Forward: one warp per token, float4 lanes across the feature dimension: each token reads two table
// rows and one positional row (3 * d * 4 B) and writes one row (d * 4 B), fully coalesced.
//
// Backward is a scatter-add of token rows g[t] into table rows, made deterministic without sorting and
// without float atomics:
//   * item rows: integer atomics elect, per item, the first token that references it and count the
//     references.  An item referenced once (the common case) is written by that token's warp.  For an item
//     referenced several times the tokens are bucketed: the elected token reserves a run of a token list
//     (one atomicAdd of the item's count), every token appends itself to its item's run (slot from an integer
//     atomic: any order), and the elected token's warp sorts the run (rank by counting, in registers) and adds the
//     rows in ascending token order.  Runs longer than a warp (an item more than 32 times in one batch) fall back
//     to a scan of the token list, also in token order.
//   * the pad row (about half of all tokens) and the positional rows (at most len_max of them, every token
//     contributes to one) are dense reductions: tokens are cut into fixed chunks, each chunk accumulates
//     its (len_max + 1) bins in shared memory in token order, and a second kernel adds the chunk partials
//     in chunk order.
// Every float sum therefore has a fixed order; the integer atomics are order independent.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

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
        if (dr.p != 0.f) drop_scale4(dr, (uint64_t)t * d + 4 * v, r);
        o4[v] = r;
    }
}

// evaluation: only position sel[b] of sequence b is needed (plus, as row n_seq, one PAD token at position 0): the
// same arithmetic as gather_fwd_kernel for those tokens, without the index-gather launches in front of it
__global__ void gather_select_fwd_kernel(const float* __restrict__ hi, const float* __restrict__ E,
                                         const float* __restrict__ P, const int64_t* __restrict__ seq,
                                         const int64_t* __restrict__ pos, const int64_t* __restrict__ sel,
                                         float* __restrict__ x, int64_t n_seq, int L, int d, float scale, int64_t pad) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b > n_seq) return;
    int64_t item = pad, ps = 0;
    if (b < n_seq) {
        const int64_t t = b * L + sel[b];
        item = seq[t];
        ps = pos[t];
    }
    const float4* h4 = reinterpret_cast<const float4*>(hi + item * d);
    const float4* e4 = reinterpret_cast<const float4*>(E + item * d);
    const float4* p4 = reinterpret_cast<const float4*>(P + ps * d);
    float4* o4 = reinterpret_cast<float4*>(x + b * d);
    for (int v = lane; v < (d >> 2); v += 32) {
        float4 a = __ldg(h4 + v), bb = __ldg(e4 + v), c = __ldg(p4 + v), r;
        r.x = __fadd_rn(__fmul_rn(__fadd_rn(a.x, bb.x), scale), c.x);
        r.y = __fadd_rn(__fmul_rn(__fadd_rn(a.y, bb.y), scale), c.y);
        r.z = __fadd_rn(__fmul_rn(__fadd_rn(a.z, bb.z), scale), c.z);
        r.w = __fadd_rn(__fmul_rn(__fadd_rn(a.w, bb.w), scale), c.w);
        o4[v] = r;
    }
}

// ---- backward -----------------------------------------------------------------------------------
constexpr int kTokChunk = 64;      // tokens per dense-bin chunk
constexpr int kSlab = 128;         // features per dense-bin block

__global__ void mark_kernel(const int64_t* __restrict__ seq, int64_t n_tok, int64_t pad, int32_t* __restrict__ first,
                            int32_t* __restrict__ cnt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tok) return;
    const int64_t k = seq[t];
    if (k == pad) return;
    atomicMin(first + k, (int32_t)t);
    atomicAdd(cnt + k, 1);
}

// the elected (first) token of every item referenced more than once reserves cnt[item] slots of the token list
__global__ void list_alloc_kernel(const int64_t* __restrict__ seq, int64_t n_tok, int64_t pad,
                                  const int32_t* __restrict__ first, const int32_t* __restrict__ cnt,
                                  int32_t* __restrict__ start, int32_t* __restrict__ total) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tok) return;
    const int64_t k = seq[t];
    if (k == pad || first[k] != (int32_t)t || cnt[k] < 2) return;
    start[k] = atomicAdd(total, cnt[k]);
}
// every token of such an item appends itself to the item's run (order inside the run is arbitrary: sorted later)
__global__ void list_fill_kernel(const int64_t* __restrict__ seq, int64_t n_tok, int64_t pad,
                                 const int32_t* __restrict__ cnt, const int32_t* __restrict__ start,
                                 int32_t* __restrict__ fill, int32_t* __restrict__ list) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tok) return;
    const int64_t k = seq[t];
    if (k == pad || cnt[k] < 2) return;
    list[start[k] + atomicAdd(fill + k, 1)] = (int32_t)t;
}

// one warp per token; lane l owns features l, l+32, ... (d <= 512)
__global__ void item_rows_kernel(const float* __restrict__ dx, const int64_t* __restrict__ seq, int64_t n_tok, int d,
                                 int64_t pad, float scale, Dropout dr, const int32_t* __restrict__ first,
                                 const int32_t* __restrict__ cnt, const int32_t* __restrict__ start,
                                 const int32_t* __restrict__ list, float* d_hi, float* d_E) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tok) return;
    const int64_t key = seq[t];
    if (key == pad || first[key] != (int32_t)t) return;           // warp-uniform
    const int nper = (d + 31) >> 5;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    auto add_row = [&](int64_t tok) {
        const float* row = dx + tok * d;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int f = lane + 32 * k;
            if (k < nper && f < d) acc[k] += scale * __ldg(row + f) * drop_scale(dr, (uint64_t)tok * d + f);
        }
    };
    const int n_ref = cnt[key];
    if (n_ref == 1) {
        add_row(t);
    } else if (n_ref <= 32) {
        // the item's run of the token list: one entry per lane, rank by counting (entries are distinct), then the
        // rows are added in ascending token order
        const int32_t mine = lane < n_ref ? list[start[key] + lane] : 0x7fffffff;
        int rank = 0;
        for (int j = 0; j < n_ref; ++j) rank += __shfl_sync(0xffffffffu, mine, j) < mine ? 1 : 0;
        for (int r = 0; r < n_ref; ++r) {
            const unsigned m = __ballot_sync(0xffffffffu, lane < n_ref && rank == r);
            add_row(__shfl_sync(0xffffffffu, mine, __ffs(m) - 1));
        }
    } else {
        // scan the token list from t on (earlier tokens cannot match); 8 independent 256-byte loads in flight
        for (int64_t base = t & ~(int64_t)31; base < n_tok; base += 32 * 8) {
            int64_t k[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t j = base + 32 * u + lane;
                k[u] = j < n_tok ? __ldg(seq + j) : -1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t j = base + 32 * u + lane;
                unsigned m = __ballot_sync(0xffffffffu, j >= t && k[u] == key);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    add_row(base + 32 * u + b);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int f = lane + 32 * k;
        if (k < nper && f < d) {
            d_hi[key * d + f] += acc[k];
            if (d_E) d_E[key * d + f] += acc[k];
        }
    }
}

// grid (chunks, feature slabs), block (kSlab, kSub) threads, dynamic smem (kSub * n_bins * kSlab floats).
// bins 0 .. L-1: positional rows (unscaled g); bin L: the pad item row (scale * g).  The chunk's tokens are
// cut into kSub consecutive sub-ranges, one per threadIdx.y, each with its own bins (a quarter of the serial
// chain of dependent loads); the kSub bin sets are then added in sub-range order = token order.
constexpr int kSub = 4;            // (fewer, = blockDim.y, when len_max is too long for 4 bin sets in shared memory)
__global__ void bin_partial_kernel(const float* __restrict__ dx, const int64_t* __restrict__ seq,
                                   const int64_t* __restrict__ pos, int64_t n_tok, int d, int L, int64_t pad,
                                   float scale, Dropout dr, float* __restrict__ partial) {
    extern __shared__ float bins[];
    const int n_bins = L + 1;
    const int f = blockIdx.y * kSlab + threadIdx.x;
    float* mine = bins + (size_t)threadIdx.y * n_bins * kSlab;
    for (int b = 0; b < n_bins; ++b) mine[b * kSlab + threadIdx.x] = 0.f;
    const int n_sub = blockDim.y;
    const int per = kTokChunk / n_sub;
    const int64_t t0 = (int64_t)blockIdx.x * kTokChunk + threadIdx.y * per;
    const int64_t t1 = t0 + per < n_tok ? t0 + per : n_tok;
    if (f < d) {
        for (int64_t t = t0; t < t1; ++t) {
            const float g = __ldg(dx + t * d + f) * drop_scale(dr, (uint64_t)t * d + f);
            const int64_t ps = pos[t];
            if (ps >= 0 && ps < L) mine[ps * kSlab + threadIdx.x] += g;
            if (seq[t] == pad) mine[L * kSlab + threadIdx.x] += scale * g;
        }
    }
    __syncthreads();
    if (f < d) {
        float* out = partial + ((int64_t)blockIdx.x * n_bins) * d + f;
        for (int b = threadIdx.y; b < n_bins; b += n_sub) {
            float s = 0.f;
            for (int k = 0; k < n_sub; ++k) s += bins[((size_t)k * n_bins + b) * kSlab + threadIdx.x];
            out[(int64_t)b * d] = s;
        }
    }
}

// thread per (bin, feature): add the chunk partials in chunk order
__global__ void bin_reduce_kernel(const float* __restrict__ partial, int64_t n_chunks, int d, int L, int64_t pad,
                                  float* d_P, float* d_hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_bins = L + 1;
    if (i >= (int64_t)n_bins * d) return;
    const int b = (int)(i / d), f = (int)(i % d);
    float s = 0.f;
    for (int64_t c = 0; c < n_chunks; ++c) s += partial[(c * n_bins + b) * d + f];
    if (b < L) {
        if (d_P) d_P[(int64_t)b * d + f] += s;
    } else {
        d_hi[pad * d + f] += s;                   // the pad row gets no direct-embedding gradient (padding_idx)
    }
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

int c2dsr_gather_select_fwd(const float* hi, const float* E, const float* P, const int64_t* seq, const int64_t* pos,
                            const int64_t* sel, float* x, int64_t n_seq, int L, int d, float scale, int64_t pad_idx,
                            void* stream) {
    if (n_seq < 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d % 4 == 0, "d must be a positive multiple of 4");
    gather_select_fwd_kernel<<<(unsigned)ceil_div(n_seq + 1, 8), 256, 0, (cudaStream_t)stream>>>(hi, E, P, seq, pos, sel, x,
                                                                                                 n_seq, L, d, scale, pad_idx);
    note_launches(1);
    return check_launch("gather_select_fwd");
}

int64_t c2dsr_gather_bwd_workspace_bytes(int64_t n_tok, int d, int64_t n_rows, int len_max) {
    if (n_tok <= 0) return 256;
    // first / cnt / start / fill per item, the token list + its allocation counter, then the dense-bin partials
    return align_up(4 * n_rows * 4, 256) + align_up((n_tok + 64) * 4, 256) +
           ceil_div(n_tok, kTokChunk) * (int64_t)(len_max + 1) * d * 4 + 256;
}

int c2dsr_gather_bwd(const float* dx, const int64_t* seq, const int64_t* pos, float* d_hi, float* d_E, float* d_P,
                     int64_t n_tok, int d, int64_t n_rows, int len_max, int64_t pad_idx, float scale, float p,
                     uint64_t seed, uint64_t tag, void* workspace, int64_t workspace_bytes, void* stream) {
    if (n_tok <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(d > 0 && d <= 512, "d must be in (0, 512]");
    C2DSR_REQUIRE(n_tok < (1ll << 31) && n_rows > 0 && len_max > 0, "bad sizes");
    int n_sub = kSub;
    while (n_sub > 1 && n_sub * (len_max + 1) * kSlab * 4 > 200 * 1024) n_sub >>= 1;
    const int smem = n_sub * (len_max + 1) * kSlab * 4;
    C2DSR_REQUIRE(smem <= 200 * 1024, "len_max too large for the dense-bin reduction (<= 399)");
    if (workspace_bytes < c2dsr_gather_bwd_workspace_bytes(n_tok, d, n_rows, len_max)) {
        set_error("gather_bwd: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    const Dropout dr = make_dropout(p, seed, tag);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* first = (int32_t*)workspace;
    int32_t* cnt = first + n_rows;
    int32_t* start = cnt + n_rows;
    int32_t* fill = start + n_rows;
    int32_t* list = (int32_t*)((char*)workspace + align_up(4 * n_rows * 4, 256));
    int32_t* total = list + n_tok;
    float* partial = (float*)((char*)list + align_up((n_tok + 64) * 4, 256));
    cudaMemsetAsync(first, 0x7f, n_rows * 4, st);
    cudaMemsetAsync(cnt, 0, 3 * n_rows * 4, st);                  // cnt, start, fill
    cudaMemsetAsync(total, 0, 4, st);
    const unsigned tb = (unsigned)ceil_div(n_tok, 256);
    mark_kernel<<<tb, 256, 0, st>>>(seq, n_tok, pad_idx, first, cnt);
    list_alloc_kernel<<<tb, 256, 0, st>>>(seq, n_tok, pad_idx, first, cnt, start, total);
    list_fill_kernel<<<tb, 256, 0, st>>>(seq, n_tok, pad_idx, cnt, start, fill, list);
    item_rows_kernel<<<(unsigned)ceil_div(n_tok, 8), 256, 0, st>>>(dx, seq, n_tok, d, pad_idx, scale, dr, first, cnt,
                                                                  start, list, d_hi, d_E);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(bin_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    const int64_t n_chunks = ceil_div(n_tok, kTokChunk);
    bin_partial_kernel<<<dim3((unsigned)n_chunks, (unsigned)ceil_div(d, kSlab)), dim3(kSlab, n_sub), smem, st>>>(
        dx, seq, pos, n_tok, d, len_max, pad_idx, scale, dr, partial);
    bin_reduce_kernel<<<(unsigned)ceil_div((int64_t)(len_max + 1) * d, 256), 256, 0, st>>>(partial, n_chunks, d,
                                                                                          len_max, pad_idx, d_P, d_hi);
    note_launches(6);
    return check_launch("gather_bwd");
}

}  // extern "C"
