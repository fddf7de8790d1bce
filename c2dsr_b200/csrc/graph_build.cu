// Adjacency builder on the device (utils/graph.py:33-96 of the reference):
// raw directed transitions (src, dst) -> duplicates summed, rows normalised (D^-1 A), CSR of A or of A^T.
//
// Offline work (run once per data set), so the design goal is exactness and determinism, not speed:
//   1. 64-bit keys (row << 32 | col) are sorted by a least-significant-digit radix sort, 8 bits per pass,
//      only over the bytes that n_rows needs.  Each pass is three kernels -- per-chunk digit histogram, one
//      exclusive scan over the (digit, chunk) table, stable scatter (a thread's rank inside its chunk is
//      the number of earlier elements of the chunk with the same digit) -- so no float or order-dependent
//      atomics are involved and the result is the unique sorted order.
//   2. run heads are flagged and scanned: unique (row, col) pairs with their multiplicity.
//   3. val = (1 / rowsum[src]) * count in fp32 with separately rounded division and multiplication, which
//      is bit-identical to the reference's scipy normalisation (inv = 1 / rowsum; val = inv * count).
//   The transposed CSR uses keys (dst << 32 | src) and still divides by the row sum of the SOURCE.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

constexpr int kSortChunk = 256;

__global__ void gb_keys_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t n,
                               int transpose, uint64_t* __restrict__ keys, int32_t* __restrict__ rowsum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = (uint32_t)src[i], d = (uint32_t)dst[i];
    keys[i] = transpose ? ((uint64_t)d << 32 | s) : ((uint64_t)s << 32 | d);
    atomicAdd(rowsum + s, 1);                                    // integer: order independent
}

// block per chunk of 256 keys: hist[digit * n_chunks + chunk]
__global__ void gb_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int64_t n_chunks,
                               int32_t* __restrict__ hist) {
    __shared__ int cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kSortChunk + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & 255], 1);
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_chunks + blockIdx.x] = cnt[threadIdx.x];
}

// in-place exclusive scan of n int32 values by ONE block (n is a few million at most); total -> *total_out
__global__ void gb_scan_kernel(int32_t* __restrict__ data, int64_t n, int32_t* __restrict__ total_out) {
    __shared__ int64_t part[1024];
    const int64_t strip = (n + blockDim.x - 1) / blockDim.x;
    const int64_t b = strip * threadIdx.x, e = b + strip < n ? b + strip : n;
    int64_t s = 0;
    for (int64_t i = b; i < e; ++i) s += data[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            const int64_t v = part[t];
            part[t] = run;
            run += v;
        }
        if (total_out) *total_out = (int32_t)run;
    }
    __syncthreads();
    int64_t run = part[threadIdx.x];
    for (int64_t i = b; i < e; ++i) {
        const int32_t v = data[i];
        data[i] = (int32_t)run;
        run += v;
    }
}

// stable scatter of one pass: block per chunk
__global__ void gb_scatter_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n, int shift,
                                  int64_t n_chunks, const int32_t* __restrict__ offs) {
    __shared__ int dig[256];
    const int64_t i = (int64_t)blockIdx.x * kSortChunk + threadIdx.x;
    uint64_t k = 0;
    int d = -1;
    if (i < n) {
        k = in[i];
        d = (int)((k >> shift) & 255);
    }
    dig[threadIdx.x] = d;
    __syncthreads();
    if (i >= n) return;
    int rank = 0;
    for (int t = 0; t < (int)threadIdx.x; ++t) rank += dig[t] == d;
    out[offs[(int64_t)d * n_chunks + blockIdx.x] + rank] = k;
}

__global__ void gb_heads_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// idx = exclusive scan of the head flags.  For each head: record its position and bump its row's entry count.
__global__ void gb_compact_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ pos, int32_t* __restrict__ col, int32_t* __restrict__ row_cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = i == 0 || keys[i] != keys[i - 1];
    if (!head) return;
    const int32_t u = idx[i];
    pos[u] = (int32_t)i;
    col[u] = (int32_t)(keys[i] & 0xffffffffu);
    atomicAdd(row_cnt + (int32_t)(keys[i] >> 32), 1);
}

__global__ void gb_values_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ pos,
                                 const int32_t* __restrict__ n_unique, int64_t n, int transpose,
                                 const int32_t* __restrict__ rowsum, float* __restrict__ val) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t nu = *n_unique;
    if (u >= nu) return;
    const int32_t p = pos[u];
    const int32_t next = u + 1 < nu ? pos[u + 1] : (int32_t)n;
    const uint64_t k = keys[p];
    const int32_t s = transpose ? (int32_t)(k & 0xffffffffu) : (int32_t)(k >> 32);
    const float inv = __fdiv_rn(1.f, (float)rowsum[s]);
    val[u] = __fmul_rn(inv, (float)(next - p));
}

struct GbLayout {
    uint64_t *ka, *kb;
    int32_t *hist, *flag, *pos, *rowsum;
    int64_t bytes;
};
static GbLayout gb_layout(void* ws, int64_t n, int64_t n_rows) {
    const int64_t n_chunks = ceil_div(n > 0 ? n : 1, kSortChunk);
    char* p = (char*)ws;
    GbLayout L;
    auto take = [&](int64_t b) {
        char* q = p;
        p += align_up(b, 256);
        return q;
    };
    L.ka = (uint64_t*)take(n * 8);
    L.kb = (uint64_t*)take(n * 8);
    L.hist = (int32_t*)take(256 * n_chunks * 4);
    L.flag = (int32_t*)take(n * 4);
    L.pos = (int32_t*)take((n + 1) * 4);
    L.rowsum = (int32_t*)take(n_rows * 4);
    L.bytes = p - (char*)ws;
    return L;
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int64_t c2dsr_graph_build_workspace_bytes(int64_t n_edges, int64_t n_rows) {
    return gb_layout(nullptr, n_edges, n_rows).bytes + 1024;
}

int c2dsr_graph_build(const int32_t* src, const int32_t* dst, int64_t n_edges, int64_t n_rows, int transpose,
                      int32_t* rowptr, int32_t* col, float* val, int32_t* nnz_out, void* workspace,
                      int64_t workspace_bytes, void* stream) {
    C2DSR_REQUIRE(n_rows > 0 && n_rows < (1ll << 31) && n_edges >= 0 && n_edges < (1ll << 31), "bad sizes");
    if (workspace_bytes < c2dsr_graph_build_workspace_bytes(n_edges, n_rows)) {
        set_error("graph_build: workspace too small");
        return C2DSR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(rowptr, 0, (n_rows + 1) * 4, st);
    cudaMemsetAsync(nnz_out, 0, 4, st);
    if (n_edges == 0) return check_launch("graph_build");
    const GbLayout L = gb_layout(workspace, n_edges, n_rows);
    const int64_t n_chunks = ceil_div(n_edges, kSortChunk);
    const unsigned eb = (unsigned)ceil_div(n_edges, 256);
    cudaMemsetAsync(L.rowsum, 0, n_rows * 4, st);
    gb_keys_kernel<<<eb, 256, 0, st>>>(src, dst, n_edges, transpose, L.ka, L.rowsum);
    int bytes = 1;
    while (bytes < 4 && (n_rows - 1) >> (8 * bytes)) ++bytes;
    uint64_t* in = L.ka;
    uint64_t* out = L.kb;
    int launches = 1;
    for (int half = 0; half < 2; ++half)
        for (int b = 0; b < bytes; ++b) {
            const int shift = 32 * half + 8 * b;
            gb_hist_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(in, n_edges, shift, n_chunks, L.hist);
            gb_scan_kernel<<<1, 1024, 0, st>>>(L.hist, 256 * n_chunks, nullptr);
            gb_scatter_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(in, out, n_edges, shift, n_chunks, L.hist);
            uint64_t* t = in;
            in = out;
            out = t;
            launches += 3;
        }
    gb_heads_kernel<<<eb, 256, 0, st>>>(in, n_edges, L.flag);
    gb_scan_kernel<<<1, 1024, 0, st>>>(L.flag, n_edges, nnz_out);
    gb_compact_kernel<<<eb, 256, 0, st>>>(in, L.flag, n_edges, L.pos, col, rowptr);
    gb_scan_kernel<<<1, 1024, 0, st>>>(rowptr, n_rows + 1, nullptr);
    gb_values_kernel<<<eb, 256, 0, st>>>(in, L.pos, nnz_out, n_edges, transpose, L.rowsum, val);
    note_launches(launches + 5);
    return check_launch("graph_build");
}

}  // extern "C"
