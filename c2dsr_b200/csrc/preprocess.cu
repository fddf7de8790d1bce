// Preprocessor on the device (reference: dataloader.py:60-228).
// Everything the reference derives from a time-sorted item list is integer bookkeeping per sequence -- domain views,
// positions, next-item targets, masks, left padding -- and is done here by one thread per sequence, bit-identical to
// the reference.  The two random ingredients (one corruption draw per position, dataloader.py:80,85; the sampled
// negatives, dataloader.py:216-224) are passed in: they come from Python's `random` stream on the host, consumed in
// the reference's order, because a seeded run must see the reference's numbers.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

// fields [n, 14, L]; keep [n] = 1 iff the reference keeps the sequence (both domains have a target)
__global__ void preprocess_train_kernel(const int64_t* __restrict__ items, const int64_t* __restrict__ offs,
                                        const int64_t* __restrict__ draws, int64_t n, int64_t na, int64_t nb, int L,
                                        int64_t* __restrict__ fields, uint8_t* __restrict__ keep) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int64_t pad = na + nb;
    const int64_t* it = items + offs[u];
    const int m = (int)(offs[u + 1] - offs[u]) - 1;          // inputs; it[1 .. m] are the step targets
    const int64_t* dr = draws + (offs[u] - u);               // one draw per input position
    int64_t* f = fields + u * 14 * (int64_t)L;
    auto row = [&](int r) { return f + (int64_t)r * L; };
    const int64_t fill[14] = {pad, pad, pad, 0, 0, 0, na, nb, na, nb, 0, 0, pad, pad};
    for (int r = 0; r < 14; ++r)
        for (int j = 0; j < L; ++j) row(r)[j] = fill[r];
    const int o = L - m;
    int ca = 0, cb = 0;
    for (int i = 0; i < m; ++i) {
        const int64_t x = it[i], g = it[i + 1];
        const bool is_a = x < na;
        row(0)[o + i] = x;
        row(3)[o + i] = i + 1;
        if (is_a) {
            row(1)[o + i] = x;
            row(4)[o + i] = ++ca;
            row(12)[o + i] = x;
            row(13)[o + i] = dr[i];
        } else {
            row(2)[o + i] = x;
            row(5)[o + i] = ++cb;
            row(12)[o + i] = dr[i];
            row(13)[o + i] = x;
        }
        row(6)[o + i] = g < na ? g : na;
        row(7)[o + i] = g >= na ? g - na : nb;
    }
    // next same-domain item is the step target; the last item of a domain takes the final target if that belongs
    // to the domain (A: < na, B: > na, strictly -- Q16), else it is blanked from the domain's input
    const int64_t last = it[m];
    int64_t next_a = last < na ? last : -1, next_b = last > na ? last - na : -1;
    int n_a = 0, n_b = 0;
    for (int i = m - 1; i >= 0; --i) {
        const int64_t x = it[i];
        if (x < na) {
            if (next_a >= 0) {
                row(8)[o + i] = next_a;
                row(10)[o + i] = next_a != na ? 1 : 0;
                n_a += next_a != na ? 1 : 0;
            } else {
                row(1)[o + i] = pad;
                row(4)[o + i] = 0;
            }
            next_a = x;
        } else {
            if (next_b >= 0) {
                row(9)[o + i] = next_b;
                row(11)[o + i] = next_b != nb ? 1 : 0;
                n_b += next_b != nb ? 1 : 0;
            } else {
                row(2)[o + i] = pad;
                row(5)[o + i] = 0;
            }
            next_b = x - na;
        }
    }
    keep[u] = (n_a > 0 && n_b > 0) ? 1 : 0;
}

// six [n, 6, L], four [n, 4] (idx_last_a, idx_last_b, domain, target), neg [n, n_neg] (picks shifted past the target)
__global__ void preprocess_eval_kernel(const int64_t* __restrict__ items, const int64_t* __restrict__ offs, int64_t n,
                                       int64_t na, int64_t nb, int L, int n_neg, int64_t* __restrict__ six,
                                       int64_t* __restrict__ four, int64_t* __restrict__ neg) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int64_t pad = na + nb;
    const int64_t* it = items + offs[u];
    const int m = (int)(offs[u + 1] - offs[u]) - 1;
    int64_t* f = six + u * 6 * (int64_t)L;
    auto row = [&](int r) { return f + (int64_t)r * L; };
    for (int r = 0; r < 6; ++r)
        for (int j = 0; j < L; ++j) row(r)[j] = r < 3 ? pad : 0;
    const int o = L - m;
    int ca = 0, cb = 0;
    int64_t la = -1, lb = -1;
    for (int i = 0; i < m; ++i) {
        const int64_t x = it[i];
        row(0)[o + i] = x;
        row(3)[o + i] = i + 1;
        if (x < na) {
            row(1)[o + i] = x;
            row(4)[o + i] = ++ca;
            la = o + i;
        } else {
            row(2)[o + i] = x;
            row(5)[o + i] = ++cb;
            lb = o + i;
        }
    }
    const int64_t last = it[m];
    const int64_t g = last < na ? last : last - na;
    four[u * 4 + 0] = la;
    four[u * 4 + 1] = lb;
    four[u * 4 + 2] = last < na ? 0 : 1;
    four[u * 4 + 3] = g;
    int64_t* ng = neg + u * (int64_t)n_neg;
    for (int j = 0; j < n_neg; ++j) ng[j] = ng[j] < g ? ng[j] : ng[j] + 1;     // in place: picks -> ids without the target
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int c2dsr_preprocess_train(const int64_t* items, const int64_t* offs, const int64_t* draws, int64_t n_seq,
                           int64_t n_item_a, int64_t n_item_b, int len_max, int64_t* fields, uint8_t* keep,
                           void* stream) {
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(len_max > 0 && n_item_a > 0 && n_item_b > 0, "bad sizes");
    preprocess_train_kernel<<<(unsigned)ceil_div(n_seq, 128), 128, 0, (cudaStream_t)stream>>>(
        items, offs, draws, n_seq, n_item_a, n_item_b, len_max, fields, keep);
    note_launches(1);
    return check_launch("preprocess_train");
}

int c2dsr_preprocess_eval(const int64_t* items, const int64_t* offs, int64_t n_seq, int64_t n_item_a,
                          int64_t n_item_b, int len_max, int n_neg, int64_t* six, int64_t* four, int64_t* neg,
                          void* stream) {
    if (n_seq <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(len_max > 0 && n_item_a > 0 && n_item_b > 0 && n_neg >= 0, "bad sizes");
    preprocess_eval_kernel<<<(unsigned)ceil_div(n_seq, 128), 128, 0, (cudaStream_t)stream>>>(
        items, offs, n_seq, n_item_a, n_item_b, len_max, n_neg, six, four, neg);
    note_launches(1);
    return check_launch("preprocess_eval");
}

}  // extern "C"
