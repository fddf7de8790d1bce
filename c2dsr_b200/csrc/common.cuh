// Shared device/host helpers for the c2dsr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define C2DSR_OK 0
#define C2DSR_ERR_ARG (-10001)
#define C2DSR_ERR_ARCH (-10002)
#define C2DSR_ERR_WORKSPACE (-10003)

namespace c2dsr {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
void note_launches(int n);                  // bookkeeping for c2dsr_launch_count()          // cudaGetLastError -> 0 / -(cudaError_t), records message

#define C2DSR_REQUIRE(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            c2dsr::set_error("%s: %s", __func__, msg);             \
            return C2DSR_ERR_ARG;                                  \
        }                                                          \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// Counter-based dropout mask.  keep(seed, tag, idx) is a pure function, so backward kernels
// regenerate the forward mask from the same (seed, tag, idx) instead of storing it.
// ---------------------------------------------------------------------------------------------
struct Dropout {
    float p;             // drop probability; 0 disables
    float inv_keep;      // 1 / (1 - p)
    uint32_t thresh;     // keep iff the element's 16-bit hash lane >= thresh
    uint64_t key;        // seed mixed with the call-site tag
};

static inline uint64_t host_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline Dropout make_dropout(float p, uint64_t seed, uint64_t tag) {
    Dropout d;
    d.p = (p > 0.f && p < 1.f) ? p : 0.f;
    double t = (double)d.p * 65536.0 + 0.5;
    d.thresh = (uint32_t)(t > 65535.0 ? 65535.0 : t);                 // 16-bit threshold
    d.inv_keep = (float)(1.0 / (1.0 - (double)d.thresh / 65536.0));
    d.key = host_mix64(seed ^ host_mix64(tag));
    return d;
}

// One 64-bit mix serves FOUR consecutive elements: element idx keeps iff the 16-bit lane (idx & 3) of
// mix(idx >> 2) is >= thresh (= round(p * 65536)); 1/(1-p) is taken from the quantised p so the mask stays
// unbiased.  All kernels use these two helpers, so the mask of element idx is the same whether it is
// generated one element or four at a time.
__device__ __forceinline__ uint64_t mix64(uint64_t key, uint64_t quad) {
    uint64_t z = key + quad * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// multiplicative mask value: 0 or 1/(1-p)
__device__ __forceinline__ float drop_scale(const Dropout& d, uint64_t idx) {
    if (d.p == 0.f) return 1.f;
    const uint64_t z = mix64(d.key, idx >> 2);
    const uint32_t h = (uint32_t)(z >> (16 * (idx & 3))) & 0xffffu;
    return h >= d.thresh ? d.inv_keep : 0.f;
}

// four consecutive elements starting at a multiple of 4
__device__ __forceinline__ void drop_scale4(const Dropout& d, uint64_t base, float4& v) {
    if (d.p == 0.f) return;
    const uint64_t z = mix64(d.key, base >> 2);
    v.x *= ((uint32_t)z & 0xffffu) >= d.thresh ? d.inv_keep : 0.f;
    v.y *= ((uint32_t)(z >> 16) & 0xffffu) >= d.thresh ? d.inv_keep : 0.f;
    v.z *= ((uint32_t)(z >> 32) & 0xffffu) >= d.thresh ? d.inv_keep : 0.f;
    v.w *= (uint32_t)(z >> 48) >= d.thresh ? d.inv_keep : 0.f;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace c2dsr
