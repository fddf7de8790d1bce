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
    uint32_t k0, k1;     // per-(seed, call-site tag) keys of the two 32-bit hash words of a quad
    const uint2* dyn;    // optional device-resident per-step key words XORed into k0 / k1 (see C2DSR_SEED_INDIRECT)
};

constexpr uint64_t kSeedIndirect = 1ull << 63;   // tag bit: ``seed`` is a device pointer to two uint32 key words

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
    d.dyn = nullptr;
    if (tag & kSeedIndirect) {            // the per-step part of the key lives on the device (CUDA-graph replay)
        d.dyn = reinterpret_cast<const uint2*>(seed);
        seed = 0;
        tag &= ~kSeedIndirect;
    }
    const uint64_t key = host_mix64(seed ^ host_mix64(tag));
    d.k0 = (uint32_t)key;
    d.k1 = (uint32_t)(key >> 32);
    if (d.k1 == d.k0) d.k1 ^= 0x9E3779B9u;
    return d;
}

// Counter-based mask.  Elements are hashed four at a time ("quad" = idx >> 2): two 32-bit words
//   w_h = fmix32(quad ^ k_h),  h = 0, 1          (fmix32 = the murmur3 finaliser: full avalanche, 7 ALU ops)
// give four 16-bit lanes; element idx keeps iff lane (idx & 3) >= thresh (= round(p * 65536)), and 1/(1-p) is
// taken from the quantised p so the mask stays unbiased.  All kernels use these helpers, so the mask of
// element idx is the same whether it is generated one element or four at a time, in forward or backward.
// (A 64-bit splitmix per quad was 2-3x the instructions and made the SpMM gather issue-bound.)
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t quad_counter(uint64_t quad) {
    return (uint32_t)quad ^ ((uint32_t)(quad >> 32) * 0x9E3779B1u);
}

// multiplicative mask value: 0 or 1/(1-p)
__device__ __forceinline__ float drop_scale(const Dropout& d, uint64_t idx) {
    if (d.p == 0.f) return 1.f;
    uint32_t k = (idx & 2) ? d.k1 : d.k0;
    if (d.dyn) {
        const uint2 kk = __ldg(d.dyn);
        k ^= (idx & 2) ? kk.y : kk.x;
    }
    const uint32_t w = fmix32(quad_counter(idx >> 2) ^ k);
    const uint32_t h = (idx & 1) ? (w >> 16) : (w & 0xffffu);
    return h >= d.thresh ? d.inv_keep : 0.f;
}

// four consecutive elements starting at a multiple of 4
__device__ __forceinline__ void drop_scale4(const Dropout& d, uint64_t base, float4& v) {
    if (d.p == 0.f) return;
    const uint32_t q = quad_counter(base >> 2);
    uint32_t k0 = d.k0, k1 = d.k1;
    if (d.dyn) {
        const uint2 kk = __ldg(d.dyn);
        k0 ^= kk.x;
        k1 ^= kk.y;
    }
    const uint32_t w0 = fmix32(q ^ k0), w1 = fmix32(q ^ k1);
    v.x *= (w0 & 0xffffu) >= d.thresh ? d.inv_keep : 0.f;
    v.y *= (w0 >> 16) >= d.thresh ? d.inv_keep : 0.f;
    v.z *= (w1 & 0xffffu) >= d.thresh ? d.inv_keep : 0.f;
    v.w *= (w1 >> 16) >= d.thresh ? d.inv_keep : 0.f;
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
