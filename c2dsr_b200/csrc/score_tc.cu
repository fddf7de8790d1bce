// K4b tensor-core path (tcgen05 + TMEM + TMA): placeholder entry points until the kernel lands.
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

// fp32 -> (hi, lo) bf16 split: hi = bf16(x), lo = bf16(x - hi).  hi*hi' + hi*lo' + lo*hi' reproduces the
// fp32 product to ~2^-17 relative.
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);   // inf / nan
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

__global__ void split_bf16_kernel(const float* __restrict__ X, int64_t rows, int d, int64_t ld_out,
                                  uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int64_t total = rows * (int64_t)d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int c = (int)(i % d);
        const float x = X[i];
        const uint16_t h = f32_to_bf16_rn(x);
        const float hf = __uint_as_float((uint32_t)h << 16);
        hi[r * ld_out + c] = h;
        if (lo) lo[r * ld_out + c] = f32_to_bf16_rn(x - hf);
    }
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int c2dsr_split_bf16(const float* X, int64_t rows, int d, int64_t ld_out, uint16_t* hi, uint16_t* lo,
                     void* stream) {
    if (rows <= 0) return C2DSR_OK;
    int64_t blocks = ceil_div(rows * (int64_t)d, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, rows, d, ld_out, hi, lo);
    note_launches(1);
    return check_launch("split_bf16");
}

int64_t c2dsr_score_tc_workspace_bytes(int64_t n_q, int64_t n_shard, int d) {
    (void)n_q; (void)n_shard; (void)d;
    return 256;
}

int c2dsr_score_target_tc(const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, const float*,
                          const int64_t*, int64_t, int64_t, int64_t, int, int, float*, void*, int64_t, void*) {
    set_error("score_target_tc: tensor-core path not built in this revision");
    return C2DSR_ERR_ARG;
}

int c2dsr_score_count_tc(const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, const float*,
                         const float*, const int64_t*, int64_t, int64_t, int64_t, int, int, int32_t*, float*,
                         int64_t, void*, int64_t, void*) {
    set_error("score_count_tc: tensor-core path not built in this revision");
    return C2DSR_ERR_ARG;
}

}  // extern "C"
