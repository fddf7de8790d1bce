// Error plumbing, device check, small elementwise / reduction kernels and the fused AdamW-amsgrad step.
#include <stdarg.h>
#include <atomic>
#include <string.h>
#include <cstdlib>
#include "common.cuh"
#include "../../include/c2dsr_b200.h"

namespace c2dsr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return C2DSR_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -(int)e;
}

// ------------------------------------------------------------------------------------------
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                             int64_t n, float a, float b) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = y ? a * x[i] + b * y[i] : a * x[i];
}

// Column sums with a fixed summation order: block = 32 columns x 8 row lanes, rows strided by 8,
// then the 8 partials are added in lane order.
__global__ void colsum_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int64_t N,
                              float* __restrict__ out, int accumulate) {
    __shared__ float part[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < N)
        for (int64_t r = threadIdx.y; r < M; r += 8) s += X[r * ldx + c];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        out[c] = accumulate ? out[c] + t : t;
    }
}

// Two-phase variant for tall matrices: grid (column blocks, row chunks of kColChunk) writes chunk
// partials, a second kernel adds them in chunk order.  Same fixed order on every run.
constexpr int kColChunk = 128;
__global__ void colsum_partial_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int64_t N,
                                      float* __restrict__ partial) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * kColChunk;
    const int64_t r1 = r0 + kColChunk < M ? r0 + kColChunk : M;
    float s = 0.f;
    if (c < N) {
#pragma unroll 4
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += X[r * ldx + c];
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        partial[(int64_t)blockIdx.y * N + c] = t;
    }
}
// block (32, 8): row group y adds chunks y, y + 8, ... of column x, then the 8 group sums are added in group
// order -- a fixed order, and 8 x shorter dependent chains than one thread per column
__global__ void colsum_final_kernel(const float* __restrict__ partial, int64_t n_chunks, int64_t N,
                                    float* __restrict__ out, int accumulate) {
    __shared__ float part[8][33];
    const int64_t c = (int64_t)blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < N) {
#pragma unroll 4
        for (int64_t k = threadIdx.y; k < n_chunks; k += 8) s += partial[k * N + c];
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += part[g][threadIdx.x];
        out[c] = accumulate ? out[c] + t : t;
    }
}

int colsum_dispatch(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, int accumulate, void* ws,
                    int64_t ws_bytes, cudaStream_t st) {
    if (N <= 0) return C2DSR_OK;
    const int64_t n_chunks = ceil_div(M, kColChunk);
    if (ws && n_chunks > 1 && n_chunks * N * 4 <= ws_bytes) {
        colsum_partial_kernel<<<dim3((unsigned)ceil_div(N, 32), (unsigned)n_chunks), dim3(32, 8), 0, st>>>(
            X, ldx, M, N, (float*)ws);
        colsum_final_kernel<<<(unsigned)ceil_div(N, 32), dim3(32, 8), 0, st>>>((const float*)ws, n_chunks, N, out,
                                                                            accumulate);
        note_launches(2);
    } else {
        colsum_kernel<<<(unsigned)ceil_div(N, 32), dim3(32, 8), 0, st>>>(X, ldx, M, N, out, accumulate);
        note_launches(1);
    }
    return check_launch("colsum");
}

// Single-block fixed-order weighted sum.
__global__ void wsum_kernel(const float* __restrict__ x, const float* __restrict__ w, int64_t n,
                            float* __restrict__ out) {
    __shared__ float part[1024];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += w ? x[i] * w[i] : x[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = part[0];
}

// ------------------------------------------------------------------------------------------
// AdamW + amsgrad over a table of tensors.  grid = (chunks, n_tensors).
__global__ void step_set_kernel(c2dsr_step_state* s, int64_t step, float lr, bool set_step, bool set_lr) {
    if (set_step) s->step = (uint64_t)step;
    if (set_lr) s->lr = lr;
}

// advance to the next step: counter + the two per-step dropout key words derived from it
__global__ void step_begin_kernel(c2dsr_step_state* s, uint64_t seed_base) {
    const uint64_t t = s->step + 1;
    s->step = t;
    const uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
    s->key[0] = fmix32((lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u) ^ (uint32_t)seed_base);
    s->key[1] = fmix32(((lo + 0x632BE5ABu) * 0xC2B2AE3Du) ^ (hi * 0x27D4EB2Fu) ^ (uint32_t)(seed_base >> 32));
}

__global__ void adamw_kernel(const c2dsr_adam_tensor* __restrict__ table, float lr, float beta1, float beta2,
                             float eps, float wd, float inv_bc1, float sqrt_bc2,
                             const c2dsr_step_state* __restrict__ state) {
    if (state) {                       // step number and learning rate live on the device (CUDA-graph replay)
        __shared__ float bc[2];
        if (threadIdx.x == 0) {        // double-precision pow once per block, as the host path computes it
            const double n = (double)state->step;
            bc[0] = (float)(1.0 / (1.0 - pow((double)beta1, n)));
            bc[1] = (float)sqrt(1.0 - pow((double)beta2, n));
        }
        __syncthreads();
        lr = state->lr;
        inv_bc1 = bc[0];
        sqrt_bc2 = bc[1];
    }
    const c2dsr_adam_tensor t = table[blockIdx.y];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float decay = 1.f - lr * wd;
    const float step_size = lr * inv_bc1;   // lr / bias_correction1
    auto update = [&](float g, float& p, float& m, float& v, float& vm) {
        p *= decay;
        m = m + (g - m) * (1.f - beta1);                  // torch: exp_avg.lerp_(grad, 1 - beta1)
        v = v * beta2 + (1.f - beta2) * g * g;            // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
        vm = fmaxf(vm, v);
        p -= step_size * (m / (sqrtf(vm) / sqrt_bc2 + eps));
    };
    // 16-byte path for tensors whose arrays are all float4-aligned (every large tensor is)
    const uintptr_t bits = (uintptr_t)t.p | (uintptr_t)t.acc | (uintptr_t)t.m | (uintptr_t)t.v | (uintptr_t)t.vmax |
                           (uintptr_t)t.g;
    if ((bits & 15) == 0 && (t.n & 3) == 0) {
        const int64_t n4 = t.n >> 2;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 g = reinterpret_cast<const float4*>(t.acc)[i];
            if (t.g) {
                const float4 gn = reinterpret_cast<const float4*>(t.g)[i];
                g.x += gn.x; g.y += gn.y; g.z += gn.z; g.w += gn.w;
                reinterpret_cast<float4*>(t.acc)[i] = g;
            }
            float4 p = reinterpret_cast<float4*>(t.p)[i], m = reinterpret_cast<float4*>(t.m)[i];
            float4 v = reinterpret_cast<float4*>(t.v)[i], vm = reinterpret_cast<float4*>(t.vmax)[i];
            update(g.x, p.x, m.x, v.x, vm.x);
            update(g.y, p.y, m.y, v.y, vm.y);
            update(g.z, p.z, m.z, v.z, vm.z);
            update(g.w, p.w, m.w, v.w, vm.w);
            reinterpret_cast<float4*>(t.p)[i] = p;
            reinterpret_cast<float4*>(t.m)[i] = m;
            reinterpret_cast<float4*>(t.v)[i] = v;
            reinterpret_cast<float4*>(t.vmax)[i] = vm;
        }
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += stride) {
        float g = t.acc[i];
        if (t.g) {
            g += t.g[i];
            t.acc[i] = g;
        }
        float p = t.p[i] * decay;
        float m = t.m[i];
        m = m + (g - m) * (1.f - beta1);                  // torch: exp_avg.lerp_(grad, 1 - beta1)
        float v = t.v[i] * beta2 + (1.f - beta2) * g * g; // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
        float vm = fmaxf(t.vmax[i], v);
        float denom = sqrtf(vm) / sqrt_bc2 + eps;
        p -= step_size * (m / denom);
        t.p[i] = p;
        t.m[i] = m;
        t.v[i] = v;
        t.vmax[i] = vm;
    }
}

// ------------------------------------------------------------------------------------------
// Data-parallel optimiser step over peer memory: gradient all-reduce, AdamW and parameter broadcast in ONE kernel.
// Every rank owns 1 / world of each large tensor.  For an element of its slice the kernel
//   - sums the gradient over the ranks: one multimem.ld_reduce on the multicast address of the gradient blocks
//     (the NVSwitch adds the eight copies and returns one value), or, without multicast, peer loads in rank order;
//   - applies AdamW-amsgrad with the slice's state (same arithmetic as adamw_kernel);
//   - stores the new value into every rank's parameter block: one multimem.st, or world peer stores.
// NVLink carries each gradient byte once and each new parameter byte once per direction, reads and writes in
// flight together; the host brackets the launch with device-side barriers over all ranks (dist.PeerStep).
__device__ __forceinline__ float4 mc_ld_reduce(const float* addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 peer_ld(const float* addr) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(float* addr, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool MC, int U>
__global__ void __launch_bounds__(256) adamw_peer_kernel(const c2dsr_peer_tensor* __restrict__ table,
                                                         const __grid_constant__ c2dsr_peer_map map, float beta1,
                                                         float beta2, float eps, float wd,
                                                         const c2dsr_step_state* __restrict__ state) {
    __shared__ float bc[2];
    if (threadIdx.x == 0) {
        const double n = (double)state->step;
        bc[0] = (float)(1.0 / (1.0 - pow((double)beta1, n)));
        bc[1] = (float)sqrt(1.0 - pow((double)beta2, n));
    }
    __syncthreads();
    const float lr = state->lr, inv_bc1 = bc[0], sqrt_bc2 = bc[1];
    const c2dsr_peer_tensor job = table[blockIdx.y];
    const c2dsr_adam_tensor t = job.t;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float decay = 1.f - lr * wd;
    const float step_size = lr * inv_bc1;
    auto update = [&](float g, float& p, float& m, float& v, float& vm) {
        p *= decay;
        m = m + (g - m) * (1.f - beta1);
        v = v * beta2 + (1.f - beta2) * g * g;
        vm = fmaxf(vm, v);
        p -= step_size * (m / (sqrtf(vm) / sqrt_bc2 + eps));
    };
    const int64_t n4 = t.n >> 2;
    auto fetch = [&](int64_t i) {                      // the gradient of elements 4 i .. 4 i + 3, summed over the ranks
        const int64_t e = job.offset + i * 4;
        float4 gn;
        if (MC) {
            gn = mc_ld_reduce(map.grad_mc + e);
        } else {
            gn = peer_ld(map.grad[0] + e);
            for (int k = 1; k < map.world; ++k) {
                const float4 x = peer_ld(map.grad[k] + e);
                gn.x += x.x; gn.y += x.y; gn.z += x.z; gn.w += x.w;
            }
        }
        return gn;
    };
    auto apply = [&](int64_t i, float4 gn) {
        const int64_t e = job.offset + i * 4;
        float4 g = reinterpret_cast<const float4*>(t.acc)[i];
        g.x += gn.x; g.y += gn.y; g.z += gn.z; g.w += gn.w;
        reinterpret_cast<float4*>(t.acc)[i] = g;
        float4 p = reinterpret_cast<float4*>(t.p)[i], m = reinterpret_cast<float4*>(t.m)[i];
        float4 v = reinterpret_cast<float4*>(t.v)[i], vm = reinterpret_cast<float4*>(t.vmax)[i];
        update(g.x, p.x, m.x, v.x, vm.x);
        update(g.y, p.y, m.y, v.y, vm.y);
        update(g.z, p.z, m.z, v.z, vm.z);
        update(g.w, p.w, m.w, v.w, vm.w);
        reinterpret_cast<float4*>(t.m)[i] = m;
        reinterpret_cast<float4*>(t.v)[i] = v;
        reinterpret_cast<float4*>(t.vmax)[i] = vm;
        if (MC) {
            mc_st(map.param_mc + e, p);
        } else {
            for (int k = 0; k < map.world; ++k) peer_st(map.param[k] + e, p);
        }
    };
    // U remote fetches in flight per thread before the first is consumed (NVLink round trips are microseconds)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (U == 4) {
        for (; i + 3 * stride < n4; i += 4 * stride) {
            const float4 g0 = fetch(i), g1 = fetch(i + stride), g2 = fetch(i + 2 * stride), g3 = fetch(i + 3 * stride);
            apply(i, g0);
            apply(i + stride, g1);
            apply(i + 2 * stride, g2);
            apply(i + 3 * stride, g3);
        }
    }
    for (; i < n4; i += stride) apply(i, fetch(i));
}

}  // namespace c2dsr

using namespace c2dsr;

extern "C" {

int c2dsr_abi_version(void) { return C2DSR_ABI_VERSION; }
int64_t c2dsr_launch_count(void) { return (int64_t)launches_so_far(); }
const char* c2dsr_last_error(void) { return g_err; }

int c2dsr_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    if (major != 10) {
        set_error("c2dsr_b200 is built for sm_100a only; device %d has compute capability major %d", dev, major);
        return C2DSR_ERR_ARCH;
    }
    return C2DSR_OK;
}

int c2dsr_axpby(const float* x, const float* y, float* out, int64_t n, float a, float b, void* stream) {
    if (n <= 0) return C2DSR_OK;
    int blocks = (int)(ceil_div(n, 256) < 148 * 16 ? ceil_div(n, 256) : 148 * 16);
    axpby_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, out, n, a, b);
    note_launches(1);
    return check_launch("axpby");
}

int c2dsr_colsum(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, int accumulate, void* workspace,
                 int64_t workspace_bytes, void* stream) {
    return colsum_dispatch(X, ldx, M, N, out, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

int c2dsr_wsum(const float* x, const float* w, int64_t n, float* out, void* stream) {
    wsum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, w, n, out);
    note_launches(1);
    return check_launch("wsum");
}

int c2dsr_step_state_bytes(void) { return (int)sizeof(c2dsr_step_state); }

int c2dsr_step_state_set(void* state, int64_t step, float lr, void* stream) {
    C2DSR_REQUIRE(state != nullptr && step >= 0, "bad arguments");
    step_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<c2dsr_step_state*>(state), step, lr, step >= 0,
                                                       true);
    note_launches(1);
    return check_launch("step_state_set");
}

int c2dsr_step_state_set_lr(void* state, float lr, void* stream) {
    C2DSR_REQUIRE(state != nullptr, "bad arguments");
    step_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<c2dsr_step_state*>(state), 0, lr, false, true);
    note_launches(1);
    return check_launch("step_state_set_lr");
}

int c2dsr_step_begin(void* state, uint64_t seed_base, void* stream) {
    C2DSR_REQUIRE(state != nullptr, "bad arguments");
    step_begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<c2dsr_step_state*>(state), seed_base);
    note_launches(1);
    return check_launch("step_begin");
}

// background != 0: the launch runs beside higher-priority work (early optimiser steps): one short-lived CTA per
// 1 024 (4 096 for the peer kernel) elements instead of a resident grid-stride grid, so that the SMs are handed
// back to the step's own chain whenever it has a kernel to place
static int64_t adam_chunks(int64_t max_n, int64_t per_cta, int background) {
    int64_t chunks = ceil_div(max_n, per_cta);
    if (!background && chunks > 148 * 8) chunks = 148 * 8;
    if (chunks > 0x7fffffff) chunks = 0x7fffffff;
    return chunks < 1 ? 1 : chunks;
}

int c2dsr_adamw_amsgrad_dyn(const c2dsr_adam_tensor* table_dev, int n_tensors, int64_t max_n, const void* state,
                            float beta1, float beta2, float eps, float weight_decay, int background, void* stream) {
    if (n_tensors <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(state != nullptr, "state must not be NULL");
    const int64_t chunks = adam_chunks(max_n, 256 * 4, background);
    adamw_kernel<<<dim3((unsigned)chunks, (unsigned)n_tensors), 256, 0, (cudaStream_t)stream>>>(
        table_dev, 0.f, beta1, beta2, eps, weight_decay, 0.f, 0.f, reinterpret_cast<const c2dsr_step_state*>(state));
    note_launches(1);
    return check_launch("adamw_amsgrad_dyn");
}

int c2dsr_adamw_amsgrad(const c2dsr_adam_tensor* table_dev, int n_tensors, int64_t max_n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int step, void* stream) {
    if (n_tensors <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(step >= 1, "step must be >= 1");
    double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    int64_t chunks = ceil_div(max_n, 256 * 4);
    if (chunks > 148 * 8) chunks = 148 * 8;
    if (chunks < 1) chunks = 1;
    adamw_kernel<<<dim3((unsigned)chunks, (unsigned)n_tensors), 256, 0, (cudaStream_t)stream>>>(
        table_dev, lr, beta1, beta2, eps, weight_decay, (float)(1.0 / bc1), (float)sqrt(bc2), nullptr);
    note_launches(1);
    return check_launch("adamw_amsgrad");
}

int c2dsr_adamw_amsgrad_peer(const c2dsr_peer_tensor* table_dev, int n_tensors, int64_t max_n,
                             const c2dsr_peer_map* map, const void* state, float beta1, float beta2, float eps,
                             float weight_decay, int background, void* stream) {
    if (n_tensors <= 0) return C2DSR_OK;
    C2DSR_REQUIRE(state != nullptr && map != nullptr, "state and map must not be NULL");
    C2DSR_REQUIRE(map->world >= 1 && map->world <= C2DSR_MAX_PEERS && map->rank >= 0 && map->rank < map->world,
                  "bad peer map");
    C2DSR_REQUIRE(max_n % 4 == 0, "slices must be multiples of 4 elements");
    const int64_t chunks = adam_chunks(max_n, background ? 256 * 16 : 256 * 4, background);
    const dim3 grid((unsigned)chunks, (unsigned)n_tensors);
    const auto* st = reinterpret_cast<const c2dsr_step_state*>(state);
    // remote fetches in flight per thread: measured on B200s, 2 ranks 3.21 (1) vs 3.28 ms (4) per step, 4 ranks 3.46 (1)
    // vs 3.42 ms (4) -- the more peers, the longer the round trips that have to be covered
    static const int forced = [] { const char* e = getenv("C2DSR_DP_UNROLL"); return e ? atoi(e) : 0; }();
    const int unroll = forced ? forced : (map->world <= 2 ? 1 : 4);
    const bool mc = map->grad_mc != nullptr && map->param_mc != nullptr;
    auto launch = [&](auto kernel) {
        kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table_dev, *map, beta1, beta2, eps, weight_decay, st);
    };
    if (mc && unroll == 4) launch(adamw_peer_kernel<true, 4>);
    else if (mc) launch(adamw_peer_kernel<true, 1>);
    else if (unroll == 4) launch(adamw_peer_kernel<false, 4>);
    else launch(adamw_peer_kernel<false, 1>);
    note_launches(1);
    return check_launch("adamw_amsgrad_peer");
}

}  // extern "C"
