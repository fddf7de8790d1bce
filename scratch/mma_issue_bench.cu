// How fast can ONE warp issue tcgen05.mma?  (scratch tooling)  The product kernels' first version ran the issue loop
// in a single lane (`if (lane == 0)`): every descriptor then lives in per-thread registers and each UTCHMMA is wrapped
// in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop -- ~20 SASS instructions and > 100 cycles per MMA, i.e. the tensor
// pipe idles half the time at N = 128.  Here: warp-uniform control flow, operands from provably uniform sources
// (shared-memory base, __shfl_sync broadcast of the TMEM base), elect.sync inside the asm.
// VARIANT 0 = single lane (old), 1 = warp-uniform.  Reports cycles per MMA for N = 64 / 128 / 256, SS and TS, with
// 1 or 2 independent accumulators.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "tc_gemm.cuh"
using namespace c2dsr::tc;

__device__ __forceinline__ void umma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

template <int VARIANT, int N, int TS, int CHAINS>
__global__ void __launch_bounds__(192, 1) k(int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    uint32_t tmem_base = *tmem_ptr;
    if (VARIANT >= 1) tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (warp == 4) {
        constexpr uint32_t idesc = make_idesc_bf16(N, false, false);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384);
        const uint64_t adesc = make_smem_desc(sa, false), bdesc = make_smem_desc(sb, false);
        const uint32_t ta = tmem_base + 480;
        long long t0 = clock64();
        if (VARIANT == 0) {
            if (lane == 0) {
                for (int it = 0; it < iters; ++it) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t d = tmem_base + ((kk % CHAINS) * N);
                        umma_bf16(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, it ? 1u : 0u);
                    }
                }
                umma_commit(&bar[0]);
            }
        } else {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t d = tmem_base + ((kk % CHAINS) * N);
                    if (TS) umma_ts_elect(d, ta + 8 * kk, bdesc + 2 * kk, idesc, it ? 1u : 0u);
                    else umma_elect(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, it ? 1u : 0u);
                }
            }
            commit_elect(&bar[0]);
        }
        mbar_wait(&bar[0], 0);
        if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int VARIANT, int N, int TS, int CHAINS>
void run(int sms, long long* dC) {
    const int iters = 4000, smem_bytes = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(k<VARIANT, N, TS, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    std::vector<long long> C(sms);
    k<VARIANT, N, TS, CHAINS><<<sms, 192, smem_bytes>>>(iters, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(C.data(), dC, sms * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = C[i] > mx ? C[i] : mx;
    const double per = (double)mx / (iters * 4.0), nominal = 128.0 * N / 256.0;
    printf("%-12s %-3s N=%-4d chains=%d  %8.1f clk/MMA  nominal %6.1f  ratio %5.2f  -> %5.1f %% of the 4096 MAC/clk/SM peak\n",
           VARIANT ? "warp-uniform" : "single-lane", TS ? "TS" : "SS", N, CHAINS, per, nominal, per / nominal, 100.0 * nominal / per);
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long* dC;
    cudaMalloc(&dC, sms * 8);
    run<0, 128, 0, 1>(sms, dC);
    run<1, 64, 0, 1>(sms, dC);  run<1, 64, 0, 2>(sms, dC);  run<1, 64, 0, 4>(sms, dC);
    run<1, 128, 0, 1>(sms, dC); run<1, 128, 0, 2>(sms, dC);
    run<1, 256, 0, 1>(sms, dC);
    run<1, 64, 1, 1>(sms, dC);  run<1, 64, 1, 2>(sms, dC);  run<1, 64, 1, 4>(sms, dC);
    run<1, 128, 1, 1>(sms, dC); run<1, 128, 1, 2>(sms, dC);
    run<1, 256, 1, 1>(sms, dC);
    return 0;
}
