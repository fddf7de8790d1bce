// Micro-benchmark + numerics check of the tcgen05.mma forms the K4a kernels use (scratch tooling, not product):
//   SS  A and B from 128B-swizzled shared memory (K-major A; K- or MN-major B)
//   TS  A from tensor memory (bf16 pairs packed along K, one row per lane), B from shared memory
// for N in {64, 128, 256}.  Every CTA (one per SM) issues `iters` rounds of one 64-wide k-block (4 MMAs of K = 16)
// back to back from a single thread and reports cycles per MMA; CTA 0 also writes D so the host can compare it with
// an exact integer reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/_bin/mma_bench
//        scratch/mma_bench.cu -I c2dsr_b200/csrc -lcuda
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "tc_gemm.cuh"

using namespace c2dsr::tc;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__host__ __device__ inline int a_val(int m, int k) { return (m * 3 + k * 5) % 7 - 3; }
__host__ __device__ inline int b_val(int n, int k) { return (n * 2 + k * 3) % 5 - 2; }

// byte offset of element (row, col) in a [rows][64] bf16 tile of 128-byte rows with the 128B swizzle (8-row atoms)
__device__ __forceinline__ uint32_t sw128(int row, int col) {
    const int chunk = (col >> 3) ^ (row & 7);
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + chunk * 16 + (col & 7) * 2);
}

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(pred));
    return pred;
}

struct Params {
    int mode;      // 0 = SS, 1 = TS
    int n;         // 64, 128, 256
    int b_mn;      // B is MN-major
    int iters;
    int extra_smem_traffic;    // 1: the other warps keep reading shared memory (like TMA writes / epilogues would)
    int chains;                // independent accumulators the MMAs rotate over (1 = every MMA depends on the previous one)
    int uniform;               // 1: warp-uniform issue loop + elect.sync; 0: a single lane runs the loop
};

__global__ void __launch_bounds__(192, 1) mma_bench_kernel(Params p, float* D_out, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                      // [128][64] bf16, K-major, 16 KB
    uint8_t* sB = smem + 16384;              // K-major: [n][64] (n * 128 B); MN-major: n / 64 groups of [64 k][64 n] (8 KB each)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = p.n;
    __nv_bfloat16* a16 = nullptr;
    (void)a16;
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int m = i >> 6, k = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(sA + sw128(m, k)) = __float2bfloat16((float)a_val(m, k));
    }
    for (int i = threadIdx.x; i < N * 64; i += blockDim.x) {
        const int n = i >> 6, k = i & 63;
        uint32_t off;
        if (p.b_mn) off = (uint32_t)((n >> 6) * 8192) + sw128(k, n & 63);     // row = k, 64 consecutive n inside the row
        else off = sw128(n, k);
        *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16((float)b_val(n, k));
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    // generic-proxy writes of the operand tiles must be visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_a = tmem_base + 480;        // TS: A lives in columns [480, 512)
    if (p.mode == 1 && warp < 4) {
        uint32_t r[32];
        const int m = warp * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const __nv_bfloat162 v = __floats2bfloat162_rn((float)a_val(m, 2 * j), (float)a_val(m, 2 * j + 1));
            r[j] = *reinterpret_cast<const uint32_t*>(&v);
        }
        tmem_st32(tmem_a + ((uint32_t)(warp * 32) << 16), r);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    long long t0 = 0, t1 = 0;
    if (warp == 4 && (p.uniform || lane == 0)) {
        // uniform = 1: the WHOLE warp runs the issue loop with identical values and only the instruction itself is
        // predicated on elect.sync, so descriptors stay in uniform registers (no per-MMA R2UR broadcast loop)
        const uint32_t idesc = make_idesc_bf16(N, false, p.b_mn != 0);
        const uint32_t kb = kstep_units(p.b_mn != 0);
        const uint64_t adesc = make_smem_desc(smem_u32(sA), false), bdesc = make_smem_desc(smem_u32(sB), p.b_mn != 0);
        t0 = clock64();
        for (int it = 0; it < p.iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = it * 4 + k;
                const uint32_t acc = idx >= p.chains ? 1u : 0u;
                const uint32_t dcol = tmem_base + (uint32_t)((idx & (p.chains - 1)) * N);
                if (!p.uniform || elect_one()) {
                    if (p.mode == 0) umma_bf16(dcol, adesc + 2 * k, bdesc + kb * k, idesc, acc);
                    else umma_bf16_ts(dcol, tmem_a + 8 * k, bdesc + kb * k, idesc, acc);
                }
            }
        }
        if (!p.uniform || elect_one()) umma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    } else if (warp == 5 && p.extra_smem_traffic) {
        // a steady stream of shared-memory reads from another warp while the MMAs run
        volatile uint4* q = reinterpret_cast<volatile uint4*>(sB + 16384);
        uint32_t acc = 0;
        for (int it = 0; it < p.iters * 8; ++it) acc += q[(it * 32 + lane) & 511].x;
        if (acc == 0x12345678u) cycles[0] = 0;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (blockIdx.x == 0 && warp < 4) {
        // D accumulated iters times the same product: out = D / iters
        for (int c = 0; c < N; c += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            for (int i = 0; i < 32; ++i) D_out[(warp * 32 + lane) * 256 + c + i] = v[i];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem_bytes = 16384 + 32768 + 1024 + 256;
    cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    float* dD;
    long long* dC;
    cudaMalloc(&dD, 128 * 256 * 4);
    cudaMalloc(&dC, sms * 8);
    std::vector<float> D(128 * 256);
    std::vector<long long> C(sms);
    const int iters = 2000;
    printf("%-4s %-4s %-5s %-6s %10s %10s %8s %s\n", "mode", "N", "B", "chains", "clk/MMA", "nominal", "ratio", "numerics");
    for (int mode = 0; mode < 2; ++mode)
        for (int bmn = 0; bmn < 2; ++bmn)
            for (int n : {64, 128, 256})
                for (int tr : {1, 2, 4, 10, 20}) {
                    const int chains = tr >= 10 ? tr / 10 : tr, uni = tr >= 10 ? 1 : 0;
                    if (chains * n > 448) continue;
                    Params p{mode, n, bmn, iters, 0, chains, uni};
                    for (int grid : {sms}) {
                        cudaMemset(dD, 0, 128 * 256 * 4);
                        mma_bench_kernel<<<grid, 192, smem_bytes>>>(p, dD, dC);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) {
                            printf("mode %d n %d bmn %d: %s\n", mode, n, bmn, cudaGetErrorString(e));
                            return 1;
                        }
                        cudaMemcpy(D.data(), dD, 128 * 256 * 4, cudaMemcpyDeviceToHost);
                        cudaMemcpy(C.data(), dC, grid * 8, cudaMemcpyDeviceToHost);
                        double worst = 0;
                        if (chains == 1)
                        for (int m = 0; m < 128; ++m)
                            for (int j = 0; j < n; ++j) {
                                long ref = 0;
                                for (int k = 0; k < 64; ++k) ref += a_val(m, k) * b_val(j, k);
                                const double d = fabs((double)D[m * 256 + j] / iters - (double)ref);
                                if (d > worst) worst = d;
                            }
                        long long mx = 0;
                        for (int i = 0; i < grid; ++i) mx = C[i] > mx ? C[i] : mx;
                        const double per = (double)mx / (iters * 4.0), nominal = 128.0 * n / 256.0;
                        printf("%-4s %-4d %-5s %-6d %10.1f %10.1f %8.2f max|err| %.3g  (grid %d)\n", mode ? "TS" : "SS", n,
                               bmn ? "MN" : "K", tr, per, nominal, per / nominal, worst, grid);
                    }
                }
    return 0;
}
