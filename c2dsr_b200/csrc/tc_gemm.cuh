// tcgen05 / TMEM / TMA building blocks for the score GEMMs (sm_100a).
//
// C[M, N] = A[M, K] * B[N, K]^T with both operands K-major bf16, fp32 accumulation in tensor memory.
// fp32 fidelity comes from a split: x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the kernel issues
// lo*hi + hi*lo + hi*hi per 64-wide k-block (the lo*lo term is below 2^-16 relative), so one staged
// k-block of {A_hi, A_lo, B_hi, B_lo} feeds three MMA groups ("passes" = 3).  passes = 1 uses hi only.
//
// Kernel anatomy (one CTA per SM, persistent over a contiguous range of output tiles):
//   warp 0      TMA producer   cp.async.bulk.tensor.2d -> 128B-swizzled smem, mbarrier expect_tx
//   warp 1      MMA issuer     one lane issues tcgen05.mma (M = 128, N = BN, K = 16), commits to mbarriers
//   warp 2      TMEM allocator 2 accumulator stages x BN columns
//   warps 4..11 epilogue       tcgen05.ld 32 lanes x 32 columns at a time; the functor consumes rows
//                              (two warps per TMEM lane quarter, each owns half of the tile's columns)
// so the epilogue of tile i overlaps the MMAs of tile i+1 and the TMA loads of tile i+2.
//
// Two operand-staging modes:
//   ARES = false  every k-block stage holds A and B tiles (any K; used for the long-K gradient GEMMs,
//                 optionally split along K into slabs whose partial sums the caller adds in order)
//   ARES = true   the whole A row block (all k-blocks, hi and lo) stays resident in TENSOR MEMORY (256 columns:
//                 bf16 pairs packed along K, one row per lane; tcgen05.mma reads A from TMEM) while the CTA
//                 walks consecutive N tiles.  Only B is streamed, and ALL of shared memory is ring: 7 stages of
//                 32 KB instead of 3, i.e. ~190 KB of B tiles in flight -- what it takes to cover the TMA latency
//                 at 42 B/clk/SM (with 3 stages the MMA warp waited on `full` half the time).  The epilogue
//                 warps load a new row block (global -> registers -> tcgen05.st) when it changes.  Needs
//                 K <= 256 and BN = 128 (the logits / score GEMMs with d <= 256).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace c2dsr {
namespace tc {

constexpr int BM = 128;        // rows of A per tile = TMEM lanes
constexpr int BK = 64;         // bf16 elements per k-block = 128 bytes = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int ARES_MAX_KB = 4; // resident A: up to 4 k-blocks (K <= 256)
constexpr int ARES_STAGES = 7; // ring depth of the resident-A kernels (BN = 128: 7 x 32 KB)
constexpr int EPI_WARPS = 8;   // epilogue warps: two per TMEM lane quarter, each takes half of the tile's columns
constexpr int EPI_PARTS = EPI_WARPS / 4;
constexpr int THREADS = 128 + 32 * EPI_WARPS;

// ex2.approx: 2 ulp, maps -inf to +0 (what the online soft-max needs)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// start fetching the line at p into L1 (the epilogues ask for the per-column parameters of a tile before they wait for
// its accumulator, so the L2 round trip overlaps the MMAs instead of stalling the first add of every chunk)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates are (inner = k element, outer = row)
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// elect.sync-predicated forms for a producer warp that runs its loop with all lanes (see umma_bf16_elect: a
// `lane == 0` branch anywhere in the role code makes ptxas treat the MMA issuer's counters as per-thread values)
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(bytes)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t"
        "}" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---- tensor memory ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The MMA-issuing warp runs its loops with ALL lanes (warp-uniform control flow) and predicates only the instruction
// itself on elect.sync: every operand then comes from provably uniform values (shared-memory base, kernel
// parameters, block index, a __shfl_sync broadcast of the TMEM base) and stays in uniform registers.  Issuing from a
// single lane (`if (lane == 0)`) made the compiler wrap each UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop
// (~20 SASS instructions, > 100 cycles per MMA): the N = 128 kernels were issue-bound at ~55 % tensor-pipe activity
// (scratch/mma_issue_bench.cu: 64.0 cycles per 128x128x16 MMA = nominal with this form).
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The shared-memory descriptor as two 32-bit halves: only the low word depends on the address (start address >> 4 in
// bits 0-13, leading byte offset >> 4 in bits 16-29); the high word (stride byte offset 1024 >> 4, descriptor
// version 1, SWIZZLE_128B) is the same constant for every operand tile of these kernels.
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, bool mn_major) {
    return ((smem_addr & 0x3FFFFu) >> 4) | ((uint32_t)(mn_major ? (BK * 128) >> 4 : 1) << 16);
}
__device__ __forceinline__ uint32_t uniform32(uint32_t x) { return __shfl_sync(0xffffffffu, x, 0); }
__device__ __forceinline__ void umma_lo_elect(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b32 hi;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b32 hi, %5;\n\t"
        "mov.b64 da, {%1, hi};\n\t"
        "mov.b64 db, {%2, hi};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "n"(kDescHi)
        : "memory");
}
__device__ __forceinline__ void umma_ts_lo_elect(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo32, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b32 hi;\n\t"
        ".reg .b64 db;\n\t"
        "mov.b32 hi, %5;\n\t"
        "mov.b64 db, {%2, hi};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo32), "r"(idesc), "r"(accumulate), "n"(kDescHi)
        : "memory");
}
// A from tensor memory (bf16 pairs packed along K, one row per lane), B from shared memory
__device__ __forceinline__ void umma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                   uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
// mbarrier arrives once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base lane + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
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

// 128-byte swizzled operand tiles, both built from 128-byte smem rows (tile base 1024-B aligned):
//   K-major   row = one M/N index, 64 consecutive K elements inside the row (TMA box 64 k x rows).
//             8 rows form a swizzle atom: SBO = 1024 B; LBO unused; the next UMMA_K = 16 elements are +32 B.
//   MN-major  row = one K index, 64 consecutive M/N elements inside the row (TMA box 64 mn x 64 k = 8 KB);
//             8 K rows form an atom: SBO = 1024 B between 8-row K groups, LBO = 8192 B between 64-wide M/N
//             groups (one TMA box each); the next UMMA_K = 16 K rows are +2048 B.
// Descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
constexpr int MN_GROUP_BYTES = BK * 128;     // one [64 k][64 mn] box
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, bool mn_major) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(mn_major ? (MN_GROUP_BYTES >> 4) : 1) << 16;     // leading byte offset
    d |= (uint64_t)(1024 >> 4) << 32;                               // stride byte offset
    d |= (uint64_t)1 << 46;                                         // descriptor version
    d |= (uint64_t)2 << 61;                                         // SWIZZLE_128B
    return d;
}
// descriptor increment (in 16-byte units) for the next UMMA_K slice of a k-block
__host__ __device__ constexpr uint32_t kstep_units(bool mn_major) { return mn_major ? (2 * 1024) >> 4 : (UMMA_K * 2) >> 4; }
// kind::f16 instruction descriptor: D fp32, A/B bf16, M = 128, N = n; bit 15 / 16 = A / B is MN-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Maps {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

struct Problem {
    int64_t M, N;       // rows of A, rows of B
    int K;              // shared inner extent (elements)
    int passes;         // 3 = hi/lo split, 1 = hi only
    int diag_only;      // 1: only tiles with m_blk == n_blk (target-score pass, BN == BM)
    int k_splits;       // >= 1: K is cut into this many slabs; the epilogue receives the slab index
    const uint16_t* a_hi_ptr;   // resident-A mode: the bf16 hi / lo matrices of A themselves ([M, lda], K-major)
    const uint16_t* a_lo_ptr;
    int64_t lda;
    const int* m_limit; // optional device-resident row count: only the first min(M, *m_limit) rows of A are computed
                        // (M stays the capacity the tensor maps and the launch were sized for; lets a captured
                        // launch follow a row count that is only known on the device)
};

template <int BN, int STAGES, bool ARES>
struct SmemLayout {
    static constexpr int A_TILE = BM * BK * 2;           // 16 KB
    static constexpr int B_TILE = BN * BK * 2;
    static constexpr int A_RES = 0;                       // (the resident A row block lives in tensor memory)
    static constexpr int STAGE = (ARES ? 0 : 2 * A_TILE) + 2 * B_TILE;
    static constexpr int B_OFF = ARES ? 0 : 2 * A_TILE;   // offset of B_hi inside a stage
    static constexpr int RING_OFF = A_RES;
    static constexpr int BARRIER_OFF = RING_OFF + STAGES * STAGE;
    static constexpr int TOTAL = BARRIER_OFF + 256 + 1024;   // barriers + tmem pointer + alignment slack
};

// Compile-time dispatch of a small runtime index: f(std::integral_constant<int, v>).  The MMA issuer uses it to turn
// the ring stage, the accumulator stage and (resident A) the k-block into constants, so that every descriptor is
// "shared-memory base + constant": values that live in uniform registers from birth, with no per-MMA R2UR traffic.
template <int N, class F>
__device__ __forceinline__ void dispatch_index(int v, F&& f) {
    if constexpr (N <= 1) {
        f(std::integral_constant<int, 0>{});
    } else {
        if (v == N - 1) f(std::integral_constant<int, N - 1>{});
        else dispatch_index<N - 1>(v, f);
    }
}

// Epilogue functor contract (one instance per epilogue thread; the thread owns A row `row`):
//   void tile_begin(int64_t m_blk, int64_t n_blk, int64_t row, int k_slab, int part);
//   void chunk(int64_t row, int64_t col0, const float (&v)[32]);   // columns col0..col0+31 of the tile
//   void tile_end(int64_t row);
// Each row of a tile is shared by EPI_PARTS threads (`part` = which BN / EPI_PARTS column range they own).
// A_MN / B_MN: the operand is MN-major (its global matrix is [K, M or N] row-major) instead of K-major.
template <int BN, int STAGES, bool ARES, bool A_MN, bool B_MN, class Epilogue>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ Maps maps, const Problem pb, Epilogue epi) {
    static_assert(!(ARES && A_MN), "the resident-A mode is K-major only");
    using L = SmemLayout<BN, STAGES, ARES>;
    // operand tile loaders: K-major = one box of [rows][64 k]; MN-major = one [64 k][64 mn] box per 64 rows
    auto load_a = [&](const CUtensorMap* map, uint64_t* bar, uint8_t* dst, int kb, int64_t m_blk) {
        if (A_MN) {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g)
                tma_load_2d_elect(map, bar, dst + g * MN_GROUP_BYTES, (int)(m_blk * BM) + 64 * g, kb * BK);
        } else {
            tma_load_2d_elect(map, bar, dst, kb * BK, (int)(m_blk * BM));
        }
    };
    auto load_b = [&](const CUtensorMap* map, uint64_t* bar, uint8_t* dst, int kb, int64_t n_blk) {
        if (B_MN) {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)
                tma_load_2d_elect(map, bar, dst + g * MN_GROUP_BYTES, (int)(n_blk * BN) + 64 * g, kb * BK);
        } else {
            tma_load_2d_elect(map, bar, dst, kb * BK, (int)(n_blk * BN));
        }
    };
    extern __shared__ __align__(1024) uint8_t smem[];       // (1024-byte aligned: what the 128B swizzle atoms need)
    uint8_t* ring = smem + L::RING_OFF;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BARRIER_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* a_full = tmem_empty + 2;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_empty + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    int64_t M_eff = pb.M;
    if (pb.m_limit) {
        const int64_t lim = *pb.m_limit;
        M_eff = lim < M_eff ? (lim > 0 ? lim : 0) : M_eff;
    }
    const int64_t m_blocks = (M_eff + BM - 1) / BM, n_blocks = (pb.N + BN - 1) / BN;
    const int ks = pb.k_splits > 1 ? pb.k_splits : 1;
    const int64_t n_tiles = pb.diag_only ? m_blocks : m_blocks * n_blocks * ks;
    const int n_kb_total = (pb.K + BK - 1) / BK;
    const int kb_per_slab = (n_kb_total + ks - 1) / ks;
    const bool split = pb.passes == 3;
    const uint32_t stage_bytes = (ARES ? 0u : (uint32_t)L::A_TILE * (split ? 2u : 1u)) + (uint32_t)L::B_TILE * (split ? 2u : 1u);
    // contiguous tile range of this CTA (consecutive tiles share the A row block)
    int64_t t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
    if (ARES && !pb.diag_only && ks == 1 && m_blocks * 2 <= (int64_t)gridDim.x) {
        // resident A with few row blocks (the ranking GEMM: ~1 000 queries against up to 10^6 items): give every
        // row block the same number of CTAs and cut the columns identically for all of them, so that the CTAs of
        // different row blocks walk the SAME column tiles in lockstep and a B tile streamed for one row block is an
        // L2 hit for the others (an even split of the flat tile list staggers them: at the 1M-item catalogue
        // the 614 MB classifier was read 9 times from HBM, 5.5 GB per launch)
        const int64_t G = (int64_t)gridDim.x / m_blocks;
        const int64_t m = blockIdx.x / G, g = blockIdx.x % G;
        if (m >= m_blocks) {
            t0 = t1 = 0;
        } else {
            const int64_t per = (n_blocks + G - 1) / G;
            const int64_t n0 = g * per < n_blocks ? g * per : n_blocks;
            const int64_t n1 = n0 + per < n_blocks ? n0 + per : n_blocks;
            t0 = m * n_blocks + n0;
            t1 = m * n_blocks + n1;
        }
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a_hi);
        tma_prefetch_desc(&maps.b_hi);
        if (split) {
            tma_prefetch_desc(&maps.a_lo);
            tma_prefetch_desc(&maps.b_lo);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], EPI_WARPS);
        }
        mbar_init(a_full, EPI_WARPS);      // (resident A is written by the epilogue warps)
        mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = ARES ? 512 : 2 * BN;   // resident A: hi in columns [256, 384), lo in [384, 512)
    constexpr uint32_t A_TMEM = 256;
    static_assert(!ARES || BN == 128, "the resident-A mode is laid out for BN = 128");
    if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    // tile t -> (m_blk, n_blk, k slab); slabs of one output tile are consecutive
    auto tile_coords = [&](int64_t t, int64_t& m_blk, int64_t& n_blk, int& slab) {
        if (pb.diag_only) {
            m_blk = n_blk = t;
            slab = 0;
        } else {
            slab = (int)(t % ks);
            const int64_t mn = t / ks;
            m_blk = mn / n_blocks;
            n_blk = mn % n_blocks;
        }
    };
    auto kb_range = [&](int slab, int& kb0, int& kb1) {
        kb0 = slab * kb_per_slab;
        kb1 = kb0 + kb_per_slab < n_kb_total ? kb0 + kb_per_slab : n_kb_total;
    };

    // (role branches are on the warp index only -- provably warp-uniform -- so that ptxas keeps the issuer's loop
    //  counters and descriptors in uniform registers; the lane test of the producer is nested inside its branch)
    if (warp == 0) {
        // ---------------- TMA producer: all lanes run the loop, elect.sync issues ----------------
        {
        int stage = 0;
        uint32_t phase = 0, a_phase = 0;
        int64_t cur_m = -1;
        for (int64_t t = t0; t < t1; ++t) {
            int64_t m_blk, n_blk;
            int slab, kb0, kb1;
            tile_coords(t, m_blk, n_blk, slab);
            kb_range(slab, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = ring + stage * L::STAGE;
                mbar_arrive_expect_tx_elect(&full[stage], stage_bytes);
                if (!ARES) {
                    load_a(&maps.a_hi, &full[stage], st, kb, m_blk);
                    if (split) load_a(&maps.a_lo, &full[stage], st + L::A_TILE, kb, m_blk);
                }
                load_b(&maps.b_hi, &full[stage], st + L::B_OFF, kb, n_blk);
                if (split) load_b(&maps.b_lo, &full[stage], st + L::B_OFF + L::B_TILE, kb, n_blk);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: the whole warp runs the loop, elect.sync picks the issuing lane ----------------
        const uint32_t smem_base = smem_u32(smem);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0, a_phase = 0;
        int64_t cur_m = -1;
        for (int64_t t = t0; t < t1; ++t) {
            int64_t m_blk, n_blk;
            int slab, kb0, kb1;
            tile_coords(t, m_blk, n_blk, slab);
            kb_range(slab, kb0, kb1);
            if (ARES && m_blk != cur_m) {
                if (cur_m >= 0) umma_commit_elect(a_empty);          // previous row block no longer needed
                mbar_wait(a_full, a_phase);                          // the epilogue warps have stored the new one
                tcgen05_fence_after();
                a_phase ^= 1;
                cur_m = m_blk;
            }
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tcgen05_fence_after();
                // descriptor low words from the (warp-uniform) stage / k-block counters: uniform-datapath arithmetic
                const uint32_t st = smem_base + (uint32_t)(L::RING_OFF + stage * L::STAGE);
                const uint32_t a_hi = ARES ? tmem_base + A_TMEM + (uint32_t)(kb * (BK / 2)) : desc_lo(st, A_MN);
                const uint32_t a_lo = ARES ? a_hi + 128u : a_hi + (uint32_t)(L::A_TILE >> 4);
                const uint32_t b_hi = desc_lo(st + L::B_OFF, B_MN), b_lo = b_hi + (uint32_t)(L::B_TILE >> 4);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                constexpr uint32_t idesc = make_idesc_bf16(BN, A_MN, B_MN);
                // per UMMA_K slice: A advances 8 TMEM columns (16 bf16) or one descriptor k-step
                constexpr uint32_t ka = ARES ? (uint32_t)(UMMA_K / 2) : kstep_units(A_MN), kbs = kstep_units(B_MN);
                auto mma = [&](uint32_t a, uint32_t b, uint32_t accumulate) {
                    if (ARES) umma_ts_lo_elect(d_tmem, a, b, idesc, accumulate);
                    else umma_lo_elect(d_tmem, a, b, idesc, accumulate);
                };
                uint32_t accum = kb > kb0 ? 1u : 0u;
                if (split) {      // small cross terms first, then the leading product
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        mma(a_lo + ka * k, b_hi + kbs * k, accum);
                        accum = 1u;
                    }
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) mma(a_hi + ka * k, b_lo + kbs * k, 1u);
                }
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    mma(a_hi + ka * k, b_hi + kbs * k, accum);
                    accum = 1u;
                }
                umma_commit_elect(&empty[stage]);          // smem stage is free once these MMAs retire
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit_elect(&tmem_full[acc]);            // accumulator complete -> epilogue
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue ----------------
        const int q = warp & 3;                      // TMEM lane quarter this warp may access (warp id mod 4)
        const int part = (warp - 4) >> 2;            // which column range of the tile
        constexpr int PART_COLS = BN / EPI_PARTS;
        int acc = 0;
        uint32_t acc_phase = 0, a_phase_epi = 0;
        int64_t cur_m_epi = -1;
        for (int64_t t = t0; t < t1; ++t) {
            int64_t m_blk, n_blk;
            int slab;
            tile_coords(t, m_blk, n_blk, slab);
            const int64_t row = m_blk * BM + q * 32 + lane;
            if (ARES && m_blk != cur_m_epi) {
                // resident A: this thread's row of the new block, bf16 pairs as stored (K-major), into its TMEM
                // lane; the `part` 0 warps store the hi half (columns [256, 384)), the `part` 1 warps the lo half
                if (cur_m_epi >= 0) {
                    mbar_wait(a_empty, a_phase_epi);             // every MMA that read the previous block is done
                    a_phase_epi ^= 1;
                    tcgen05_fence_after();
                }
                cur_m_epi = m_blk;
                const uint16_t* src = part == 0 ? pb.a_hi_ptr : pb.a_lo_ptr;
                const bool live = row < pb.M && (part == 0 || split);
                const uint4* r4 = reinterpret_cast<const uint4*>(src + row * pb.lda);
                const int n_vec = pb.K >> 3;                     // 16-byte vectors in the row (K % 8 == 0)
                const uint32_t a_addr = tmem_base + ((uint32_t)(q * 32) << 16) + A_TMEM + (uint32_t)(part * 128);
#pragma unroll 1
                for (int c = 0; c < ARES_MAX_KB; ++c) {          // 32 columns = 64 bf16 = one k-block per store
                    uint32_t w[32];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const int vi = c * 8 + v;
                        uint4 x = make_uint4(0u, 0u, 0u, 0u);
                        if (live && vi < n_vec) x = __ldg(r4 + vi);
                        w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
                    }
                    tmem_st32(a_addr + (uint32_t)(32 * c), w);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full);
            }
            epi.tile_begin(m_blk, n_blk, row, slab, part);
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + part * PART_COLS);
#pragma unroll 1
            for (int c = 0; c < PART_COLS / 32; ++c) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), v);
                epi.chunk(row, n_blk * BN + part * PART_COLS + c * 32, v);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            epi.tile_end(row);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
}  // namespace c2dsr
