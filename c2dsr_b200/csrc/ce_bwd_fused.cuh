// K4a backward, fused: dZ = (softmax(H W^T + b) - onehot) * coef never leaves the SM.
//
// Two launches of ONE kernel template share the work of trainer.py:131-156's backward through the logits:
//   TRANSPOSED = false   X = H (rows m, stationary), Y = W (rows n, streamed):  OUT = dH[m, :] = sum_n dZ[m, n] W[n, :]
//   TRANSPOSED = true    X = W (rows n, stationary), Y = H (rows m, streamed):  OUT = dW[n, :] = sum_m dZ[m, n] H[m, :]
//                        (+ db[n] = sum_m dZ[m, n], a running row sum of the epilogue threads)
// For every 128 x 128 tile (x block, y tile) of the logits:
//   S  = X Y^T                      tcgen05.mma SS, K = d <= 256, bf16 hi/lo split (3 MMAs per product), fp32 in TMEM
//   P  = dZ tile                    epilogue warps: tcgen05.ld -> exp2 / scale / minus one-hot -> bf16 hi, lo -> tcgen05.st
//                                   IN PLACE over the S columns they were read from (tensor memory, not shared memory)
//   OUT += P Y                      tcgen05.mma with A = P from TENSOR MEMORY and B = the same Y k-block tile, read
//                                   MN-major from the very layout the first product read K-major (no transposed copy)
// OUT (128 x d fp32 = 256 TMEM columns) stays resident while the CTA walks the y tiles of its x block; it is flushed
// once per (x block, CTA) as a partial that a small kernel adds in a fixed order (deterministic).  TMEM: two S / P
// stages of 128 columns + OUT 256 columns = 512.  Shared memory: X hi/lo resident (128 KB) + a 3-stage ring of Y
// k-block tiles (hi + lo, 32 KB each); every Y k-block passes through the ring twice (once per product).
// The MMA warp interleaves "S of tile j + 1" with "OUT of tile j", so the epilogue of a tile overlaps tensor work.
#pragma once
#include "tc_gemm.cuh"

namespace c2dsr {
namespace tc {

constexpr int FB_STAGES = 3;
constexpr int FB_X_BYTES = ARES_MAX_KB * 2 * BM * BK * 2;      // 128 KB: X hi / lo, 4 k-blocks
constexpr int FB_STAGE_BYTES = 2 * BM * BK * 2;                // 32 KB: one Y k-block, hi + lo
constexpr int FB_BAR_OFF = FB_X_BYTES + FB_STAGES * FB_STAGE_BYTES;
constexpr int FB_SMEM = FB_BAR_OFF + 256 + 1024;
constexpr int FB_OUT_COL = 256;                                // TMEM column of OUT

struct CeBwdProblem {
    int64_t X_rows, Y_rows;       // rows of the stationary / streamed operand
    int d;                        // K of the first product = N of the second (<= 256)
    int passes;                   // 3 = hi / lo split, 1 = hi only
    // per-row / per-column soft-max parameters, padded to multiples of 128 with values that make dZ = 0:
    const float* l2s;             // [M_pad]  lse * log2(e)  (1e30 for rows that carry no target or lie past M)
    const float* cfs;             // [M_pad]  d loss / d loss_row (0 for such rows)
    const int* g32;               // [M_pad]  target column (-1 for such rows)
    const float* bl;              // [N_pad]  bias * log2(e)  (-1e30 past N)
    float* out_part;              // [x_blocks][max_slots][128][d]
    float* db_part;               // TRANSPOSED: [x_blocks][max_slots][2][128] row sums of dZ^T (per column half)
    int max_slots;
};

// two fp32 -> one register of two bf16 (low half = first value), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// CTA c owns tiles [n_tiles * c / grid, n_tiles * (c + 1) / grid); the CTA that owns tile T:
__host__ __device__ inline int64_t fb_cta_of_tile(int64_t T, int64_t n_tiles, int64_t grid) {
    return ((T + 1) * grid + n_tiles - 1) / n_tiles - 1;
}

template <bool TRANSPOSED>
__global__ void __launch_bounds__(THREADS, 1)
ce_bwd_kernel(const __grid_constant__ Maps maps, const CeBwdProblem pb) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + FB_BAR_OFF);
    uint64_t* empty = full + FB_STAGES;
    uint64_t* s_full = empty + FB_STAGES;      // [2] S of a tile complete -> epilogue
    uint64_t* p_full = s_full + 2;             // [2] P of a tile written   -> MMA
    uint64_t* x_full = p_full + 2;
    uint64_t* x_empty = x_full + 1;
    uint64_t* out_full = x_empty + 1;
    uint64_t* out_empty = out_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(out_empty + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int64_t x_blocks = (pb.X_rows + BM - 1) / BM, y_tiles = (pb.Y_rows + BM - 1) / BM;
    const int64_t n_tiles = x_blocks * y_tiles;
    const int n_kb = (pb.d + BK - 1) / BK;
    const bool split = pb.passes == 3;
    const int64_t t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
    const int n_my = (int)(t1 - t0);
    const uint32_t stage_tx = (uint32_t)(BM * BK * 2) * (split ? 2u : 1u);

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < FB_STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&s_full[s], 1);
                mbar_init(&p_full[s], EPI_WARPS);
            }
            mbar_init(x_full, 1);
            mbar_init(x_empty, 1);
            mbar_init(out_full, 1);
            mbar_init(out_empty, EPI_WARPS);
            fence_barrier_init();
        }
    } else if (warp == 2) {
        tmem_alloc(tmem_ptr, 512);
    } else if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.a_hi);
            tma_prefetch_desc(&maps.b_hi);
            if (split) {
                tma_prefetch_desc(&maps.a_lo);
                tma_prefetch_desc(&maps.b_lo);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    uint8_t* ring = smem + FB_X_BYTES;

    // Role branches test the warp index only, and every role runs its loops with all lanes (elect.sync picks the
    // issuing lane): see tc_gemm.cuh for why that matters for the MMA issue rate.
    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0;
        uint32_t phase = 0, x_phase = 0;
        int64_t cur_x = -1;
        auto stream_y = [&](int64_t y_tile) {           // the 4 k-blocks of one Y tile through the ring
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = ring + stage * FB_STAGE_BYTES;
                mbar_arrive_expect_tx_elect(&full[stage], stage_tx);
                tma_load_2d_elect(&maps.b_hi, &full[stage], st, kb * BK, (int)(y_tile * BM));
                if (split) tma_load_2d_elect(&maps.b_lo, &full[stage], st + BM * BK * 2, kb * BK, (int)(y_tile * BM));
                if (++stage == FB_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        };
        for (int j = 0; j <= n_my; ++j) {
            if (j < n_my) {
                const int64_t t = t0 + j, x_blk = t / y_tiles, y_tile = t % y_tiles;
                if (x_blk != cur_x) {
                    mbar_wait(x_empty, x_phase ^ 1);          // first products of the previous x block are done
                    mbar_arrive_expect_tx_elect(x_full, (uint32_t)n_kb * (uint32_t)(BM * BK * 2) * (split ? 2u : 1u));
                    for (int kb = 0; kb < n_kb; ++kb) {
                        tma_load_2d_elect(&maps.a_hi, x_full, smem + kb * 2 * BM * BK * 2, kb * BK, (int)(x_blk * BM));
                        if (split)
                            tma_load_2d_elect(&maps.a_lo, x_full, smem + kb * 2 * BM * BK * 2 + BM * BK * 2, kb * BK,
                                              (int)(x_blk * BM));
                    }
                    x_phase ^= 1;
                    cur_x = x_blk;
                }
                stream_y(y_tile);                             // for S of tile j
            }
            if (j >= 1) stream_y((t0 + j - 1) % y_tiles);     // again, for OUT of tile j - 1
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t smem_base = smem_u32(smem);
        constexpr uint32_t idesc_s = make_idesc_bf16(BM, false, false);      // S: N = 128, both operands K-major
        constexpr uint32_t idesc_o = make_idesc_bf16(BK, false, true);       // OUT chunk: N = 64, B MN-major
        constexpr uint32_t tile16 = (uint32_t)(BM * BK * 2) >> 4;            // hi -> lo distance in descriptor units
        int stage = 0;
        uint32_t phase = 0, x_phase = 0, out_phase = 0;
        int64_t cur_x = -1;
        int seg = 0;
        for (int j = 0; j <= n_my; ++j) {
            if (j < n_my) {
                const int64_t t = t0 + j, x_blk = t / y_tiles;
                if (x_blk != cur_x) {
                    mbar_wait(x_full, x_phase);
                    x_phase ^= 1;
                    cur_x = x_blk;
                }
                const uint32_t s_tmem = tmem_base + (uint32_t)((j & 1) * BM);
                // (the S stage is free: its previous content, P of tile j - 2, was consumed by MMAs issued earlier,
                //  and tcgen05.mma instructions execute in issue order)
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_hi = desc_lo(smem_base + (uint32_t)(kb * 2 * BM * BK * 2), false), a_lo = a_hi + tile16;
                    const uint32_t b_hi = desc_lo(smem_base + (uint32_t)(FB_X_BYTES + stage * FB_STAGE_BYTES), false);
                    const uint32_t b_lo = b_hi + tile16;
                    uint32_t accum = kb > 0 ? 1u : 0u;
                    if (split) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            umma_lo_elect(s_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc_s, accum);
                            accum = 1u;
                        }
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) umma_lo_elect(s_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc_s, 1u);
                    }
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_lo_elect(s_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc_s, accum);
                        accum = 1u;
                    }
                    umma_commit_elect(&empty[stage]);
                    if (++stage == FB_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_elect(&s_full[j & 1]);
                const bool last_of_x = (j + 1 == n_my) || ((t + 1) / y_tiles != x_blk);
                if (last_of_x) umma_commit_elect(x_empty);      // X may be overwritten once these MMAs retire
            }
            if (j >= 1) {
                const int jj = j - 1;
                const int64_t t = t0 + jj, x_blk = t / y_tiles;
                const bool first_of_x = (jj == 0) || ((t - 1) / y_tiles != x_blk);
                const bool last_of_x = (jj + 1 == n_my) || ((t + 1) / y_tiles != x_blk);
                mbar_wait(&p_full[jj & 1], (uint32_t)((jj >> 1) & 1));
                tcgen05_fence_after();
                if (first_of_x && seg > 0) {                    // the epilogue has flushed the previous OUT
                    mbar_wait(out_empty, out_phase);
                    out_phase ^= 1;
                    tcgen05_fence_after();
                }
                const uint32_t p_tmem = tmem_base + (uint32_t)((jj & 1) * BM);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t o_tmem = tmem_base + (uint32_t)(FB_OUT_COL + kb * BK);
                    // B = the Y k-block read MN-major: rows are the K index (y), +2048 B per 16 rows
                    const uint32_t b_hi = desc_lo(smem_base + (uint32_t)(FB_X_BYTES + stage * FB_STAGE_BYTES), true);
                    const uint32_t b_lo = b_hi + tile16;
                    uint32_t accum = first_of_x ? 0u : 1u;
#pragma unroll
                    for (int ks = 0; ks < BM / UMMA_K; ++ks) {
                        // P of y rows [16 ks, 16 ks + 16): 32-column chunk ks / 2 = [hi 16 columns | lo 16 columns]
                        const uint32_t p_hi = p_tmem + (uint32_t)(32 * (ks >> 1) + 8 * (ks & 1)), p_lo = p_hi + 16;
                        const uint32_t koff = (uint32_t)(ks * ((UMMA_K * 128) >> 4));
                        if (split) {
                            umma_ts_lo_elect(o_tmem, p_lo, b_hi + koff, idesc_o, accum);
                            umma_ts_lo_elect(o_tmem, p_hi, b_lo + koff, idesc_o, 1u);
                            accum = 1u;
                        }
                        umma_ts_lo_elect(o_tmem, p_hi, b_hi + koff, idesc_o, accum);
                        accum = 1u;
                    }
                    umma_commit_elect(&empty[stage]);
                    if (++stage == FB_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (last_of_x) {
                    umma_commit_elect(out_full);
                    ++seg;
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: S -> P in place, OUT flush ----------------
        const int q = warp & 3;                         // TMEM lane quarter
        const int part = (warp - 4) >> 2;               // which 64-column half of the tile
        constexpr float kL2e = 1.4426950408889634f;
        uint32_t out_phase = 0;
        int seg = 0;
        float rowsum = 0.f;
        int64_t cur_x = -1;
        float l2_row = 1e30f, cf_row = 0.f, bl_row = -1e30f;
        int g_row = -1;
        const int64_t first_cta_tiles = n_tiles;        // (silence unused warnings in one of the instantiations)
        (void)first_cta_tiles;
        for (int j = 0; j < n_my; ++j) {
            const int64_t t = t0 + j, x_blk = t / y_tiles, y_tile = t % y_tiles;
            const int64_t row = x_blk * BM + q * 32 + lane;        // x index owned by this thread
            if (x_blk != cur_x) {
                cur_x = x_blk;
                rowsum = 0.f;
                if (TRANSPOSED) {
                    bl_row = __ldg(pb.bl + row);                    // (padded arrays: no bounds test)
                } else {
                    l2_row = __ldg(pb.l2s + row);
                    cf_row = __ldg(pb.cfs + row);
                    g_row = __ldg(pb.g32 + row);
                }
            }
            {   // the per-column parameters of this tile: start the loads before waiting for S
                const int64_t c0 = y_tile * BM + part * 64;
                if (TRANSPOSED) {
                    prefetch_l1(pb.l2s + c0); prefetch_l1(pb.l2s + c0 + 32);
                    prefetch_l1(pb.cfs + c0); prefetch_l1(pb.cfs + c0 + 32);
                    prefetch_l1(pb.g32 + c0); prefetch_l1(pb.g32 + c0 + 32);
                } else {
                    prefetch_l1(pb.bl + c0); prefetch_l1(pb.bl + c0 + 32);
                }
            }
            mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
            tcgen05_fence_after();
            const uint32_t s_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((j & 1) * BM + part * 64);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float v[32];
                tmem_ld32(s_addr + (uint32_t)(32 * c), v);
                const int64_t col0 = y_tile * BM + part * 64 + c * 32;       // y index of v[0]
                if (TRANSPOSED) {
                    const float4* l4 = reinterpret_cast<const float4*>(pb.l2s + col0);
                    const float4* c4 = reinterpret_cast<const float4*>(pb.cfs + col0);
                    const int4* g4 = reinterpret_cast<const int4*>(pb.g32 + col0);
                    const int r32 = (int)row;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 l = __ldg(l4 + i), cf = __ldg(c4 + i);
                        const int4 g = __ldg(g4 + i);
                        float e0 = ex2_approx(fmaf(v[4 * i], kL2e, bl_row - l.x)) * cf.x;
                        float e1 = ex2_approx(fmaf(v[4 * i + 1], kL2e, bl_row - l.y)) * cf.y;
                        float e2 = ex2_approx(fmaf(v[4 * i + 2], kL2e, bl_row - l.z)) * cf.z;
                        float e3 = ex2_approx(fmaf(v[4 * i + 3], kL2e, bl_row - l.w)) * cf.w;
                        e0 -= g.x == r32 ? cf.x : 0.f;
                        e1 -= g.y == r32 ? cf.y : 0.f;
                        e2 -= g.z == r32 ? cf.z : 0.f;
                        e3 -= g.w == r32 ? cf.w : 0.f;
                        v[4 * i] = e0; v[4 * i + 1] = e1; v[4 * i + 2] = e2; v[4 * i + 3] = e3;
                        rowsum += (e0 + e1) + (e2 + e3);
                    }
                } else {
                    const float4* b4 = reinterpret_cast<const float4*>(pb.bl + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = __ldg(b4 + i);
                        v[4 * i] = ex2_approx(fmaf(v[4 * i], kL2e, b.x - l2_row)) * cf_row;
                        v[4 * i + 1] = ex2_approx(fmaf(v[4 * i + 1], kL2e, b.y - l2_row)) * cf_row;
                        v[4 * i + 2] = ex2_approx(fmaf(v[4 * i + 2], kL2e, b.z - l2_row)) * cf_row;
                        v[4 * i + 3] = ex2_approx(fmaf(v[4 * i + 3], kL2e, b.w - l2_row)) * cf_row;
                    }
                    if ((int64_t)g_row >= col0 && (int64_t)g_row < col0 + 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i == (int64_t)g_row) v[i] -= cf_row;
                    }
                }
                // bf16 hi / lo pairs: registers [0, 16) = hi of columns (2 r, 2 r + 1), [16, 32) = lo
                uint32_t pk[32];
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float a = v[2 * r], b = v[2 * r + 1];
                    const uint32_t h = pack_bf16x2(a, b);
                    pk[r] = h;
                    pk[16 + r] = split ? pack_bf16x2(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xffff0000u)) : 0u;
                }
                tmem_st32(s_addr + (uint32_t)(32 * c), pk);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[j & 1]);
            const bool last_of_x = (j + 1 == n_my) || ((t + 1) / y_tiles != x_blk);
            if (last_of_x) {
                // OUT of this x block (the part of it this CTA accumulated) -> partial slab `slot`
                const int64_t first_cta = fb_cta_of_tile(x_blk * y_tiles, n_tiles, gridDim.x);
                const int64_t slot = (int64_t)blockIdx.x - first_cta;
                mbar_wait(out_full, out_phase);
                out_phase ^= 1;
                tcgen05_fence_after();
                float* dst = pb.out_part + ((x_blk * pb.max_slots + slot) * BM + (q * 32 + lane)) * (int64_t)pb.d;
                const uint32_t o_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(FB_OUT_COL + part * 128);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int col = part * 128 + c * 32;
                    if (col >= pb.d) break;
                    float v[32];
                    tmem_ld32(o_addr + (uint32_t)(32 * c), v);
                    if (col + 32 <= pb.d) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            reinterpret_cast<float4*>(dst + col)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col + i < pb.d) dst[col + i] = v[i];
                    }
                }
                if (TRANSPOSED)
                    pb.db_part[((x_blk * pb.max_slots + slot) * 2 + part) * BM + q * 32 + lane] = rowsum;
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(out_empty);
                ++seg;
            }
        }
        (void)seg;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
}  // namespace c2dsr
