// Generic warp-specialised tcgen05 GEMM used by the rule contraction (K2), its two adjoints (K3, K4) and the
// diagnostics entry point.  D[128 x BLOCK_N] tiles, fp16 operands staged by TMA (128-byte swizzle) into a
// STAGES-deep mbarrier ring, fp32 accumulators double-buffered in tensor memory, epilogue in 4 warps.
//   warp 0 : TMA producer (one lane)        warp 1 : tcgen05.mma issuer (one lane)
//   warp 2 : TMEM allocator                 warps 4-7 : epilogue (TMEM -> registers -> global), one row per thread
// Operand "major-ness":
//   K-major  X[rows][K]  : smem tile = rows x 64 halves, 128 B per row, 8-row swizzle groups of 1024 B
//   MN-major X[K][rows]  : smem tile = (rows/64) atoms, each 64 k-rows x 128 B (64 row-elements), atoms 8192 B apart
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace rbn {

static constexpr int kBM = 128;      // tile rows (TMEM lanes)
static constexpr int kBK = 64;       // halves per k-block = one 128-byte swizzle span

struct GemmShape {
    int m_tiles, n_tiles, k_blocks;
    // 2-CTA kernel only: tiles [0, full_tiles) are computed whole; every later tile is split along K into tail_splits
    // work units whose partial sums the epilogue adds atomically (wave-quantisation fix).  tail_splits <= 1: no split.
    int full_tiles, tail_splits;
    // 2-CTA kernel only: every unit reports partial = true (N = 512: the max-shift epilogue needs both halves of a row,
    // so all tiles add into the fp32 buffer that k_rule_finalize turns into chart cells)
    int all_partial;
    // L2 eviction priority of the two operand streams (ptx::kEvict*; 0 = no hint)
    unsigned long long hint_a, hint_b;
    // 2-CTA kernel with segmented accumulation only: every unit of the launch is at most one segment long, so a unit needs
    // ONE accumulator and consecutive units alternate between the two -- the epilogue of a unit (partial sums -> red.add)
    // runs behind the next unit's MMAs instead of in front of them
    int short_units;
    // 2-CTA kernel only, stream-K: > 0 = the launch's (tile, k-block) space, tile-major, is one linear range cut into equal
    // pieces of sk_per k-blocks, one per CTA pair; a piece that crosses a tile boundary becomes two units.  Every unit is
    // partial.  Evens out launches with fewer tiles than CTA pairs at the price of at most (pairs + tiles) partial sums.
    int sk_per;
};

struct GemmUnit {
    int tile, kb0, kb1;
    bool partial;
};
__device__ __forceinline__ int gemm_units(const GemmShape& s) {
    const int total = s.m_tiles * s.n_tiles;
    if (s.tail_splits <= 1) return total;
    return s.full_tiles + (total - s.full_tiles) * s.tail_splits;
}
__device__ __forceinline__ GemmUnit gemm_unit(const GemmShape& s, int u) {
    GemmUnit r;
    if (s.tail_splits <= 1 || u < s.full_tiles) {
        r.tile = u; r.kb0 = 0; r.kb1 = s.k_blocks; r.partial = s.all_partial != 0;
    } else {
        const int v = u - s.full_tiles;
        const int per = (s.k_blocks + s.tail_splits - 1) / s.tail_splits;
        r.tile = s.full_tiles + v / s.tail_splits;
        r.kb0 = (v % s.tail_splits) * per;
        r.kb1 = min(s.k_blocks, r.kb0 + per);
        r.partial = true;
    }
    return r;
}

// The units of one CTA pair, in order; all three warp roles of the 2-CTA kernel walk the same sequence.
struct PairUnits {
    const GemmShape& s;
    int unit, npairs, total;
    long long lin, lin_end;
    __device__ PairUnits(const GemmShape& shape, int pair, int np) : s(shape), unit(pair), npairs(np), total(gemm_units(shape)) {
        const long long all = (long long)shape.m_tiles * shape.n_tiles * shape.k_blocks;
        lin = (long long)pair * shape.sk_per;
        lin_end = lin + shape.sk_per < all ? lin + shape.sk_per : all;
    }
    __device__ bool next(GemmUnit& gu) {
        if (s.sk_per > 0) {
            if (lin >= lin_end) return false;
            gu.tile = (int)(lin / s.k_blocks);
            gu.kb0 = (int)(lin - (long long)gu.tile * s.k_blocks);
            const long long left = lin_end - lin;
            gu.kb1 = (gu.kb0 + left < s.k_blocks) ? (int)(gu.kb0 + left) : s.k_blocks;
            gu.partial = true;
            lin += gu.kb1 - gu.kb0;
            return true;
        }
        while (unit < total) {
            gu = gemm_unit(s, unit);
            unit += npairs;
            if (gu.kb0 < gu.kb1) return true;       // (an empty K range cannot happen for sane splits)
        }
        return false;
    }
};

template <int BLOCK_N, int STAGES, int EPI_SMEM>
struct GemmSmem {
    static constexpr int A_BYTES = kBM * kBK * 2;
    static constexpr int B_BYTES = BLOCK_N * kBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFF = STAGES * STAGE_BYTES;          // staging tiles of a TMA-store epilogue
    static constexpr int BAR_OFF = EPI_OFF + EPI_SMEM;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;            // + barriers + alignment slack
};

__host__ __device__ constexpr uint32_t tmem_cols_pow2(int c) {
    return c <= 32 ? 32u : c <= 64 ? 64u : c <= 128 ? 128u : c <= 256 ? 256u : 512u;
}

// SEG > 0: segmented accumulation.  tcgen05.mma adds into its fp32 accumulator with truncation; over a K-chain of
// 65536 that is a systematic -3.9e-4 relative bias (measured).  With SEG k-blocks per segment, segment 0
// accumulates into accumulator B, every later segment into accumulator A (fresh), and the epilogue warps fold A
// into B with round-to-nearest adds (tcgen05.ld / tcgen05.st) between segments; the last segment is folded in the
// final epilogue read.  SEG == 0: plain accumulation, the two accumulators double-buffer consecutive tiles.
template <bool A_MN, bool B_MN, int BLOCK_N, int STAGES, int SEG, class Epi>
__global__ void __launch_bounds__(256, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape shape,
          const __grid_constant__ Epi epi) {
    using S = GemmSmem<BLOCK_N, STAGES, Epi::kSmem>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* segfull = tempty + 2;
    uint64_t* segfree = segfull + 1;
    uint32_t* tmem_slot = (uint32_t*)(segfree + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(2 * BLOCK_N);
    constexpr int SEGD = SEG > 0 ? SEG : 1;   // divisor that is never zero (SEG == 0 branches are dead)

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull[s], 1);
            ptx::mbar_init(&tempty[s], 128);
        }
        ptx::mbar_init(segfull, 1);
        ptx::mbar_init(segfree, 128);
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tmA);
        ptx::tma_prefetch_desc(&tmB);
    }
    if (warp == 2) ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_sync();                        // nothing above reads or writes global memory

    const int total_tiles = shape.m_tiles * shape.n_tiles;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile / shape.m_tiles, mt = tile - nt * shape.m_tiles;
                for (int kb = 0; kb < shape.k_blocks; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * S::STAGE_BYTES;
                    uint8_t* sB = sA + S::A_BYTES;
                    ptx::mbar_expect_tx(&full[stage], S::STAGE_BYTES);
                    if (!A_MN) {
                        ptx::tma_load_2d(sA, &tmA, &full[stage], kb * kBK, mt * kBM);
                    } else {
#pragma unroll
                        for (int a = 0; a < kBM / 64; ++a)
                            ptx::tma_load_2d(sA + a * 8192, &tmA, &full[stage], mt * kBM + a * 64, kb * kBK);
                    }
                    if (!B_MN) {
                        ptx::tma_load_2d(sB, &tmB, &full[stage], kb * kBK, nt * BLOCK_N);
                    } else {
#pragma unroll
                        for (int a = 0; a < BLOCK_N / 64; ++a)
                            ptx::tma_load_2d(sB + a * 8192, &tmB, &full[stage], nt * BLOCK_N + a * 64, kb * kBK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(kBM, BLOCK_N, A_MN, B_MN, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            uint32_t flush_waits = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                for (int kb = 0; kb < shape.k_blocks; ++kb) {
                    uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                    bool seg_start = (kb == 0);
                    if (SEG > 0) {
                        const int seg = kb / SEGD;
                        seg_start = (kb - seg * SEGD) == 0;
                        d_tmem = tmem_base + (seg == 0 ? 0 : BLOCK_N);
                        if (seg_start && seg >= 2) {           // accumulator A must have been folded into B
                            ptx::mbar_wait(segfree, flush_waits & 1);
                            ++flush_waits;
                            ptx::tc_fence_after();
                        }
                    }
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sA = ptx::smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint32_t sB = sA + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t ad = A_MN ? ptx::make_smem_desc(sA + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sA + k * 32, 16, 1024);
                        const uint64_t bd = B_MN ? ptx::make_smem_desc(sB + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sB + k * 32, 16, 1024);
                        ptx::mma_f16_ss(d_tmem, ad, bd, idesc, (uint32_t)(!(seg_start && k == 0)));
                    }
                    ptx::mma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (SEG > 0) {
                        const int seg = kb / SEGD;
                        const bool seg_end = ((kb + 1) % SEGD == 0) && (kb + 1 < shape.k_blocks);
                        if (seg_end && seg >= 1) ptx::mma_commit(segfull);   // a non-final segment in A: fold it
                    }
                }
                ptx::mma_commit(&tfull[acc]);
                if (SEG == 0) {
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                } else {
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t flushes = 0, nstore = 0;
        uint8_t* epi_smem = smem + S::EPI_OFF;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile / shape.m_tiles, mt = tile - nt * shape.m_tiles;
            uint32_t t0 = tmem_base + acc * BLOCK_N + lane_off, t1 = 0;
            if (SEG > 0) {
                const int nseg = (shape.k_blocks + SEGD - 1) / SEGD;
                t0 = tmem_base + lane_off;
                const uint32_t tA = tmem_base + BLOCK_N + lane_off;
                for (int sg = 1; sg + 1 < nseg; ++sg) {       // fold every non-final segment held in A into B
                    ptx::mbar_wait(segfull, flushes & 1);
                    ++flushes;
                    ptx::tc_fence_after();
                    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                        float v[32];
                        ptx::tmem_ld32_sum(t0, tA, c0, v);
                        ptx::tmem_st32(t0 + c0, v);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(segfree);
                }
                if (nseg >= 2) t1 = tA;
            }
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            epi(mt, nt, q * 32 + lane, t0, t1, epi_smem, nstore, false);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty[acc]);
            if (SEG == 0) {
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            } else {
                acc_phase ^= 1;
            }
        }
        if (Epi::kSmem > 0 && threadIdx.x == 128) ptx::bulk_wait_all();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// =======================================================================================================================
// 2-CTA variant (cta_group::2): a pair of CTAs on one TPC computes a 256 x BLOCK_N tile.  Each CTA stages its own 128
// rows of A and HALF of B (BLOCK_N/2 rows), so per-SM shared-memory traffic (TMA fills + MMA operand reads) drops
// from ~190 B/cycle -- above the 128 B/cycle an SM can move, which capped the 1-CTA kernel near 70 % -- to ~125.
// The even CTA (leader) issues the MMAs; its `full` barriers collect the bytes of both CTAs' TMA loads; commits are
// multicast to both CTAs; the odd CTA's epilogue threads arrive remotely on the leader's `tempty` / `segfree`.
// =======================================================================================================================
template <int BLOCK_N, int STAGES, int EPI_SMEM>
struct Gemm2Smem {
    static constexpr int A_BYTES = kBM * kBK * 2;
    static constexpr int B_BYTES = (BLOCK_N / 2) * kBK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = EPI_OFF + EPI_SMEM;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <bool A_MN, bool B_MN, int BLOCK_N, int STAGES, int SEG, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
k_tc_gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape shape,
           const __grid_constant__ Epi epi) {
    using S = Gemm2Smem<BLOCK_N, STAGES, Epi::kSmem>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;       // [2] (index 1: short-unit mode only)
    uint64_t* tempty = tfull + 2;           // [2]
    uint64_t* segfull = tempty + 2;
    uint64_t* segfree = segfull + 1;
    uint32_t* tmem_slot = (uint32_t*)(segfree + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = (rank == 0);
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(SEG > 0 ? 2 * BLOCK_N : BLOCK_N);   // SEG == 0 uses one accumulator
    constexpr int SEGD = SEG > 0 ? SEG : 1;
    constexpr int HALF_N = BLOCK_N / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], 256);
        }
        ptx::mbar_init(segfull, 1);
        ptx::mbar_init(segfree, 256);
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tmA);
        ptx::tma_prefetch_desc(&tmB);
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                         // peer barriers are initialised before any remote arrive / multicast
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_sync();                        // nothing above reads or writes global memory

    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;     // (m_tiles counts 256-row tiles here)
    const bool short_mode = SEG > 0 && shape.short_units != 0;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            PairUnits units(shape, pair, npairs);
            GemmUnit gu;
            while (units.next(gu)) {
                const int mt = gu.tile / shape.n_tiles, nt = gu.tile - mt * shape.n_tiles;   // n fastest: the pairs that
                const int row0 = mt * 256 + (int)rank * kBM;                 // share an A tile run side by side (L2 hit)
                const int col0 = nt * BLOCK_N + (int)rank * HALF_N;          // this CTA's half of B
                for (int kb = gu.kb0; kb < gu.kb1; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * S::STAGE_BYTES;
                    uint8_t* sB = sA + S::A_BYTES;
                    if (leader) ptx::mbar_expect_tx(&full[stage], 2 * S::STAGE_BYTES);
                    const uint64_t ha = shape.hint_a ? shape.hint_a : ptx::kEvictNormal;
                    const uint64_t hb = shape.hint_b ? shape.hint_b : ptx::kEvictNormal;
                    if (!A_MN) {
                        ptx::tma_load_2d_pair_hint(sA, &tmA, &full[stage], kb * kBK, row0, ha);
                    } else {
#pragma unroll
                        for (int a = 0; a < kBM / 64; ++a)
                            ptx::tma_load_2d_pair_hint(sA + a * 8192, &tmA, &full[stage], row0 + a * 64, kb * kBK, ha);
                    }
                    if (!B_MN) {
                        ptx::tma_load_2d_pair_hint(sB, &tmB, &full[stage], kb * kBK, col0, hb);
                    } else {
#pragma unroll
                        for (int a = 0; a < HALF_N / 64; ++a)
                            ptx::tma_load_2d_pair_hint(sB + a * 8192, &tmB, &full[stage], col0 + a * 64, kb * kBK, hb);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(256, BLOCK_N, A_MN, B_MN, 0);
            int stage = 0;
            uint32_t phase = 0, it = 0, flush_waits = 0;
            PairUnits units(shape, pair, npairs);
            GemmUnit gu;
            while (units.next(gu)) {
                const uint32_t acc = short_mode ? (it & 1) : 0;              // accumulator (and barrier pair) of this unit
                const uint32_t acc_phase = short_mode ? ((it >> 1) & 1) : (it & 1);
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                for (int kbg = gu.kb0; kbg < gu.kb1; ++kbg) {
                    const int kb = kbg - gu.kb0;
                    const int nkb = gu.kb1 - gu.kb0;
                    uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                    bool seg_start = (kb == 0);
                    if (SEG > 0 && !short_mode) {
                        const int seg = kb / SEGD;
                        seg_start = (kb - seg * SEGD) == 0;
                        d_tmem = tmem_base + (seg == 0 ? 0 : BLOCK_N);
                        if (seg_start && seg >= 2) {
                            ptx::mbar_wait(segfree, flush_waits & 1);
                            ++flush_waits;
                            ptx::tc_fence_after();
                        }
                    }
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sA = ptx::smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint32_t sB = sA + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t ad = A_MN ? ptx::make_smem_desc(sA + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sA + k * 32, 16, 1024);
                        const uint64_t bd = B_MN ? ptx::make_smem_desc(sB + k * 2048, 8192, 1024)
                                                 : ptx::make_smem_desc(sB + k * 32, 16, 1024);
                        ptx::mma_f16_ss_pair(d_tmem, ad, bd, idesc, (uint32_t)(!(seg_start && k == 0)));
                    }
                    ptx::mma_commit_pair(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (SEG > 0 && !short_mode) {
                        const int seg = kb / SEGD;
                        const bool seg_end = ((kb + 1) % SEGD == 0) && (kb + 1 < nkb);
                        if (seg_end && seg >= 1) ptx::mma_commit_pair(segfull);
                    }
                }
                ptx::mma_commit_pair(&tfull[acc]);
                ++it;
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t it = 0, flushes = 0, nstore = 0;
        uint8_t* epi_smem = smem + S::EPI_OFF;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        PairUnits units(shape, pair, npairs);
        GemmUnit gu;
        while (units.next(gu)) {
            const int mt = gu.tile / shape.n_tiles, nt = gu.tile - mt * shape.n_tiles;
            const uint32_t acc = short_mode ? (it & 1) : 0;
            const uint32_t acc_phase = short_mode ? ((it >> 1) & 1) : (it & 1);
            const uint32_t t0 = tmem_base + acc * BLOCK_N + lane_off;
            uint32_t t1 = 0;
            if (SEG > 0 && !short_mode) {
                const int nseg = (gu.kb1 - gu.kb0 + SEGD - 1) / SEGD;
                const uint32_t tA = tmem_base + BLOCK_N + lane_off;
                for (int sg = 1; sg + 1 < nseg; ++sg) {
                    ptx::mbar_wait(segfull, flushes & 1);
                    ++flushes;
                    ptx::tc_fence_after();
                    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                        float v[32];
                        ptx::tmem_ld32_sum(t0, tA, c0, v);
                        ptx::tmem_st32(t0 + c0, v);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive_leader(segfree);
                }
                if (nseg >= 2) t1 = tA;
            }
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            epi(mt * 2 + (int)rank, nt, q * 32 + lane, t0, t1, epi_smem, nstore, gu.partial);
            ptx::tc_fence_before();
            ptx::mbar_arrive_leader(&tempty[acc]);
            ++it;
        }
        if (Epi::kSmem > 0 && threadIdx.x == 128) ptx::bulk_wait_all();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                         // nobody leaves while the peer may still read its smem / barriers
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, TMEM_COLS);
}

template <bool A_MN, bool B_MN, int BLOCK_N, int STAGES, int SEG, class Epi>
int launch_tc_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmShape shape, const Epi& epi, int num_sms,
                    cudaStream_t st, const char* name, int kclass) {
    using S = Gemm2Smem<BLOCK_N, STAGES, Epi::kSmem>;
    auto kern = k_tc_gemm2<A_MN, B_MN, BLOCK_N, STAGES, SEG, Epi>;
    RBN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    int tiles = shape.m_tiles * shape.n_tiles;
    if (shape.tail_splits > 1) tiles = shape.full_tiles + (tiles - shape.full_tiles) * shape.tail_splits;
    int pairs = num_sms / 2;
    if (shape.sk_per > 0) {
        const long long all = (long long)shape.m_tiles * shape.n_tiles * shape.k_blocks;
        tiles = (int)((all + shape.sk_per - 1) / shape.sk_per);       // pieces; the caller sized sk_per for `pairs` of them
    }
    if (tiles < pairs) pairs = tiles;
    if (pairs < 1) return RBN_OK;
    { ProfScope ps(kclass, st); launch_pdl(kern, dim3(2 * pairs), dim3(256), S::TOTAL, st, tmA, tmB, shape, epi); }
    RBN_LAUNCH_CHECK(name);
    return RBN_OK;
}

// ---- host: TMA descriptor encoding (driver entry point fetched at run time; libcuda is never linked) --------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn();

// 2-D fp16 tensor [outer][inner] (inner contiguous), box = {64, box_outer}, 128-byte swizzle, OOB reads give zeros.
int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_outer);
// 2-D fp32 tensor [outer][inner], box = {box_inner, box_outer}, no swizzle (target of TMA reduce-add stores).
int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer);

// 3-D fp32 view [d2][d1][d0] (d0 contiguous), box = {box0, 1, box2}, no swizzle: rows that are a fixed column of
// consecutive d2 slices (the right children of a span in the start-major outside chart).
int make_tmap_3d_f32(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box2);

template <bool A_MN, bool B_MN, int BLOCK_N, int STAGES, int SEG, class Epi>
int launch_tc_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmShape shape, const Epi& epi, int num_sms,
                   cudaStream_t st, const char* name, int kclass) {
    using S = GemmSmem<BLOCK_N, STAGES, Epi::kSmem>;
    auto kern = k_tc_gemm<A_MN, B_MN, BLOCK_N, STAGES, SEG, Epi>;
    RBN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    int tiles = shape.m_tiles * shape.n_tiles;
    int grid = tiles < num_sms ? tiles : num_sms;
    if (grid < 1) return RBN_OK;
    { ProfScope ps(kclass, st); launch_pdl(kern, dim3(grid), dim3(256), S::TOTAL, st, tmA, tmB, shape, epi); }
    RBN_LAUNCH_CHECK(name);
    return RBN_OK;
}

}  // namespace rbn
