// Span-diagonal tensor-core path (tcgen05 + TMEM + TMA) of the inside and outside passes.
//
// Per span diagonal (width w), spans ("items", m = b * (L-w+1) + i) are processed in chunks:
//   inside :  K1 split-sum    Y[m][b,c]  = sum_j c_j * Lhat_j[b] * Rhat_j[c]          (per-item GEMM, K = w-1)
//             K2 rule GEMM    out[m][a]  = sum_{b,c} Y[m][b,c] * Wt[a][b,c]           (items x N^2 x N) + max-shift epilogue
//   outside:  K1 recompute Y, prepG (scaled upstream gradient), then
//             K3 gY[m][b,c]   = sum_a Ghat[m][a] * Wk[b,c][a]                        (items x N x N^2)
//             K4 gW[b,c][a]  += sum_m Y[m][b,c] * Gabs[m][a]                         (N^2 x items x N) expected rule counts
//             K5 alpha[i,j][b]+= c_j 2^eG sum_c gY[m][b,c] Rhat_j[c]; alpha[j,k][c] += c_j 2^eG sum_b gY[m][b,c] Lhat_j[b]
// Reference arithmetic being replaced: /root/reference/rbnet/pcfg.py:305-314 (per-split contraction), 387-398
// (structural mixture) and the autograd traversal of sequential.py:36.
// Scaling: every chart cell is mantissa (max in [0.5,1)) x 2^exp; operands are fp16 copies of the mantissas in a
// start-major (LC) and an end-major (RC) layout; W is pre-scaled per parent (wexp) so its fp16 copy keeps 11 bits.
#include "common.cuh"
#include "tc_gemm.cuh"
#include <limits.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

namespace rbn {

// kernels shared with the CUDA-core path (same chart layout)
__global__ void k_emission(int, int, int, const float*, const int*, const int*, float*, int*);
__global__ void k_root(int, int, const float*, const int*, const float*, const int*, float*, float*, float*, float*, int*);
__global__ void k_outside_seed(int, int, const float*, const int*, const float*, const float*, const float*, float*, float*);
__global__ void k_emission_bwd(int, int, int, const int*, const int*, const int*, const float*, const float*, float*);

// ---- TMA descriptor encoding ------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    });
    return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_outer) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return RBN_E_CUDA;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {64, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu stride=%llu box_outer=%u", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_outer);
        return RBN_E_CUDA;
    }
    return RBN_OK;
}

int make_tmap_2d_f32(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return RBN_E_CUDA;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(f32) failed (%d)", (int)r);
        return RBN_E_CUDA;
    }
    return RBN_OK;
}

int make_tmap_3d_f32(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box2) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return RBN_E_CUDA;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
    cuuint32_t box[3] = {box0, 1, box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d f32) failed (%d)", (int)r);
        return RBN_E_CUDA;
    }
    return RBN_OK;
}

// CTA-pair (cta_group::2) GEMMs for the rule contraction and the count GEMM; RBN_TC_1CTA=1 selects the 1-CTA kernels
static bool use_pairs() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RBN_TC_1CTA");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// =======================================================================================================================
// epilogues of the generic GEMM
// =======================================================================================================================
struct EpiStoreF32 {   // diagnostics: D -> float [M][N]
    static constexpr int kSmem = 0;
    float* D;
    int M, N, block_n;
    __device__ void operator()(int mt, int nt, int row, uint32_t t0, uint32_t t1, uint8_t* es, uint32_t& nstore, bool partial) const {
        (void)es; (void)nstore; (void)partial;
        int r = mt * kBM + row;
        for (int c0 = 0; c0 < block_n; c0 += 32) {
            float v[32];
            ptx::tmem_ld32_sum(t0, t1, c0, v);
            if (r < M)
                for (int j = 0; j < 32; ++j) {
                    int c = nt * block_n + c0 + j;
                    if (c < N) D[(size_t)r * N + c] = v[j];
                }
        }
    }
};

// K2: out[m][a] -> chart mantissas (fp32), exponents, fp16 operand copies
struct EpiRule {
    static constexpr int kSmem = 0;
    int N, L, Lp, w, m0, cnt;
    const int* lengths;
    const float* wscale; // W = Wt * wscale[a], wscale[a] = 2^wexp[a]
    const int* item_e;   // E_m
    float* cv;
    int* ce;
    __half* lc;
    __half* rc;
    float* acc;          // [chunk][N] fp32 partial sums of K-split tail tiles (finished by k_rule_finalize)
    int bn;              // parents per tile (N, or 256 when N = 512: every tile is then 'partial' and nt selects the half)
    __device__ void operator()(int mt, int nt, int row, uint32_t t0, uint32_t t1, uint8_t* es, uint32_t& nstore, bool partial) const {
        (void)es; (void)nstore;
        const int t = mt * kBM + row;
        if (partial) {
            float* dst = acc + (size_t)t * N + nt * bn;
            for (int c0 = 0; c0 < bn; c0 += 32) {
                float v[32];
                ptx::tmem_ld32_sum(t0, t1, c0, v);
                if (t < cnt) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(v[j]), "f"(v[j + 1]),
                                     "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                }
            }
            return;
        }
        const int nI = L - w + 1;
        const int m = m0 + t;
        const int b = m / nI, i = m - b * nI, k = i + w;
        const bool valid = (t < cnt) && (k <= lengths[t < cnt ? b : 0]);
        float mx = 0.f;
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            ptx::tmem_ld32_sum(t0, t1, c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j] * wscale[c0 + j]);
        }
        int e_new = 0;
        if (mx > 0.f) frexpf(mx, &e_new);
        const float rs = ldexpf(1.0f, -e_new);
        const size_t cell = (size_t)b * cells_per_seq(L) + cell_off(w, L) + i;
        float* cvp = cv + cell * N;
        __half* lcp = lc + ((size_t)((size_t)b * L + i) * Lp + k) * N;
        __half* rcp = rc + ((size_t)((size_t)b * (L + 1) + k) * Lp + i) * N;
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            ptx::tmem_ld32_sum(t0, t1, c0, v);
            if (valid) {
                __align__(16) __half h[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = v[j] * wscale[c0 + j] * rs;
                    h[j] = __float2half_rn(v[j]);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cvp + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    *reinterpret_cast<uint4*>(lcp + c0 + j) = *reinterpret_cast<const uint4*>(h + j);
                    *reinterpret_cast<uint4*>(rcp + c0 + j) = *reinterpret_cast<const uint4*>(h + j);
                }
            }
        }
        if (valid) ce[cell] = item_e[t] + e_new;
    }
};

// K3: gY (fp16) [items][N*N], written as full 128-byte lines by TMA stores of smem-staged [128 x 64] tiles
struct EpiGY {
    static constexpr int kSmem = 32768;
    CUtensorMap tm;      // gY as [cnt rows][N*N cols], box {64, 128}
    __device__ void operator()(int mt, int nt, int row, uint32_t t0, uint32_t t1, uint8_t* es, uint32_t& nstore, bool partial) const {
        (void)partial;
        for (int c0 = 0; c0 < 256; c0 += 64) {
            float v[64];
            ptx::tmem_ld32_sum(t0, t1, c0, v);
            ptx::tmem_ld32_sum(t0, t1, c0 + 32, v + 32);
            ptx::stage_store_tile(es, &tm, nt * 256 + c0, mt * kBM, row, v, nstore, 3, threadIdx.x == 128, true);
            ++nstore;
        }
    }
};

// K4: gW[(b,c)][a] += D * scale, scale = 2^-shift of this launch's gradient rows (k_gabs_convert)
static constexpr int kK2Stages = 7;      // TMA ring of the CTA-pair rule / count GEMMs: 7 x 32 KB is what fits (more bytes in flight)
static constexpr int kSegBlocks = 64;   // k-blocks (of 64) per accumulation segment: K chain of 4096 -> bias ~2e-5
struct EpiCounts {
    static constexpr int kSmem = 0;
    float* gW;
    int N;
    int bn;              // parents per tile (N, or 256 when N = 512)
    const float* scale;  // device scalar written by k_gabs_convert
    __device__ void operator()(int mt, int nt, int row, uint32_t t0, uint32_t t1, uint8_t* es, uint32_t& nstore, bool partial) const {
        (void)es; (void)nstore; (void)partial;
        float* dst = gW + ((size_t)mt * kBM + row) * N + nt * bn;
        for (int c0 = 0; c0 < bn; c0 += 32) {
            float v[32];
            ptx::tmem_ld32_sum(t0, t1, c0, v);
            const float sc = __ldg(scale);
            if (partial) {        // several K-split units add into the same tile
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(v[j] * sc),
                                 "f"(v[j + 1] * sc), "f"(v[j + 2] * sc), "f"(v[j + 3] * sc) : "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = *reinterpret_cast<float4*>(dst + c0 + j);
                    o.x += v[j] * sc;
                    o.y += v[j + 1] * sc;
                    o.z += v[j + 2] * sc;
                    o.w += v[j + 3] * sc;
                    *reinterpret_cast<float4*>(dst + c0 + j) = o;
                }
            }
        }
    }
};

// =======================================================================================================================
// K1: split-sum, one span per CTA iteration
// =======================================================================================================================
// Operands are K-slabs of up to 64 splits: TMA boxes of 64 chart rows x 64 nonterminals land as MN-major atoms
// (64 k-rows x 128 B, 128-byte swizzle).  Four "fix-up" warps fold the per-split scale c_j = 2^(e_ij + e_jk - E)
// into the left-child slab in shared memory (and zero the rows past the last split); one thread issues the MMAs.
//   warps 0-3 epilogue (TMEM -> fp16 -> Y)   warps 4-7 fix-up   warp 8 TMA producer   warp 9 MMA issuer / TMEM owner
//   warps 10-13 (RBN_TC_K1_GROUPS=2, one CTA per SM only): optional second epilogue group, one accumulator (= half of a
//   span) per group.  The kernel follows the SM clock (137 ms per C3 step on a 1.75 GHz box, 167 ms on a 1.5 GHz one), so
//   the epilogue (8 tiles of [128 x 64] per span: TMEM load, fp16 pack, swizzled st.shared, barrier, TMA store) looked like
//   the longest stage of the per-span pipeline; measured, the second group does not help (see launch_k1) and is off.
template <int MINB>
static constexpr int k1_threads() { return MINB == 1 ? 448 : 320; }

// Span m of a launch as (sequence b, start i), stepped without a division per item (the per-span control code is what
// bounds this kernel at N = 64: 2742 warp instructions per span in round 2's source-level profile, two integer
// divisions per role among them).
struct SpanIter {
    int b, i, sb, si, nI;
    __device__ SpanIter(int m_first, int step, int nI_) : nI(nI_) {
        b = m_first / nI_; i = m_first - b * nI_;
        sb = step / nI_; si = step - sb * nI_;
    }
    __device__ void next() { b += sb; i += si; if (i >= nI) { i -= nI; ++b; } }
    __device__ void succ(int& b1, int& i1) const { b1 = b; i1 = i + 1; if (i1 >= nI) { i1 = 0; ++b1; } }   // span m + 1
};

// MINB: co-resident CTAs per SM the register budget is cut for (2 for small N).  PAIR: the N = 64 pair mode below, as a
// template parameter so that the sub-block / half / K-slab loops of the general kernel fold away.
template <int MINB, bool PAIR>
__global__ void __launch_bounds__(k1_threads<MINB>(), MINB)
k_tc_splitsum(const __grid_constant__ CUtensorMap tmLC, const __grid_constant__ CUtensorMap tmRC,
              const __grid_constant__ CUtensorMap tmLC1, const __grid_constant__ CUtensorMap tmRC1,   // second K-slab (rows = Kpad - 64)
              const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmYtail,
              int N, int L, int Lp, int w, int m0, int cnt, int stages, const int* __restrict__ lengths,
              const int* __restrict__ ce, __half* __restrict__ Y, int* __restrict__ item_e, int nslots, uint64_t y_pol, int nbuf,
              int j0, int ns) {
    // [j0, j0 + ns): the window of a span's w - 1 splits this launch sums (ns <= 128: two K slabs of 64).  Spans of up to
    // 129 tokens are one window (j0 = 0, ns = w - 1); wider spans take one launch per window, each with its own scale
    // exponent, folded together by k_y_combine (launch_split_sum).  Pair mode: always the whole span.
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int Kpad = (ns + 15) & ~15;
    const int parts = PAIR ? 1 : (Kpad + 63) / 64;
    // N > 256: the N x N split-sum is cut into (N/256)^2 sub-blocks of 256 x 256, each handled like a span of its own
    const int NS = PAIR ? 64 : (N > 256 ? 256 : N);          // nonterminals per sub-block side
    const int nblk = PAIR ? 1 : N / NS;
    const int nblk_sh = nblk >> 1;             // nblk is 1 or 2: sub / nblk = sub >> nblk_sh
    // nslots == 1 (one 256-column accumulator): the two 128-row halves of a sub-block cannot accumulate side by side
    // over the K slabs, so every half becomes a sub-item of its own (its left-child slab has 128 nonterminals; the
    // right-child slab is loaded once per half)
    const bool hsplit = !PAIR && (nslots == 1) && (NS == 256);
    const int RB = hsplit ? 128 : NS;          // left nonterminals (rows of Y) per sub-item
    const int nsub = PAIR ? 1 : (N / RB) * nblk;
    const int halves = (PAIR || hsplit) ? 1 : (NS + 127) / 128;
    // N = 64: a span fills half of the 128-row MMA and its per-span hand-offs (barriers, TMEM drain, store) cost more
    // issue slots than its bytes cost HBM time.  Pair mode puts TWO consecutive spans of the diagonal into one item:
    // A = [L(t); L(t+1)] (two 64-row atoms), B = [R(t); R(t+1)], one M = 128 x N = 128 MMA chain; the diagonal blocks of
    // the accumulator are the two split-sums (the off-diagonal blocks are never read), and the 128 x 64 staging tile is
    // exactly Y[t], Y[t+1] -- one full-box store.  Needs all splits in one K slab (w - 1 <= 64).
    constexpr bool pairm = PAIR;               // host: N == 64 and Kpad <= 64
    constexpr int per = PAIR ? 2 : 1;          // spans per CTA iteration
    const int ACC = pairm ? 128 : NS;          // accumulator columns per slot
    const int a_atoms = halves * 2;            // A operand covers halves * 128 rows (zero beyond the loaded atoms)
    const int l_atoms = pairm ? 2 : RB / 64;   // left-child atoms actually loaded
    const int b_atoms = pairm ? 2 : NS / 64;
    // an operand atom is 64 nonterminals x up to 64 splits (128 B per split row).  Pair mode packs atoms at Kpad rows
    // instead of 64, so that the narrow diagonals (where most spans are) get a deeper ring out of the same shared memory:
    // a pair is a ~2 us chain of load -> fix-up -> MMA, and only the ring depth overlaps consecutive pairs
    const uint32_t atom_bytes = pairm ? (uint32_t)Kpad * 128 : 8192;
    const uint32_t stage_bytes = (uint32_t)(a_atoms + b_atoms) * atom_bytes;
    uint8_t* ybuf = smem + (size_t)stages * stage_bytes;       // nbuf x 16 KB staging tiles for the Y stores
    uint8_t* tail = ybuf + (size_t)nbuf * 16384;
    uint64_t* tma_full = (uint64_t*)tail;         // [stages]
    uint64_t* fix_full = tma_full + stages;       // [stages]
    uint64_t* empty = fix_full + stages;          // [stages]
    uint64_t* acc_full = empty + stages;          // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    int* red = (int*)(tmem_slot + 2);             // [4]
    float* cj = (float*)(red + 4);                // [128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // nslots = 2: two accumulators (the halves of a span, or consecutive spans when a span has one half); nslots = 1:
    // a single accumulator (256 TMEM columns), used when this kernel shares its SM with a GEMM CTA (co-resident schedule)
    const uint32_t TMEM_COLS = tmem_cols_pow2(nslots * ACC);
    const int nI = L - w + 1;
    const size_t C = cells_per_seq(L);
    const int step = (int)gridDim.x * per;     // spans between consecutive items of this CTA
    // span t of this launch, at (sequence b, start i), exists and lies inside its sequence
    auto span_ok = [&](int t, int b, int i) { return t < cnt && i + w <= lengths[b]; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            ptx::mbar_init(&tma_full[s], 1);
            ptx::mbar_init(&fix_full[s], 128);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int h = 0; h < 2; ++h) {
            ptx::mbar_init(&acc_full[h], 1);
            ptx::mbar_init(&acc_empty[h], 128);
        }
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tmLC);
        ptx::tma_prefetch_desc(&tmRC);
    }
    // zero the A atoms that lie beyond N (N = 64 or 192) in every stage: TMA never writes them
    if (a_atoms > l_atoms)
        for (int sidx = 0; sidx < stages; ++sidx)
            for (uint32_t off = (uint32_t)l_atoms * 8192 + threadIdx.x * 16; off < (uint32_t)a_atoms * 8192; off += blockDim.x * 16)
                *reinterpret_cast<uint4*>(smem + (size_t)sidx * stage_bytes + off) = make_uint4(0, 0, 0, 0);
    if (warp == 9) ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_sync();                        // nothing above reads or writes global memory

    if (warp == 8) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            SpanIter it(m0 + blockIdx.x * per, step, nI);
            for (int t = blockIdx.x * per; t < cnt; t += step, it.next()) {
                const int b = it.b, i = it.i, k = i + w;
                if (pairm) {
                    int b1, i1;
                    it.succ(b1, i1);
                    const bool ok0 = span_ok(t, b, i), ok1 = span_ok(t + 1, b1, i1);
                    if (!ok0 && !ok1) continue;
                    // a span beyond its sequence borrows its partner's (finite) rows; its factors c_j are all zero
                    const int ba = ok0 ? b : b1, ia = ok0 ? i : i1, bb = ok1 ? b1 : b, ib = ok1 ? i1 : i;
                    const int lrow0 = (int)(((size_t)ba * L + ia) * Lp + (ia + 1)), rrow0 = (int)(((size_t)ba * (L + 1) + ia + w) * Lp + (ia + 1));
                    const int lrow1 = (int)(((size_t)bb * L + ib) * Lp + (ib + 1)), rrow1 = (int)(((size_t)bb * (L + 1) + ib + w) * Lp + (ib + 1));
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sL = smem + (size_t)stage * stage_bytes;
                    uint8_t* sR = sL + 2 * atom_bytes;
                    ptx::mbar_expect_tx(&tma_full[stage], 4u * (uint32_t)Kpad * 128);
                    ptx::tma_load_2d(sL, &tmLC, &tma_full[stage], 0, lrow0);
                    ptx::tma_load_2d(sL + atom_bytes, &tmLC, &tma_full[stage], 0, lrow1);
                    ptx::tma_load_2d(sR, &tmRC, &tma_full[stage], 0, rrow0);
                    ptx::tma_load_2d(sR + atom_bytes, &tmRC, &tma_full[stage], 0, rrow1);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                    continue;
                }
                if (k > lengths[b]) continue;
                const int lc_row = (int)(((size_t)b * L + i) * Lp + (i + 1) + j0);
                const int rc_row = (int)(((size_t)b * (L + 1) + k) * Lp + (i + 1) + j0);
                for (int sub = 0; sub < nsub; ++sub) {
                    const int bcol = (sub >> nblk_sh) * RB, ccol = (sub & (nblk - 1)) * NS;
                    for (int part = 0; part < parts; ++part) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sL = smem + (size_t)stage * stage_bytes;
                        uint8_t* sR = sL + (size_t)a_atoms * 8192;
                        // the box of this slab holds exactly Ks = min(64, Kpad - 64 part) chart rows (its atom stays 8 KB apart)
                        const int Ks = min(64, Kpad - part * 64);
                        ptx::mbar_expect_tx(&tma_full[stage], (uint32_t)(l_atoms + b_atoms) * (uint32_t)Ks * 128);
                        for (int a = 0; a < l_atoms; ++a)
                            ptx::tma_load_2d(sL + a * 8192, part ? &tmLC1 : &tmLC, &tma_full[stage], bcol + a * 64, lc_row + part * 64);
                        for (int a = 0; a < b_atoms; ++a)
                            ptx::tma_load_2d(sR + a * 8192, part ? &tmRC1 : &tmRC, &tma_full[stage], ccol + a * 64, rc_row + part * 64);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ---------------- fix-up: fold c_j into the left-child slab ----------------
        const int p = threadIdx.x - 128;                // 0..127
        int stage = 0;
        uint32_t phase = 0;
        if constexpr (PAIR) {
            // Pair mode: thread (span psp, split pj) scales ITS OWN row of its span's left-child atom, and every warp finds
            // its span's scale exponent by itself (a thread also loads the exponents of split pj ^ 32, so the 32 lanes of
            // a warp see all 64 splits): no shared factors and no named barrier in this role.
            const int psp = p >> 6, pj = p & 63;
            auto fetch2 = [&](int t, int b, int i, int& ea, int& eb, int& oa, int& ob) {     // (b, i): span t of the launch
                ea = oa = INT_MIN; eb = ob = 0;
                if (t >= cnt || i + w > lengths[b]) return;
                const int* ce_seq = ce + (size_t)b * C;
                const int u = pj + 1, v = (pj ^ 32) + 1;
                if (u < w) { ea = ce_seq[cell_off(u, L) + i]; eb = ce_seq[cell_off(w - u, L) + i + u]; }
                if (v < w) { oa = ce_seq[cell_off(v, L) + i]; ob = ce_seq[cell_off(w - v, L) + i + v]; }
            };
            SpanIter it(m0 + blockIdx.x * 2, step, nI);
            auto my_span = [&](int& b, int& i) { b = it.b; i = it.i; if (psp) it.succ(b, i); };   // this thread's span of the item
            int ea, eb, oa, ob, bm, im;
            my_span(bm, im);
            fetch2(blockIdx.x * 2 + psp, bm, im, ea, eb, oa, ob);
            for (int t = blockIdx.x * 2; t < cnt; t += step) {
                const int sft = (ea == INT_MIN) ? INT_MIN : ea + eb;
                const int oft = (oa == INT_MIN) ? INT_MIN : oa + ob;
                bool ok = span_ok(t, it.b, it.i);
                if (!ok) { int b1, i1; it.succ(b1, i1); ok = span_ok(t + 1, b1, i1); }
                it.next();
                my_span(bm, im);
                fetch2(t + step + psp, bm, im, ea, eb, oa, ob);      // in flight while this item is processed
                if (!ok) continue;
                int E = max(sft, oft);
                for (int o = 16; o > 0; o >>= 1) E = max(E, __shfl_xor_sync(0xffffffffu, E, o));
                // c_j = 2^(sft - E), exponent clamped at -60 (zero in fp16 either way); 0 for rows past the last split and
                // for every row of a span beyond its sequence
                const float c = (sft != INT_MIN) ? __int_as_float((max(sft - E, -60) + 127) << 23) : 0.f;
                if (pj == 0 && E != INT_MIN) item_e[t + psp] = E;
                ptx::mbar_wait(&tma_full[stage], phase);
                if (pj < Kpad && c != 1.0f) {
                    uint8_t* rowp = smem + (size_t)stage * stage_bytes + (uint32_t)psp * atom_bytes + (uint32_t)pj * 128;
                    const __half2 sc = __float2half2_rn(c);
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        uint4* ptr = reinterpret_cast<uint4*>(rowp + ((ch ^ (pj & 7)) << 4));
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (c != 0.f) {
                            v = *ptr;
                            __half2* hv = reinterpret_cast<__half2*>(&v);
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) hv[qd] = __hmul2(hv[qd], sc);
                        }
                        *ptr = v;
                    }
                }
                ptx::fence_proxy_async_smem();
                ptx::mbar_arrive(&fix_full[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        } else {
        // 16-byte chunks per left-child slab row
        const int chunks_per_row = RB / 8;
        const int psp = 0, pj = p;
        // e_left + e_right of this thread's split of span t (INT_MIN: no such split / span beyond its sequence)
        // (kept as two registers and added at the point of use, so that the loads stay in flight across the iteration)
        auto fetch_sft = [&](int t, int b, int i, int& ea, int& eb) {       // (b, i): span t of the launch
            ea = INT_MIN; eb = 0;
            if (t >= cnt || pj >= ns || i + w > lengths[b]) return;
            const int* ce_seq = ce + (size_t)b * C;
            const int u = j0 + pj + 1;
            ea = ce_seq[cell_off(u, L) + i];
            eb = ce_seq[cell_off(w - u, L) + i + u];
        };
        SpanIter it(m0 + blockIdx.x * per, step, nI);
        auto my_span = [&](int& b, int& i) { b = it.b; i = it.i; if (pairm && psp) it.succ(b, i); };   // this thread's span of the item
        int ea_next, eb_next, bm, im;
        my_span(bm, im);
        fetch_sft(blockIdx.x * per + psp, bm, im, ea_next, eb_next);
        for (int t = blockIdx.x * per; t < cnt; t += step) {
            const int sft = (ea_next == INT_MIN) ? INT_MIN : ea_next + eb_next;
            bool ok = span_ok(t, it.b, it.i);
            if (pairm && !ok) { int b1, i1; it.succ(b1, i1); ok = span_ok(t + 1, b1, i1); }
            it.next();
            my_span(bm, im);
            fetch_sft(t + step + psp, bm, im, ea_next, eb_next);   // in flight while this item is processed (global-load latency)
            if (!ok) continue;
            int mxs = sft;
            for (int o = 16; o > 0; o >>= 1) mxs = max(mxs, __shfl_xor_sync(0xffffffffu, mxs, o));
            ptx::named_bar_sync(1, 128);                // every fix-up thread is done with the previous span's cj / red
            if (lane == 0) red[warp - 4] = mxs;
            ptx::named_bar_sync(1, 128);
            // the span's scale exponent: over all four warps, or over the two warps that own the span in pair mode
            const int E = pairm ? max(red[2 * psp], red[2 * psp + 1]) : max(max(red[0], red[1]), max(red[2], red[3]));
            cj[p] = (sft != INT_MIN) ? ldexpf(1.0f, max(sft - E, -60)) : 0.f;
            if (pj == 0 && E != INT_MIN) item_e[t + psp] = E;
            ptx::named_bar_sync(1, 128);
            for (int sp = 0, part = 0; sp < nsub * parts; ++sp, part = (part + 1 == parts) ? 0 : part + 1) {
                const int Ks = min(64, Kpad - part * 64);
                ptx::mbar_wait(&tma_full[stage], phase);
                uint8_t* sL = smem + (size_t)stage * stage_bytes;
                for (int idx = p; idx < Ks * chunks_per_row; idx += 128) {
                    const int r = idx / chunks_per_row, c16 = idx - r * chunks_per_row;
                    const float c = cj[pairm ? ((c16 >> 3) * 64 + r) : (part * 64 + r)];
                    if (c == 1.0f) continue;
                    uint4* ptr = reinterpret_cast<uint4*>(sL + (uint32_t)(c16 >> 3) * atom_bytes + (uint32_t)r * 128 +
                                                         (uint32_t)(((c16 & 7) ^ (r & 7)) << 4));
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (c != 0.f) {
                        v = *ptr;
                        const __half2 sc = __float2half2_rn(c);
                        __half2* hv = reinterpret_cast<__half2*>(&v);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) hv[qd] = __hmul2(hv[qd], sc);
                    }
                    *ptr = v;
                }
                ptx::fence_proxy_async_smem();
                ptx::mbar_arrive(&fix_full[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        }
    } else if (warp == 9) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = ptx::make_idesc_f16(128, ACC, true, true, 0);
            int stage = 0;
            uint32_t phase = 0, nacc = 0;       // nacc: accumulators used so far; slot = n & 1, its phase = (n >> 1) & 1
            SpanIter it(m0 + blockIdx.x * per, step, nI);
            for (int t = blockIdx.x * per; t < cnt; t += step, it.next()) {
                bool ok = span_ok(t, it.b, it.i);
                if (pairm && !ok) { int b1, i1; it.succ(b1, i1); ok = span_ok(t + 1, b1, i1); }
                if (!ok) continue;
                for (int sub = 0; sub < nsub; ++sub) {
                for (int part = 0; part < parts; ++part) {
                    const int Ks = min(64, Kpad - part * 64);
                    ptx::mbar_wait(&fix_full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t aL = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t aR = aL + (uint32_t)a_atoms * atom_bytes;
                    for (int h = 0; h < halves; ++h) {
                        const uint32_t na = nacc + h;
                        const uint32_t slot = nslots == 2 ? (na & 1) : 0, slot_phase = nslots == 2 ? ((na >> 1) & 1) : (na & 1);
                        if (part == 0) {
                            ptx::mbar_wait(&acc_empty[slot], slot_phase ^ 1);
                            ptx::tc_fence_after();
                        }
                        for (int k16 = 0; k16 < Ks / 16; ++k16) {
                            const uint64_t ad = ptx::make_smem_desc(aL + (uint32_t)(h * 2) * atom_bytes + k16 * 2048, atom_bytes, 1024);
                            const uint64_t bd = ptx::make_smem_desc(aR + k16 * 2048, atom_bytes, 1024);
                            ptx::mma_f16_ss(tmem_base + slot * ACC, ad, bd, idesc, (uint32_t)!(part == 0 && k16 == 0));
                        }
                        if (part == parts - 1) ptx::mma_commit(&acc_full[slot]);
                    }
                    ptx::mma_commit(&empty[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                nacc += halves;
                }
            }
        }
    } else if (warp < 4 || warp >= 10) {
        // ---------------- epilogue: TMEM -> fp16 -> smem tile -> TMA store into Y[t][b][c] ----------------
        // Two groups (warps 0-3 and, when present, 10-13): with two accumulators, group g drains accumulator g -- half g of
        // every span when a span has two halves, every other sub-item otherwise -- through its own staging tile.
        const int grp = warp >= 10 ? 1 : 0;
        const bool two_groups = (blockDim.x > 320) && (nslots == 2);
        if (grp == 1 && !two_groups) goto k1_done;
        {
        uint32_t nacc = 0, nstore = 0;
        const int qd = warp & 3;                      // TMEM lane quarter this warp may touch
        const int row = qd * 32 + lane;
        const int gt = grp ? (int)threadIdx.x - 320 : (int)threadIdx.x;      // thread index within the group
        const bool issuer = (gt == 0);
        // two groups: the four staging tiles are split two and two (one store in flight per group) when there are four
        uint8_t* gbuf = two_groups ? ybuf + grp * (nbuf == 4 ? 32768 : 16384) : ybuf;
        SpanIter it(m0 + blockIdx.x * per, step, nI);
        for (int t = blockIdx.x * per; t < cnt; t += step, it.next()) {
            bool ok = span_ok(t, it.b, it.i);
            if (pairm && !ok) { int b1, i1; it.succ(b1, i1); ok = span_ok(t + 1, b1, i1); }
            if (!ok) {
                // span beyond its sequence (ragged batch): its Y rows are still read (as A rows of the rule GEMM whose
                // output row is dropped, as K rows of the count GEMM against a zero gradient row), so they must be
                // finite -- write zeros; nothing else ever clears these buffers.  (Pair mode with ONE span outside: its
                // zero factors make its block of the accumulator zero and the pair's store writes it.)
                if (grp == 0) {
                    uint4* dst = reinterpret_cast<uint4*>(Y + (size_t)t * N * N);
                    const int nq = ((pairm && t + 1 < cnt) ? 2 : 1) * N * N / 8;
                    for (int q = gt; q < nq; q += 128) dst[q] = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            for (int sub = 0; sub < nsub; ++sub) {
                const int brow = (sub >> nblk_sh) * RB, ccol = (sub & (nblk - 1)) * NS;
                for (int h = 0; h < halves; ++h) {
                    const uint32_t na = nacc + h;
                    const uint32_t slot = nslots == 2 ? (na & 1) : 0, slot_phase = nslots == 2 ? ((na >> 1) & 1) : (na & 1);
                    if (two_groups && (int)slot != grp) continue;        // the other group's accumulator
                    ptx::mbar_wait(&acc_full[slot], slot_phase);
                    ptx::tc_fence_after();
                    // pair mode: rows 64-127 are span t + 1, whose split-sum is columns 64-127 of the accumulator
                    const uint32_t taddr = tmem_base + slot * ACC + (pairm ? (uint32_t)(qd >> 1) * 64 : 0u) + ((uint32_t)(qd * 32) << 16);
                    // only 64 valid rows: N = 192's second half, or a pair whose second span is past the launch
                    const bool tail_half = pairm ? (t + 1 >= cnt) : (h * 128 + 128 > NS);
                    for (int c0 = 0; c0 < NS; c0 += 64) {
                        float v[64];
                        ptx::tmem_ld32(taddr + c0, v);
                        ptx::tmem_ld32(taddr + c0 + 32, v + 32);
                        ptx::tmem_ld_wait();
                        if (PAIR) {              // the whole accumulator is in registers: hand it back before packing / storing
                            ptx::tc_fence_before();
                            ptx::mbar_arrive(&acc_empty[slot]);
                        }
                        if (PAIR)                // two co-resident CTAs: two staging tiles
                            ptx::stage_store_tile(gbuf, tail_half ? &tmYtail : &tmY, ccol + c0, t * N + brow + h * 128, row, v, nstore, 3, issuer, true, y_pol);
                        else if (two_groups && nbuf == 4)
                            ptx::stage_store_tile<2>(gbuf, tail_half ? &tmYtail : &tmY, ccol + c0, t * N + brow + h * 128, row, v, nstore, 3 + grp, issuer, true, y_pol);
                        else if (two_groups)     // one staging tile per group (the two groups' stores interleave)
                            ptx::stage_store_tile<1>(gbuf, tail_half ? &tmYtail : &tmY, ccol + c0, t * N + brow + h * 128, row, v, nstore, 3 + grp, issuer, true, y_pol);
                        else if (nbuf == 4)
                            ptx::stage_store_tile<4>(gbuf, tail_half ? &tmYtail : &tmY, ccol + c0, t * N + brow + h * 128, row, v, nstore, 3, issuer, true, y_pol);
                        else if (nbuf == 3)
                            ptx::stage_store_tile<3>(gbuf, tail_half ? &tmYtail : &tmY, ccol + c0, t * N + brow + h * 128, row, v, nstore, 3, issuer, true, y_pol);
                        else
                            ptx::stage_store_tile(gbuf, tail_half ? &tmYtail : &tmY, ccol + c0, t * N + brow + h * 128, row, v, nstore, 3, issuer, true, y_pol);
                        ++nstore;
                    }
                    if (!PAIR) {
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(&acc_empty[slot]);
                    }
                }
                nacc += halves;
            }
        }
        if (issuer) ptx::bulk_wait_all();
        }
    }
k1_done:
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// =======================================================================================================================
// K5: child gradients, one span per CTA iteration, TMA-fed
// =======================================================================================================================
// Operand ring: a stage is a 16 KB gY tile + the Npad x 64 child-row tile (Npad * 128 B, at most 16 KB).  The ring takes
// whatever shared memory is left beside the reduce staging tiles: 5 stages at 128 splits, 8 (the cap) from 33 splits down.
// A span is 16 stage loads; with 4 stages of 32 KB narrow diagonals were bound by the bytes in flight per SM (5.8 us per
// span at w = 2 against 3 us of HBM time), which is where most spans are.
static constexpr int kK5MaxStages = 8;

__global__ void __launch_bounds__(384, 1)
k_tc_childgrad(const __grid_constant__ CUtensorMap tmGYk,   // gY as [rows=(t,b)][c], box {64,128}: A of the left-child GEMM
               const __grid_constant__ CUtensorMap tmGYm,   // same array, box {64,64}: MN-major A of the right-child GEMM
               const __grid_constant__ CUtensorMap tmRC,    // RC rows, box {64, Npad}
               const __grid_constant__ CUtensorMap tmLC,    // LC rows, box {64, Npad}
               int N, int L, int Lp, int w, int m0, int cnt, const int* __restrict__ lengths,
               const __grid_constant__ CUtensorMap tmAL,    // start-major outside chart (fp32) [rows (b,start,end)][N], box {128, 32}
               const __grid_constant__ CUtensorMap tmAR,    // the SAME chart as [b*L + start][end][N], box {128, 1, 32}: right children
               const __grid_constant__ CUtensorMap tmALt,   // same views, tail_rows instead of 32: last chunk of a span's splits
               const __grid_constant__ CUtensorMap tmARt,
               const int* __restrict__ ce, const int* __restrict__ item_e, const int* __restrict__ gexp,
               uint64_t red_pol, uint64_t gy2_pol, int stages, int j0, int ns) {
    // [j0, j0 + ns): the window of a span's w - 1 splits whose children this launch serves (ns <= 128; wider spans take one
    // launch per window: the windows' child gradients are independent outputs, gY is read once per window)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = 16384 + (((ns + 15) & ~15) * 128);
    uint8_t* rbuf = smem + stages * stage_bytes;        // [2 groups][2] x 16 KB tiles [32 splits][128 nonterminals] fp32
    uint64_t* full = (uint64_t*)(rbuf + 65536);
    uint64_t* empty = full + kK5MaxStages;
    uint64_t* afull = empty + kK5MaxStages;  // [4]
    uint64_t* aempty = afull + 4;            // [4]
    uint32_t* tmem_slot = (uint32_t*)(aempty + 4);
    float* fac = (float*)(tmem_slot + 2);    // [2 groups][2][128] split factors c_j * 2^eG of the current / next span

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Npad = (ns + 15) & ~15;
    const int halves = (N + 127) / 128;
    const int units = 2 * halves;            // unit = kind * halves + h ; kind 0 = left-child, 1 = right-child gradient
    const int kblocks = N / 64;
    const int nI = L - w + 1;
    const size_t C = cells_per_seq(L);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 4; ++s) {
            ptx::mbar_init(&afull[s], 1);
            ptx::mbar_init(&aempty[s], 128);
        }
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tmGYk);
        ptx::tma_prefetch_desc(&tmGYm);
        ptx::tma_prefetch_desc(&tmRC);
        ptx::tma_prefetch_desc(&tmLC);
    }
    if (warp == 2) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_sync();                        // nothing above reads or writes global memory
    const uint32_t b_bytes = (uint32_t)Npad * 128;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < cnt; t += gridDim.x) {
                const int m = m0 + t;
                const int b = m / nI, i = m - b * nI, k = i + w;
                if (k > lengths[b]) continue;
                const int rc_row = (int)(((size_t)b * (L + 1) + k) * Lp + (i + 1) + j0);
                const int lc_row = (int)(((size_t)b * L + i) * Lp + (i + 1) + j0);
                for (int u = 0; u < units; ++u) {
                    const int kind = u / halves, h = u - kind * halves;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * stage_bytes;
                        uint8_t* sB = sA + 16384;
                        ptx::mbar_expect_tx(&full[stage], 16384 + b_bytes);
                        if (kind == 0) {
                            ptx::tma_load_2d(sA, &tmGYk, &full[stage], kb * 64, t * N + h * 128);
                            ptx::tma_load_2d(sB, &tmRC, &full[stage], kb * 64, rc_row);
                        } else {
                            // second (and last) read of this span's gY: nothing will want these lines again
                            ptx::tma_load_2d_hint(sA, &tmGYm, &full[stage], h * 128, t * N + kb * 64, gy2_pol);
                            ptx::tma_load_2d_hint(sA + 8192, &tmGYm, &full[stage], h * 128 + 64, t * N + kb * 64, gy2_pol);
                            ptx::tma_load_2d(sB, &tmLC, &full[stage], kb * 64, lc_row);
                        }
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_k = ptx::make_idesc_f16(128, Npad, false, false, 0);
            const uint32_t idesc_m = ptx::make_idesc_f16(128, Npad, true, false, 0);
            // This one thread issues 64 MMAs per span; on the narrow diagonals (w <= 33, 44 % of all spans) ITS instruction
            // count, not HBM, set the pace: 126 SASS instructions per stage, most of them descriptor arithmetic, 5.85 us per
            // span against 3.7 us of HBM time (source-level profile at w = 8).  The descriptors are therefore kept as
            // templates (zero address) plus the stage's address >> 4, stepped incrementally.
            const uint64_t tmpl_k = ptx::make_smem_desc(0, 16, 1024), tmpl_m = ptx::make_smem_desc(0, 8192, 1024);
            // (18 address bits, as make_smem_desc keeps: in a cluster launch the CTA rank sits above them)
            const uint32_t base16 = (ptx::smem_u32(smem) & 0x3FFFFu) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
            uint32_t sA16 = base16;       // shared-memory address >> 4 of the current stage
            int stage = 0;
            uint32_t phase = 0;
            uint32_t nunit = 0;           // running unit counter: accumulator slot = nunit % 4, its phase = (nunit / 4) & 1
            for (int t = blockIdx.x; t < cnt; t += gridDim.x) {
                const int m = m0 + t;
                const int b = m / nI, i = m - b * nI;
                if (i + w > lengths[b]) continue;
                for (int u = 0; u < units; ++u, ++nunit) {
                    const bool kind = u >= halves;
                    const uint64_t a_tmpl = kind ? tmpl_m : tmpl_k;
                    const uint32_t a_step = kind ? (2048u >> 4) : (32u >> 4);      // next 16 of K: 2 KB (MN-major) / 32 B (K-major)
                    const uint32_t idesc = kind ? idesc_m : idesc_k;
                    const uint32_t slot = nunit & 3, slot_phase = (nunit >> 2) & 1;
                    ptx::mbar_wait(&aempty[slot], slot_phase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + slot * 128;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&full[stage], phase);
                        ptx::tc_fence_after();
                        const uint64_t ad = a_tmpl + sA16, bd = tmpl_k + (sA16 + (16384u >> 4));
                        ptx::mma_f16_ss(d_tmem, ad, bd, idesc, (uint32_t)(kb != 0));
                        ptx::mma_f16_ss(d_tmem, ad + a_step, bd + 2, idesc, 1u);
                        ptx::mma_f16_ss(d_tmem, ad + 2 * a_step, bd + 4, idesc, 1u);
                        ptx::mma_f16_ss(d_tmem, ad + 3 * a_step, bd + 6, idesc, 1u);
                        ptx::mma_commit(&empty[stage]);
                        if (++stage == stages) { stage = 0; phase ^= 1; sA16 = base16; } else sA16 += stage16;
                    }
                    ptx::mma_commit(&afull[slot]);
                }
            }
        }
    } else if (warp >= 4) {
        // Epilogue: D[x, split] (TMEM) * factor -> smem tile [32 splits][128 nonterminals] -> TMA reduce-add into the
        // outside chart.  Left-child gradients go to the start-major chart (rows of one start are contiguous),
        // right-child gradients to the end-major chart; the fp32 add happens at L2, the SM never reads the chart.
        // Two groups of four warps (4-7, 8-11) take alternate units.
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int p = (threadIdx.x - 128) & 127;
        const bool issuer = (p == 0);
        uint32_t nunit0 = 0, nred = 0;      // nunit0: running unit counter at the start of the current span
        int fbuf = 0;
        uint8_t* mybuf = rbuf + grp * 32768;
        // exponent pieces of the split factor c_j * 2^eG of this thread's split of span t, fetched one span ahead and
        // combined at the point of use so that the loads stay in flight (INT_MIN: padding split, factor 0)
        auto fetch_fac = [&](int t, int& e0, int& e1, int& e2, int& e3) {
            e0 = INT_MIN; e1 = e2 = e3 = 0;
            if (t >= cnt || p >= ns) return;
            const int m = m0 + t;
            const int b = m / nI, i = m - b * nI;
            if (i + w > lengths[b]) return;
            const int* ce_seq = ce + (size_t)b * C;
            const int us = j0 + p + 1;
            e0 = ce_seq[cell_off(us, L) + i];
            e1 = ce_seq[cell_off(w - us, L) + i + us];
            e2 = item_e[t];
            e3 = gexp[t];
        };
        int e0n, e1n, e2n, e3n;
        fetch_fac(blockIdx.x, e0n, e1n, e2n, e3n);
        for (int t = blockIdx.x; t < cnt; t += gridDim.x) {
            const float f_cur = (e0n == INT_MIN) ? 0.f : ldexpf(1.0f, min(max(e0n + e1n - e2n + e3n, -126), 126));
            fetch_fac(t + gridDim.x, e0n, e1n, e2n, e3n);
            const int m = m0 + t;
            const int b = m / nI, i = m - b * nI, k = i + w;
            if (k > lengths[b]) continue;
            float* facb = fac + (grp * 2 + fbuf) * 128;
            facb[p] = f_cur;
            ptx::named_bar_sync(2 + grp, 128);
            const int rowL = (int)(((size_t)b * L + i) * Lp + (i + 1) + j0);
            for (int u = grp; u < units; u += 2) {
                const int kind = u / halves, h = u - kind * halves;
                const uint32_t slot = (nunit0 + u) & 3, slot_phase = ((nunit0 + u) >> 2) & 1;
                ptx::mbar_wait(&afull[slot], slot_phase);
                ptx::tc_fence_after();
                const int xl = q * 32 + lane;                    // row of the 128-wide tile
                const uint32_t taddr = tmem_base + slot * 128 + ((uint32_t)(q * 32) << 16);
                for (int c0 = 0; c0 < ns; c0 += 32) {
                    float v[32];
                    ptx::tmem_ld32(taddr + c0, v);
                    float* tile = reinterpret_cast<float*>(mybuf + (nred & 1) * 16384);
                    if (issuer) ptx::bulk_wait_read<1>();        // the reduce that used this tile two chunks ago has read it
                    ptx::named_bar_sync(2 + grp, 128);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tile[j * 128 + xl] = v[j] * facb[c0 + j];
                    ptx::fence_proxy_async_smem();
                    ptx::named_bar_sync(2 + grp, 128);
                    if (issuer) {
                        const bool last = (c0 + 32 >= ns);          // the tail box covers only the rows that hold splits
                        // left children (i, j): consecutive rows of start i.  Right children (j, k): row `end = k` of the
                        // consecutive starts j -- both land in ONE chart, so the two contributions a cell receives per
                        // diagonal meet in L2 (they are issued within a few microseconds of each other)
                        if (kind == 0) ptx::tma_reduce_add_2d(last ? &tmALt : &tmAL, tile, h * 128, rowL + c0, red_pol);
                        else ptx::tma_reduce_add_3d(last ? &tmARt : &tmAR, tile, h * 128, k, b * L + (i + 1) + j0 + c0, red_pol);
                        ptx::bulk_commit();
                    }
                    ++nred;
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(&aempty[slot]);
            }
            nunit0 += units;
            fbuf ^= 1;
        }
        if (issuer) ptx::bulk_wait_all();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// =======================================================================================================================
// K3: gY[m][(b,c)] = sum_a Ghat[m][a] * Wk[(b,c)][a]  -- A-resident variant
// =======================================================================================================================
// Output-heavy GEMM (K = N only): a CTA keeps the upstream-gradient rows of 256 spans resident in shared memory
// (two 128-row M tiles) and streams 128-row tiles of Wk past them, so every Wk tile feeds 256 spans (halves the
// L2 -> SM operand traffic that bounded the generic kernel) and the gradient rows are loaded once per work unit.
//   warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM owner, warps 4-7 epilogue (TMEM -> fp16 -> smem -> TMA store)
static constexpr int kGYStages = 4;

__global__ void __launch_bounds__(256, 1)
k_tc_gy(const __grid_constant__ CUtensorMap tmG,     // Ghat [cnt rows][N], box {64, 128}
        const __grid_constant__ CUtensorMap tmW,     // Wk [N*N rows][N], box {64, 128}
        const __grid_constant__ CUtensorMap tmOut,   // gY [cnt rows][N*N], box {64, 128}
        int N, int supertiles, int ranges) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int kblocks = N / 64;
    const int ntiles_total = N * N / 128;
    const int ntiles_per_range = ntiles_total / ranges;
    uint8_t* sAres = smem;                                        // [2 mtiles][kblocks] x 16 KB
    uint8_t* sBring = sAres + (size_t)2 * kblocks * 16384;        // [stages] x 16 KB
    uint8_t* ebuf = sBring + kGYStages * 16384;                   // 2 x 16 KB staging tiles
    uint64_t* full = (uint64_t*)(ebuf + 32768);
    uint64_t* empty = full + kGYStages;
    uint64_t* tfull = empty + kGYStages;      // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* a_full = tempty + 2;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(a_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGYStages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull[s], 1);
            ptx::mbar_init(&tempty[s], 128);
        }
        ptx::mbar_init(a_full, 1);
        ptx::mbar_init(a_empty, 1);
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tmG);
        ptx::tma_prefetch_desc(&tmW);
        ptx::tma_prefetch_desc(&tmOut);
    }
    if (warp == 2) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_sync();                        // nothing above reads or writes global memory
    const int units = supertiles * ranges;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int st = unit / ranges, r = unit - st * ranges;
                ptx::mbar_wait(a_empty, aphase ^ 1);
                ptx::mbar_expect_tx(a_full, (uint32_t)2 * kblocks * 16384);
                for (int mt = 0; mt < 2; ++mt)
                    for (int kb = 0; kb < kblocks; ++kb)
                        ptx::tma_load_2d(sAres + (size_t)(mt * kblocks + kb) * 16384, &tmG, a_full, kb * 64, st * 256 + mt * 128);
                aphase ^= 1;
                for (int nt = r * ntiles_per_range; nt < (r + 1) * ntiles_per_range; ++nt)
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        ptx::mbar_expect_tx(&full[stage], 16384);
                        ptx::tma_load_2d(sBring + stage * 16384, &tmW, &full[stage], kb * 64, nt * 128);
                        if (++stage == kGYStages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = ptx::make_idesc_f16(128, 128, false, false, 0);
            int stage = 0, buf = 0;
            uint32_t phase = 0, aphase = 0, bphase = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int st = unit / ranges, r = unit - st * ranges;
                (void)st;
                ptx::mbar_wait(a_full, aphase);
                aphase ^= 1;
                ptx::tc_fence_after();
                const uint32_t aA = ptx::smem_u32(sAres);
                for (int nt = r * ntiles_per_range; nt < (r + 1) * ntiles_per_range; ++nt) {
                    ptx::mbar_wait(&tempty[buf], bphase ^ 1);
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&full[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sB = ptx::smem_u32(sBring + stage * 16384);
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint32_t sA = aA + (uint32_t)(mt * kblocks + kb) * 16384;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::mma_f16_ss(tmem_base + buf * 256 + mt * 128, ptx::make_smem_desc(sA + k * 32, 16, 1024),
                                                ptx::make_smem_desc(sB + k * 32, 16, 1024), idesc, (uint32_t)((kb | k) != 0));
                        }
                        ptx::mma_commit(&empty[stage]);
                        if (++stage == kGYStages) { stage = 0; phase ^= 1; }
                    }
                    ptx::mma_commit(&tfull[buf]);
                    if (++buf == 2) { buf = 0; bphase ^= 1; }
                }
                ptx::mma_commit(a_empty);          // the resident A tiles may be replaced once these MMAs retire
            }
        }
    } else if (warp >= 4) {
        // NOTE (measured): this kernel is shared-memory-bandwidth bound (MMA operand reads + TMA fills + staging
        // ~250 B/cycle against 128 B/cycle); a second epilogue group made it slower, see DESIGN.md section 4.
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool issuer = (threadIdx.x == 128);
        int buf = 0;
        uint32_t bphase = 0, nstore = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int st = unit / ranges, r = unit - st * ranges;
            for (int nt = r * ntiles_per_range; nt < (r + 1) * ntiles_per_range; ++nt) {
                ptx::mbar_wait(&tfull[buf], bphase);
                ptx::tc_fence_after();
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t taddr = tmem_base + buf * 256 + mt * 128 + ((uint32_t)(q * 32) << 16);
                    for (int c0 = 0; c0 < 128; c0 += 64) {
                        float v[64];
                        ptx::tmem_ld32(taddr + c0, v);
                        ptx::tmem_ld32(taddr + c0 + 32, v + 32);
                        ptx::tmem_ld_wait();
                        ptx::stage_store_tile(ebuf, &tmOut, nt * 128 + c0, st * 256 + mt * 128, row, v, nstore, 3, issuer, true);
                        ++nstore;
                    }
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tempty[buf]);
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
        }
        if (issuer) ptx::bulk_wait_all();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// K3, CTA-pair variant (cta_group::2): the pair keeps the gradient rows of 256 spans resident (128 per CTA) and every
// CTA stages only HALF of each 256-row Wk tile, which halves both the TMA fill and the MMA operand reads per SM --
// the 1-CTA kernel above is shared-memory-bandwidth bound (~250 B/cycle needed, 128 available).
template <int NBUF, int STG>      // staging tiles per epilogue group and Wk ring stages: whatever fits beside the resident
                                  // gradient rows (N <= 256: 2 and 6, N = 512: 1 and 4)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
k_tc_gy2(const __grid_constant__ CUtensorMap tmG,     // Ghat [cnt rows][N], box {64, 128}
         const __grid_constant__ CUtensorMap tmW,     // Wk [N*N rows][N], box {64, 128}
         const __grid_constant__ CUtensorMap tmOut,   // gY [cnt rows][N*N], box {64, 128}
         int N, int supertiles, int ranges, uint64_t out_pol, uint64_t w_pol) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int kblocks = N / 64;
    const int ntiles_total = N * N / 256;
    // range r of a supertile covers Wk tiles [r * ntiles / ranges, (r + 1) * ntiles / ranges): any number of ranges,
    // sizes differ by at most one tile
    auto range_lo = [&](int r) { return (int)(((long long)r * ntiles_total) / ranges); };
    uint8_t* sAres = smem;                                        // [kblocks] x 16 KB (this CTA's 128 rows)
    uint8_t* sBring = sAres + (size_t)kblocks * 16384;            // [stages] x 16 KB (this CTA's half of a Wk tile)
    uint8_t* ebuf = sBring + STG * 16384;                   // [2 groups][NBUF] x 16 KB staging tiles
    uint64_t* full = (uint64_t*)(ebuf + 2 * NBUF * 16384);
    uint64_t* empty = full + STG;
    uint64_t* tfull = empty + STG;      // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* a_full = tempty + 2;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(a_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = (rank == 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STG; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull[s], 1);
            ptx::mbar_init(&tempty[s], 512);
        }
        ptx::mbar_init(a_full, 1);
        ptx::mbar_init(a_empty, 1);
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tmG);
        ptx::tma_prefetch_desc(&tmW);
        ptx::tma_prefetch_desc(&tmOut);
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_sync();                        // nothing above reads or writes global memory
    const int units = supertiles * ranges;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            for (int unit = pair; unit < units; unit += npairs) {
                const int st = unit / ranges, r = unit - st * ranges;
                ptx::mbar_wait(a_empty, aphase ^ 1);
                if (leader) ptx::mbar_expect_tx(a_full, (uint32_t)2 * kblocks * 16384);
                for (int kb = 0; kb < kblocks; ++kb)
                    ptx::tma_load_2d_pair(sAres + (size_t)kb * 16384, &tmG, a_full, kb * 64, st * 256 + (int)rank * 128);
                aphase ^= 1;
                for (int nt = range_lo(r); nt < range_lo(r + 1); ++nt)
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        if (leader) ptx::mbar_expect_tx(&full[stage], 2 * 16384);
                        ptx::tma_load_2d_pair_hint(sBring + stage * 16384, &tmW, &full[stage], kb * 64, nt * 256 + (int)rank * 128, w_pol);
                        if (++stage == STG) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = ptx::make_idesc_f16(256, 256, false, false, 0);
            int stage = 0, buf = 0;
            uint32_t phase = 0, aphase = 0, bphase = 0;
            for (int unit = pair; unit < units; unit += npairs) {
                const int st = unit / ranges, r = unit - st * ranges;
                (void)st;
                ptx::mbar_wait(a_full, aphase);
                aphase ^= 1;
                ptx::tc_fence_after();
                const uint32_t aA = ptx::smem_u32(sAres);
                for (int nt = range_lo(r); nt < range_lo(r + 1); ++nt) {
                    ptx::mbar_wait(&tempty[buf], bphase ^ 1);
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&full[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sB = ptx::smem_u32(sBring + stage * 16384);
                        const uint32_t sA = aA + (uint32_t)kb * 16384;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::mma_f16_ss_pair(tmem_base + buf * 256, ptx::make_smem_desc(sA + k * 32, 16, 1024),
                                                 ptx::make_smem_desc(sB + k * 32, 16, 1024), idesc, (uint32_t)((kb | k) != 0));
                        ptx::mma_commit_pair(&empty[stage]);
                        if (++stage == STG) { stage = 0; phase ^= 1; }
                    }
                    ptx::mma_commit_pair(&tfull[buf]);
                    if (++buf == 2) { buf = 0; bphase ^= 1; }
                }
                ptx::mma_commit_pair(a_empty);
            }
        }
    } else if (warp >= 4) {
        // two epilogue groups per CTA: warps 4-7 drain columns 0-127, warps 8-11 columns 128-255 of every accumulator
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const bool issuer = ((threadIdx.x - 128) & 127) == 0;
        uint8_t* mybuf = ebuf + grp * (NBUF * 16384);
        int buf = 0;
        uint32_t bphase = 0, nstore = 0;
        for (int unit = pair; unit < units; unit += npairs) {
            const int st = unit / ranges, r = unit - st * ranges;
            for (int nt = range_lo(r); nt < range_lo(r + 1); ++nt) {
                ptx::mbar_wait(&tfull[buf], bphase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + buf * 256 + grp * 128 + ((uint32_t)(q * 32) << 16);
                float v[128];
#pragma unroll
                for (int c0 = 0; c0 < 128; c0 += 32) ptx::tmem_ld32(taddr + c0, v + c0);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive_leader(&tempty[buf]);           // the accumulator is in registers: release it early
#pragma unroll
                for (int c0 = 0; c0 < 128; c0 += 64) {
                    ptx::stage_store_tile<NBUF>(mybuf, &tmOut, nt * 256 + grp * 128 + c0, st * 256 + (int)rank * 128, row, v + c0, nstore, 3 + grp, issuer, true, out_pol);
                    ++nstore;
                }
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
        }
        if (issuer) ptx::bulk_wait_all();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, 512);
}

// K3, CTA-pair variant with the gradient rows in TENSOR MEMORY (tcgen05.mma with a TMEM A operand).  The kernels above
// are shared-memory-bandwidth bound at N = 512 (MMA operand reads of A and B + TMA fills + store staging = 128 B/clk per
// SM); here A costs nothing, the Wk ring needs 8 KB per k-block and CTA, and the accumulators are 128 columns wide:
//   TMEM columns [0, N/2)            this CTA's 128 gradient rows, two fp16 per column (written by tcgen05.st)
//   TMEM columns [N/2 + 128 buf ...) two accumulators of 128 (b,c) columns
//   warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM owner, warps 4-7 load A and drain columns 0-63,
//   warps 8-11 drain columns 64-127 of every accumulator
static constexpr int kGY3Stages = 12;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
k_tc_gy3(const __grid_constant__ CUtensorMap tmW,     // Wk [N*N rows][N], box {64, 64}
         const __grid_constant__ CUtensorMap tmOut,   // gY [cnt rows][N*N], box {64, 128}
         const __half* __restrict__ ghat,             // [>= cnt rows][N]
         int cnt, int N, int supertiles, int ranges) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int kblocks = N / 64;
    const int ntiles_total = N * N / 128;
    const int ntiles_per_range = ntiles_total / ranges;
    uint8_t* sBring = smem;                                       // [stages] x 8 KB (this CTA's 64 rows of a Wk tile)
    uint8_t* ebuf = sBring + kGY3Stages * 8192;                   // [2 groups][2] x 16 KB staging tiles
    uint64_t* full = (uint64_t*)(ebuf + 65536);
    uint64_t* empty = full + kGY3Stages;
    uint64_t* tfull = empty + kGY3Stages;     // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* a_full = tempty + 2;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(a_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = (rank == 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGY3Stages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull[s], 1);
            ptx::mbar_init(&tempty[s], 512);
        }
        ptx::mbar_init(a_full, 256);
        ptx::mbar_init(a_empty, 1);
        ptx::fence_barrier_init();
        ptx::tma_prefetch_desc(&tmW);
        ptx::tma_prefetch_desc(&tmOut);
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_sync();                        // nothing above reads or writes global memory
    const uint32_t acc_base = tmem_base + (uint32_t)(N / 2);
    const int units = supertiles * ranges;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = pair; unit < units; unit += npairs) {
                const int r = unit % ranges;
                for (int nt = r * ntiles_per_range; nt < (r + 1) * ntiles_per_range; ++nt)
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        if (leader) ptx::mbar_expect_tx(&full[stage], 2 * 8192);
                        ptx::tma_load_2d_pair(sBring + stage * 8192, &tmW, &full[stage], kb * 64, nt * 128 + (int)rank * 64);
                        if (++stage == kGY3Stages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = ptx::make_idesc_f16(256, 128, false, false, 0);
            int stage = 0, buf = 0;
            uint32_t phase = 0, aphase = 0, bphase = 0;
            for (int unit = pair; unit < units; unit += npairs) {
                const int r = unit % ranges;
                ptx::mbar_wait(a_full, aphase);
                aphase ^= 1;
                ptx::tc_fence_after();
                for (int nt = r * ntiles_per_range; nt < (r + 1) * ntiles_per_range; ++nt) {
                    ptx::mbar_wait(&tempty[buf], bphase ^ 1);
                    ptx::tc_fence_after();
                    for (int kb = 0; kb < kblocks; ++kb) {
                        ptx::mbar_wait(&full[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t sB = ptx::smem_u32(sBring + stage * 8192);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            ptx::mma_f16_ts_pair(acc_base + buf * 128, tmem_base + (uint32_t)(kb * 32 + k * 8),
                                                 ptx::make_smem_desc(sB + k * 32, 16, 1024), idesc, (uint32_t)((kb | k) != 0));
                        ptx::mma_commit_pair(&empty[stage]);
                        if (++stage == kGY3Stages) { stage = 0; phase ^= 1; }
                    }
                    ptx::mma_commit_pair(&tfull[buf]);
                    if (++buf == 2) { buf = 0; bphase ^= 1; }
                }
                ptx::mma_commit_pair(a_empty);         // the gradient rows may be replaced once these MMAs retire
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const bool issuer = ((threadIdx.x - 128) & 127) == 0;
        uint8_t* mybuf = ebuf + grp * 32768;
        int buf = 0;
        uint32_t bphase = 0, nstore = 0, aphase = 0;
        for (int unit = pair; unit < units; unit += npairs) {
            const int st = unit / ranges, r = unit - st * ranges;
            if (grp == 0) {
                // this CTA's 128 gradient rows -> TMEM (one row per thread, 32 columns = 64 halves per store)
                ptx::mbar_wait(a_empty, aphase ^ 1);
                aphase ^= 1;
                ptx::tc_fence_after();
                const int grow = st * 256 + (int)rank * 128 + row;
                const uint4* src = reinterpret_cast<const uint4*>(ghat + (size_t)grow * N);
                for (int c0 = 0; c0 < N / 2; c0 += 32) {
                    uint32_t v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint4 x = make_uint4(0, 0, 0, 0);
                        if (grow < cnt) x = __ldg(src + (c0 >> 2) + j);
                        v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
                    }
                    ptx::tmem_st32u(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive_leader(a_full);
            }
            for (int nt = r * ntiles_per_range; nt < (r + 1) * ntiles_per_range; ++nt) {
                ptx::mbar_wait(&tfull[buf], bphase);
                ptx::tc_fence_after();
                const uint32_t taddr = acc_base + buf * 128 + grp * 64 + ((uint32_t)(q * 32) << 16);
                float v[64];
                ptx::tmem_ld32(taddr, v);
                ptx::tmem_ld32(taddr + 32, v + 32);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive_leader(&tempty[buf]);           // the accumulator is in registers: release it early
                ptx::stage_store_tile(mybuf, &tmOut, nt * 128 + grp * 64, st * 256 + (int)rank * 128, row, v, nstore, 3 + grp, issuer, true);
                ++nstore;
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
        }
        if (issuer) ptx::bulk_wait_all();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, 512);
}

// =======================================================================================================================
// element-wise helpers of the tensor-core path
// =======================================================================================================================
// per-parent exponent of max_{b,c} W[b][c][a]
// (and, in wmax_bits[N + a], the number of nonzero rules of parent a)
__global__ void k_w_colmax(int N, const float* __restrict__ W, int* __restrict__ wmax_bits) {
    // grid over (b,c) rows in slabs; thread a
    size_t rows = (size_t)N * N;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        float mx = 0.f;
        int nnz = 0;
        for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
            const float x = W[r * N + a];
            mx = fmaxf(mx, x);
            nnz += (x != 0.f);
        }
        atomicMax(&wmax_bits[a], __float_as_int(mx));     // non-negative floats order like ints
        atomicAdd(&wmax_bits[N + a], nnz);
    }
}
// tcgen05.mma adds into its fp32 accumulator with TRUNCATION: a chain of non-negative products comes out low by
// kTruncPerKBlock per k-block of 64 (measured: K = 1024 / 4096 / 8192 / 65536 -> 5.5e-6 / 2.4e-5 / 4.8e-5 / 3.9e-4,
// tools/experiments/exp_bias.py; linear in the chain length, independent of the data's scale).  The inside pass feeds one
// such GEMM into the next 127 times.  log Z only moves by ~3e-3, but posteriors of a node at depth d of a derivation
// come out HIGH by d times the per-level bias (the outside recursion has K = 256 chains and no such bias to cancel
// it): +4e-4 on the expected counts at L = 128, i.e. half of the 1e-3 budget, as a systematic error.  The rule GEMM's
// epilogue therefore multiplies parent a by 1 + kTruncPerKBlock * (k-blocks per accumulation chain) * density[a], where
// density[a] = fraction of nonzero rules of parent a (adding an exact zero does not truncate).  RBN_TC_KAPPA overrides
// the constant (0 switches the compensation off).
static constexpr float kTruncPerKBlock = 3.7e-7f;
__global__ void k_w_exp(int N, const int* __restrict__ wmax_bits, int* __restrict__ wexp, float* __restrict__ wscale,
                        float* __restrict__ wdens, float kc_full) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    float mx = __int_as_float(wmax_bits[a]);
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    wexp[a] = e;
    const float dens = (float)wmax_bits[N + a] / ((float)N * (float)N);
    wdens[a] = dens;
    wscale[a] = ldexpf(1.0f, e);
    wscale[N + a] = ldexpf(1.0f, e) * (1.0f + kc_full * dens);     // whole tiles: chains of kSegBlocks k-blocks
}
// Wk[(b,c)][a] = W * 2^-wexp[a]  and its transpose Wt[a][(b,c)]  (32x32 tiles through shared memory)
__global__ void k_w_copies(int N, const float* __restrict__ W, const int* __restrict__ wexp, __half* __restrict__ wk,
                           __half* __restrict__ wt) {
    __shared__ __half tile[32][33];
    size_t rows = (size_t)N * N;
    size_t r0 = (size_t)blockIdx.x * 32;
    int a0 = blockIdx.y * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        size_t r = r0 + rr;
        int a = a0 + threadIdx.x;
        __half h = __float2half_rn(ldexpf(W[r * N + a], -wexp[a]));
        wk[r * N + a] = h;
        tile[rr][threadIdx.x] = h;
    }
    __syncthreads();
    for (int aa = threadIdx.y; aa < 32; aa += blockDim.y) wt[(size_t)(a0 + aa) * rows + r0 + threadIdx.x] = tile[threadIdx.x][aa];
}
// fp16 operand copies of the width-1 cells
__global__ void k_copy_width1(int N, int L, int Lp, const int* __restrict__ lengths, const float* __restrict__ cv,
                              __half* __restrict__ lc, __half* __restrict__ rc) {
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    const float* src = cv + ((size_t)b * cells_per_seq(L) + i) * N;
    __half* l = lc + ((size_t)((size_t)b * L + i) * Lp + (i + 1)) * N;
    __half* r = rc + ((size_t)((size_t)b * (L + 1) + (i + 1)) * Lp + i) * N;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        __half h = __float2half_rn(src[a]);
        l[a] = h;
        r[a] = h;
    }
}
// max-shift epilogue of the K-split tail tiles of the rule GEMM: acc[t][a] (fp32 sums) -> chart, exponents, fp16 copies.
// Splits per split-sum / child-gradient launch (two K slabs of 64 chart rows; 128 fix-up threads, 128 factors).
static constexpr int kMaxWindow = 128;

// Y[t] <- Y[t] 2^(E[t] - E') + Yp[t] 2^(Ep[t] - E'),  E[t] <- E' = max(E[t], Ep[t]): folds the partial split-sum of one more
// window of splits (spans wider than 129 tokens) into the span's split-sum.  One CTA per span, 16-byte accesses.
__global__ void __launch_bounds__(256)
k_y_combine(int N, int L, int w, int m0, const int* __restrict__ lengths, __half* __restrict__ Y, int* __restrict__ E,
            const __half* __restrict__ Yp, const int* __restrict__ Ep) {
    const int t = blockIdx.x;
    const int nI = L - w + 1;
    const int m = m0 + t;
    const int b = m / nI, i = m - b * nI;
    if (i + w > lengths[b]) return;                 // span beyond its sequence: both parts are zeros, no exponent was written
    const int ea = E[t], eb = Ep[t];
    const int e = max(ea, eb);
    const __half2 sa = __float2half2_rn(ldexpf(1.0f, max(ea - e, -60))), sb = __float2half2_rn(ldexpf(1.0f, max(eb - e, -60)));
    uint4* y = reinterpret_cast<uint4*>(Y + (size_t)t * N * N);
    const uint4* yp = reinterpret_cast<const uint4*>(Yp + (size_t)t * N * N);
    for (int idx = threadIdx.x; idx < N * N / 8; idx += blockDim.x) {
        uint4 a = y[idx], p = yp[idx];
        __half2* ha = reinterpret_cast<__half2*>(&a);
        const __half2* hp = reinterpret_cast<const __half2*>(&p);
#pragma unroll
        for (int q = 0; q < 4; ++q) ha[q] = __hfma2(ha[q], sa, __hmul2(hp[q], sb));
        y[idx] = a;
    }
    __syncthreads();                                // every thread has read E[t]
    if (threadIdx.x == 0) E[t] = e;
}

// One WARP per span (8 spans per CTA, N / 32 values per lane, float4 accesses): no block-level synchronisation and an
// eighth of the CTAs of the former one-CTA-per-span version (20 -> ~8 us per diagonal of 16 k spans; 508 launches per step).
static constexpr int kFinWarps = 8;
__global__ void __launch_bounds__(kFinWarps * 32)
k_rule_finalize(int N, int L, int Lp, int w, int m0, int t_begin, int cnt, const int* __restrict__ lengths,
                const float* __restrict__ wscale, const int* __restrict__ item_e,
                const float* __restrict__ acc, float* __restrict__ cv, int* __restrict__ ce,
                __half* __restrict__ lc, __half* __restrict__ rc, const float* __restrict__ wdens, float kc) {
    ptx::grid_dep_sync();
    const int lane = threadIdx.x & 31;
    const int t = t_begin + blockIdx.x * kFinWarps + (threadIdx.x >> 5);
    if (t >= cnt) return;
    const int nI = L - w + 1;
    const int m = m0 + t;
    const int b = m / nI, i = m - b * nI, k = i + w;
    if (k > lengths[b]) return;
    constexpr int kMaxQuads = 4;                     // N <= 512: N / 128 float4 per lane
    const int quads = N >> 7;                        // (N = 64, 192: the tail quad is handled by the lanes that own it)
    float4 x[kMaxQuads];
    float mx = 0.f;
    const float* src = acc + (size_t)t * N;
#pragma unroll
    for (int q = 0; q < kMaxQuads; ++q) {
        const int a = (q * 32 + lane) * 4;
        x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q <= quads && a < N) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(src + a));      // written by red.global.add: read at L2
            const float4 sc = *reinterpret_cast<const float4*>(wscale + a);
            const float4 dn = *reinterpret_cast<const float4*>(wdens + a);
            // kc: truncation compensation of this launch's accumulation chains
            x[q] = make_float4(v.x * sc.x * (1.0f + kc * dn.x), v.y * sc.y * (1.0f + kc * dn.y),
                               v.z * sc.z * (1.0f + kc * dn.z), v.w * sc.w * (1.0f + kc * dn.w));
            mx = fmaxf(mx, fmaxf(fmaxf(x[q].x, x[q].y), fmaxf(x[q].z, x[q].w)));
        }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e_new = 0;
    if (mx > 0.f) frexpf(mx, &e_new);
    const size_t cell = (size_t)b * cells_per_seq(L) + cell_off(w, L) + i;
    float* cvp = cv + cell * N;
    __half* lcp = lc + ((size_t)((size_t)b * L + i) * Lp + k) * N;
    __half* rcp = rc + ((size_t)((size_t)b * (L + 1) + k) * Lp + i) * N;
#pragma unroll
    for (int q = 0; q < kMaxQuads; ++q) {
        const int a = (q * 32 + lane) * 4;
        if (q <= quads && a < N) {
            const float4 v = make_float4(ldexpf(x[q].x, -e_new), ldexpf(x[q].y, -e_new), ldexpf(x[q].z, -e_new), ldexpf(x[q].w, -e_new));
            *reinterpret_cast<float4*>(cvp + a) = v;
            const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
            uint2 hp;
            hp.x = *reinterpret_cast<const uint32_t*>(&h0);
            hp.y = *reinterpret_cast<const uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(lcp + a) = hp;
            *reinterpret_cast<uint2*>(rcp + a) = hp;
        }
    }
    if (lane == 0) ce[cell] = item_e[t] + e_new;
}

// root seed into the start-major outside chart (and the prior gradient)
__global__ void k_tc_outside_seed(int N, int L, int Lp, const float* __restrict__ prior, const int* __restrict__ lengths,
                                  const float* __restrict__ cv, const float* __restrict__ ws_zroot,
                                  const float* __restrict__ coef, float* __restrict__ alphaL, float* __restrict__ gPrior) {
    int b = blockIdx.x;
    int n = lengths[b];
    size_t cell = (size_t)b * cells_per_seq(L) + cell_off(n, L);
    size_t row = ((size_t)b * L + 0) * Lp + n;
    float z = ws_zroot[b];
    float f = coef[b] / (z > 0.f ? z : 1.0f);
    if (f == 0.f) return;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        alphaL[row * N + a] = f * prior[a];
        float m = cv[cell * N + a];
        if (m != 0.f) atomicAdd(&gPrior[a], f * m);
    }
}
__global__ void k_tc_emission_bwd(int N, int L, int Lp, int nt, const int* __restrict__ tokens,
                                  const int* __restrict__ lengths, const int* __restrict__ ce,
                                  const float* __restrict__ alphaL, float* __restrict__ gT) {
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    size_t cell = (size_t)b * cells_per_seq(L) + i;
    size_t rl = ((size_t)b * L + i) * Lp + (i + 1);
    int e = ce[cell];
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        float g = ldexpf(alphaL[rl * N + a], -e);
        if (g == 0.f) continue;
        for (int t = 0; t < nt; ++t) {
            int y = tokens[((size_t)b * L + i) * nt + t];
            if (y >= 0) atomicAdd(&gT[(size_t)y * N + a], g);
        }
    }
}
// upstream gradient of one chunk:  G = alpha * 2^-e_new ;  Ghat = G * 2^wexp[a] normalised per item (fp16, operand of the
// gY GEMM) ; G itself in fp32 plus the running max |G| of the launch group (k_gabs_convert makes the count GEMM's operand)
// One WARP per span (8 spans per CTA), N / 32 values per lane: no block-level synchronisation, an eighth of the CTAs.
static constexpr int kPrepWarps = 8;
// Lane `lane` owns the N / 32 CONSECUTIVE nonterminals [lane * per, (lane + 1) * per): its loads and stores are 8- / 16-byte
// vectors (per is 2, 4, 6, 8 or 16), an eighth of the memory instructions of an interleaved mapping.
__global__ void __launch_bounds__(kPrepWarps * 32)
k_prep_g(int N, int L, int Lp, int w, int m0, int cnt, int rows, int gabs_rows, const int* __restrict__ lengths, const int* __restrict__ ce,
         const int* __restrict__ item_e, const int* __restrict__ wexp, const float* __restrict__ alpha,
         __half* __restrict__ ghat, float* __restrict__ g32, int* __restrict__ gmax_bits, int* __restrict__ gexp) {
    ptx::grid_dep_sync();
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * kPrepWarps + (threadIdx.x >> 5);
    if (t >= rows) return;                          // rows = cnt rounded up to 64: the pad rows are written as zeros
    // gabs_rows: rows of `g32` this launch owns (= rows for a slot buffer; = cnt when the rows are packed behind the
    // previous unit's in the deferred-count array, where the next unit's rows follow immediately)
    const int nI = L - w + 1;
    const int m = m0 + t;
    const int b = m / nI, i = m - b * nI;
    const bool valid = (t < cnt) && (i + w <= lengths[t < cnt ? b : 0]);
    constexpr int kMaxPerLane = 16;                 // N <= 512
    float g[kMaxPerLane], gs[kMaxPerLane];
    float mx = 0.f, mxg = 0.f;
    const int per = N >> 5;
    const int a0 = lane * per;
    if (valid) {
        const size_t cell = (size_t)b * cells_per_seq(L) + cell_off(w, L) + i;
        const float2* row = reinterpret_cast<const float2*>(alpha + (((size_t)b * L + i) * Lp + (i + w)) * N + a0);
        const int2* we = reinterpret_cast<const int2*>(wexp + a0);
        float2 rv[kMaxPerLane / 2];
#pragma unroll
        for (int q = 0; q < kMaxPerLane / 2; ++q)
            if (2 * q < per) rv[q] = row[q];        // issued before the exponents arrive (independent loads)
        const int e_new = ce[cell] - item_e[t];
#pragma unroll
        for (int q = 0; q < kMaxPerLane / 2; ++q) {
            if (2 * q < per) {
                const int2 ex = we[q];
                g[2 * q] = ldexpf(rv[q].x, -e_new);
                g[2 * q + 1] = ldexpf(rv[q].y, -e_new);
                gs[2 * q] = ldexpf(g[2 * q], ex.x);
                gs[2 * q + 1] = ldexpf(g[2 * q + 1], ex.y);
                mx = fmaxf(mx, fmaxf(fabsf(gs[2 * q]), fabsf(gs[2 * q + 1])));
                mxg = fmaxf(mxg, fmaxf(fabsf(g[2 * q]), fabsf(g[2 * q + 1])));
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < kMaxPerLane; ++q) g[q] = gs[q] = 0.f;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mxg = fmaxf(mxg, __shfl_xor_sync(0xffffffffu, mxg, o));
    }
    int eg = 0;
    if (mx > 0.f) frexpf(mx, &eg);
    if (mxg > 0.f && mxg <= 3.0e38f) {                // non-negative floats order like ints
        const int bits = __float_as_int(mxg);
        if (lane == 0 && bits > *(volatile int*)gmax_bits) atomicMax(gmax_bits, bits);
    }
    __half2* hrow = reinterpret_cast<__half2*>(ghat + (size_t)t * N + a0);
    float2* grow = reinterpret_cast<float2*>(g32 + (size_t)t * N + a0);
#pragma unroll
    for (int q = 0; q < kMaxPerLane / 2; ++q) {
        if (2 * q < per) {
            hrow[q] = __floats2half2_rn(ldexpf(gs[2 * q], -eg), ldexpf(gs[2 * q + 1], -eg));
            if (t < gabs_rows) grow[q] = make_float2(g[2 * q], g[2 * q + 1]);     // (cached spans: packed rows, no pad rows)
        }
    }
    if (lane == 0) gexp[t] = eg;
}

// Operand of the count GEMM: G rows in fp16 under ONE power-of-two scale for the whole launch (the GEMM sums over spans,
// so a per-span scale cannot be factored out).  The scale puts the largest |G| of the launch at 2^14: every entry down
// to 2^-28 of it stays a NORMAL fp16 number.  (Round 1 used a fixed 2^-8, which left the bulk of the posterior mass --
// entries of 1e-5 .. 1e-2 -- in the subnormal range, which the tensor core does not honour: -6e-4 on the counts at
// N = 512.)  `scale` receives the inverse factor for the GEMM's epilogue.
__global__ void k_gabs_convert(long long n, const float* __restrict__ g32, __half* __restrict__ g16, const int* __restrict__ gmax_bits,
                               float* __restrict__ scale) {
    const float mx = __int_as_float(*gmax_bits);
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    const int shift = 14 - e;                       // max * 2^shift in [2^13, 2^14)
    if (blockIdx.x == 0 && threadIdx.x == 0) *scale = ldexpf(1.0f, -shift);
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(g32 + i);
        __half2 lo = __floats2half2_rn(ldexpf(v.x, shift), ldexpf(v.y, shift));
        __half2 hi = __floats2half2_rn(ldexpf(v.z, shift), ldexpf(v.w, shift));
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(g16 + i) = o;
    }
}

// =======================================================================================================================
// drivers
// =======================================================================================================================
// default: bit 4 (rule / count GEMM operand streams: Y evict_first, W evict_last).  Measured (round 2, ncu DRAM bytes per
// launch): without it the 2 GB of Y streaming through L2 per launch keep evicting the 33 MB of W, which is then re-read
// from HBM ~8 times per launch (164 KB read per span against 131 KB of Y); no change in kernel time (these GEMMs are
// tensor-bound), less HBM traffic and power.  Bits 1, 2, 8 changed nothing measurable and stay off.
static constexpr int kDefaultHints = 4;
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
int device_sm_count() { return sm_count(); }
const TcTune& tc_tune() {
    static TcTune t;
    static std::once_flag once;
    std::call_once(once, [] {
        t.overlap = env_int("RBN_TC_OVERLAP", 0);
        if (t.overlap < 0 || t.overlap > 3) t.overlap = 0;
        t.groups = env_int("RBN_TC_GROUPS", 2);
        t.fwd_sms0 = env_int("RBN_TC_FWD_SMS0", 40) & ~1;
        t.bwd_sms0 = env_int("RBN_TC_BWD_SMS0", 74) & ~1;
        t.nslot = env_int("RBN_TC_SLOTS", 3);
        t.k3_stream = env_int("RBN_TC_K3_STREAM", 0) != 0;
        t.hints = env_int("RBN_TC_HINTS", kDefaultHints);
        t.pdl = env_int("RBN_PDL", 1) != 0;
        const char* ek = getenv("RBN_TC_KAPPA");
        t.kappa = (ek && *ek) ? (float)atof(ek) : kTruncPerKBlock;
        if (t.groups < 1) t.groups = 1;
        if (t.nslot < 2) t.nslot = 2;
        if (t.nslot > 8) t.nslot = 8;
    });
    return t;
}

// work unit of a pass: one chunk of spans of one diagonal and one group of sequences
struct Unit {
    int w, g, m0, cnt, slot;
    long long coff;      // first span of this unit in the split-sum cache, -1: not cached (Y lives in the slot buffer)
};

struct TcCtx {
    const RbnDesc* d;
    const WsLayout* ws;
    char* base;
    int N, L, B, Lp, chunk;
    float* cv;
    int* ce;
    float* alpha;    // start-major outside chart: root seed + left- and right-child contributions
    __half *lc, *rc, *wt, *wk;
    int *wexp, *wbits;
    float* wscale;     // [2][N]: 2^wexp[a], and the same times the truncation compensation of whole tiles
    float* wdens;      // [N] fraction of nonzero rules per parent
    float kappa;       // truncation bias per k-block of an accumulation chain (kTruncPerKBlock or RBN_TC_KAPPA)
    const int* lengths;
    // per-chunk buffers, `nslot` copies
    __half* y(int s) const { return (__half*)(base + ws->y) + (size_t)s * chunk * N * N; }
    __half* gy(int s) const { return (__half*)(base + ws->gy) + (size_t)s * chunk * N * N; }
    __half* ghat(int s) const { return (__half*)(base + ws->ghat) + (size_t)s * chunk * N; }
    __half* gabs(int s) const { return (__half*)(base + ws->gabs) + (size_t)s * chunk * N; }
    float* acc(int s) const { return (float*)(base + ws->acc) + (size_t)s * chunk * N; }
    int* item_e(int s) const { return (int*)(base + ws->item_e) + (size_t)s * chunk; }
    int* gexp(int s) const { return (int*)(base + ws->gexp) + (size_t)s * chunk; }
    // split-sum and span exponents E_m of a unit: in the cache when the unit has a place there, else in its slot
    __half* y_of(const Unit& u) const { return u.coff >= 0 ? (__half*)(base + ws->ycache) + (size_t)u.coff * N * N : y(u.slot); }
    // absolutely scaled upstream gradient rows of a unit: cached units keep them (packed like their Y) until the ONE count
    // GEMM over all cached spans at the end of the outside pass
    __half* gabs_of(const Unit& u) const { return u.coff >= 0 ? (__half*)(base + ws->gabs_all) + (size_t)u.coff * N : gabs(u.slot); }
    float* g32_of(const Unit& u) const {
        return u.coff >= 0 ? (float*)(base + ws->g32_all) + (size_t)u.coff * N : (float*)(base + ws->g32) + (size_t)u.slot * chunk * N;
    }
    // [0]: max |G| (float bits) of the kept spans, [1 + slot]: of the unit in that slot; gscale likewise (written by k_gabs_convert)
    int* gmax(int i) const { return (int*)(base + ws->gmax) + i; }
    float* gscale(int i) const { return (float*)(base + ws->gmax) + 16 + i; }
    int* e_of(const Unit& u) const { return u.coff >= 0 ? (int*)(base + ws->ycache_e) + (size_t)u.coff : item_e(u.slot); }
};

static void fill_ctx(TcCtx& c, const RbnDesc* d, const WsLayout& ws, char* base, const int* lengths) {
    c.d = d; c.ws = &ws; c.base = base; c.N = d->N; c.L = d->L; c.B = d->B; c.Lp = ws.Lp; c.chunk = ws.chunk;
    c.cv = (float*)(base + ws.chart_val); c.ce = (int*)(base + ws.chart_exp); c.alpha = (float*)(base + ws.alpha);
    c.lc = (__half*)(base + ws.lc); c.rc = (__half*)(base + ws.rc); c.wt = (__half*)(base + ws.wt);
    c.wk = (__half*)(base + ws.wk);
    c.wexp = (int*)(base + ws.wexp); c.wbits = (int*)(base + ws.wbits); c.wscale = (float*)(base + ws.wscale);
    c.wdens = (float*)(base + ws.wdens);
    c.kappa = tc_tune().kappa;
    c.lengths = lengths;
}

// ---- two launch streams with disjoint SM partitions ------------------------------------------------------------------------
// Stream 0 is the caller's; stream 1 and the events come from a process-wide pool (created on first use, never
// destroyed; a pass returns them as soon as its last launch is queued -- later records / launches on a recycled
// event or stream simply order after the earlier ones).
struct AuxPool {
    std::mutex mu;
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> events;
};
static AuxPool& aux_pool(int dev) {
    static AuxPool pools[16];
    return pools[dev & 15];
}

struct Sched {
    cudaStream_t st[2];
    int sms[2];
    bool two;                 // two streams in use
    bool pair_launch;         // launch the one-CTA-per-SM kernels as clusters of 2 so that they fill whole TPCs
    int dev;
    int err;
    std::vector<cudaEvent_t> taken;

    Sched() : two(false), pair_launch(false), dev(0), err(RBN_OK) { st[0] = st[1] = nullptr; sms[0] = sms[1] = 0; }
    ~Sched() { if (two) end(); }      // error returns inside a pass still join stream 1 and hand the pool resources back

    // mode 1: disjoint SM partitions (sms0 | rest);  mode 2: both streams use every SM, their kernels are the
    // 'half SM' variants that leave room for one CTA of the other stream (co-resident schedule);  mode 3: both streams
    // use every SM with the ordinary kernels -- they cannot share an SM, so only the tail of one kernel overlaps the
    // head of the other stream's next kernel
    int begin(cudaStream_t user, int mode, int sms0) {
        const bool overlap = mode != 0;
        const int total = sm_count();
        st[0] = st[1] = user;
        sms[0] = sms[1] = total;
        two = false;
        if (!overlap) return RBN_OK;
        if (sms0 < 2) sms0 = 2;
        if (sms0 > total - 2) sms0 = (total - 2) & ~1;
        RBN_CUDA_CHECK(cudaGetDevice(&dev));
        AuxPool& p = aux_pool(dev);
        {
            std::lock_guard<std::mutex> lk(p.mu);
            if (!p.streams.empty()) { st[1] = p.streams.back(); p.streams.pop_back(); two = true; }
        }
        if (!two) {
            RBN_CUDA_CHECK(cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking));
            two = true;
        }
        pair_launch = (mode == 1);
        if (mode == 1) {
            sms[0] = sms0;
            sms[1] = (total - sms0) & ~1;
        }
        cudaEvent_t e = record(0);            // fork: stream 1 starts after everything queued on the caller's stream
        wait(1, e);
        return err;
    }
    cudaEvent_t record(int s) {
        if (!two) return nullptr;
        cudaEvent_t e = nullptr;
        {
            AuxPool& p = aux_pool(dev);
            std::lock_guard<std::mutex> lk(p.mu);
            if (!p.events.empty()) { e = p.events.back(); p.events.pop_back(); }
        }
        if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            set_error("cudaEventCreate failed");
            err = RBN_E_CUDA;
            return nullptr;
        }
        taken.push_back(e);
        if (cudaEventRecord(e, st[s]) != cudaSuccess) { set_error("cudaEventRecord failed"); err = RBN_E_CUDA; }
        return e;
    }
    void wait(int s, cudaEvent_t e) {
        if (!two || !e) return;
        if (cudaStreamWaitEvent(st[s], e, 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); err = RBN_E_CUDA; }
    }
    // join: the caller's stream continues after everything queued on stream 1; hand the stream and events back
    int end() {
        if (!two) return err;
        cudaEvent_t e = record(1);
        wait(0, e);
        AuxPool& p = aux_pool(dev);
        std::lock_guard<std::mutex> lk(p.mu);
        p.streams.push_back(st[1]);
        for (cudaEvent_t ev : taken) p.events.push_back(ev);
        taken.clear();
        two = false;
        return err;
    }
};


// Diagonal w of sequence group g covers items [b0 * nI, b1 * nI); it is cut into chunks of at most `chunk` items
// (whole waves of `wave` items when possible).
static void diag_units(std::vector<Unit>& out, int w, int L, int B, int groups, int chunk, int wave, int nslot, int& counter) {
    const int nI = L - w + 1;
    int per = chunk;
    if (wave > 0 && per > wave) per = per / wave * wave;
    for (int g = 0; g < groups; ++g) {
        const int b0 = (int)((long long)B * g / groups), b1 = (int)((long long)B * (g + 1) / groups);
        const int i0 = b0 * nI, i1 = b1 * nI;
        for (int m0 = i0; m0 < i1; m0 += per) {
            Unit u;
            u.w = w; u.g = g; u.m0 = m0; u.cnt = (i1 - m0 < per) ? i1 - m0 : per; u.slot = counter % nslot; u.coff = -1;
            ++counter;
            out.push_back(u);
        }
    }
}

// Every unit of a pass, per diagonal.  The inside and the outside pass of one workspace build the SAME plan (same
// descriptor, same `active`), so a unit's place in the split-sum cache is known to both without any device state.
// The cache goes to the widest diagonals first: their recompute is the most expensive one (most splits per span).
static long long plan_units(std::vector<std::vector<Unit>>& by_w, int L, int B, const int* active, int groups, int chunk, int wave,
                            long long cache_spans) {
    by_w.assign((size_t)L + 1, std::vector<Unit>());
    int counter = 0;
    for (int w = 2; w <= L; ++w) diag_units(by_w[w], w, L, active ? active[w] : B, groups, chunk, wave, 1, counter);
    long long used = 0;
    for (int w = L; w >= 2 && used < cache_spans; --w)
        for (Unit& u : by_w[w]) {
            if (used + u.cnt > cache_spans) continue;       // a later, smaller unit may still fit
            u.coff = used;
            used += u.cnt;
        }
    return used;           // spans kept: they are rows [0, used) of the cache, in the order the outside pass visits them
}

template <class... KArgs, class... Args>
static cudaError_t launch_maybe_paired(void (*kern)(KArgs...), bool paired, int grid, int threads, size_t smem, cudaStream_t st,
                                       Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[2];
    unsigned na = 0;
    if (paired) grid = (grid + 1) & ~1;
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    if (paired) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 2;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (tc_tune().pdl) {               // both kernels launched through here call ptx::grid_dep_sync() after their prologue
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static int prep_weights(TcCtx& c, const float* W, cudaStream_t st) {
    int N = c.N;
    int* bits = c.wbits;
    RBN_CUDA_CHECK(cudaMemsetAsync(bits, 0, (size_t)2 * N * sizeof(int), st));
    { ProfScope ps(KC_OTHER, st); k_w_colmax<<<256, 256, 0, st>>>(N, W, bits); }
    RBN_LAUNCH_CHECK("k_w_colmax");
    { ProfScope ps(KC_OTHER, st); k_w_exp<<<(N + 127) / 128, 128, 0, st>>>(N, bits, c.wexp, c.wscale, c.wdens, c.kappa * kSegBlocks); }
    RBN_LAUNCH_CHECK("k_w_exp");
    dim3 grid((unsigned)((size_t)N * N / 32), (unsigned)(N / 32));
    { ProfScope ps(KC_OTHER, st); k_w_copies<<<grid, dim3(32, 8), 0, st>>>(N, W, c.wexp, c.wk, c.wt); }
    RBN_LAUNCH_CHECK("k_w_copies");
    return RBN_OK;
}

// Split-sum launch over the window [j0, j0 + ns) of every span's splits; Y / exponents go to (Ydst, Edst).
static int launch_k1_window(TcCtx& c, const Unit& u, cudaStream_t st, int sms, bool paired, bool half_sm, int j0, int ns,
                            __half* Ydst, int* Edst) {
    int N = c.N, L = c.L, B = c.B;
    const int w = u.w, m0 = u.m0, cnt = u.cnt;
    const int Kpad = (ns + 15) & ~15;
    const uint32_t rows0 = (uint32_t)(Kpad < 64 ? Kpad : 64), rows1 = (uint32_t)(Kpad > 64 ? Kpad - 64 : 16);
    CUtensorMap tmLC, tmRC, tmLC1, tmRC1;
    int rc = make_tmap_2d(&tmLC, c.lc, (uint64_t)N, (uint64_t)B * L * c.Lp, (uint64_t)N * 2, rows0);
    if (rc) return rc;
    rc = make_tmap_2d(&tmRC, c.rc, (uint64_t)N, (uint64_t)B * (L + 1) * c.Lp, (uint64_t)N * 2, rows0);
    if (rc) return rc;
    rc = make_tmap_2d(&tmLC1, c.lc, (uint64_t)N, (uint64_t)B * L * c.Lp, (uint64_t)N * 2, rows1);
    if (rc) return rc;
    rc = make_tmap_2d(&tmRC1, c.rc, (uint64_t)N, (uint64_t)B * (L + 1) * c.Lp, (uint64_t)N * 2, rows1);
    if (rc) return rc;
    CUtensorMap tmY, tmYtail;
    rc = make_tmap_2d(&tmY, Ydst, (uint64_t)N, (uint64_t)cnt * N, (uint64_t)N * 2, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmYtail, Ydst, (uint64_t)N, (uint64_t)cnt * N, (uint64_t)N * 2, 64);
    if (rc) return rc;
    const int NS = N > 256 ? 256 : N;
    int halves = (NS + 127) / 128;
    const bool pairm = (N == 64) && (Kpad <= 64) && (ns == w - 1);      // two spans per CTA iteration (see k_tc_splitsum)
    const int items = pairm ? (cnt + 1) / 2 : cnt;
    size_t stage_bytes = pairm ? (size_t)4 * Kpad * 128 : (size_t)(halves * 2 + NS / 64) * 8192;
    // small N: a span is a few KB of operands and 8-32 KB of output, the per-span hand-offs dominate -> co-resident CTAs
    static const int k1_ctas_env = env_int("RBN_TC_K1_CTAS", 0);
    int ctas = k1_ctas_env > 0 ? (k1_ctas_env > 2 ? 2 : k1_ctas_env) : (N <= 128 ? 2 : 1);
    if (paired) ctas = 1;
    // staging tiles of the Y stores: 4 (three stores in flight) when the CTA has the SM to itself, else 2
    static const int nbuf_env = env_int("RBN_TC_K1_NBUF", 0);
    int nbuf = (nbuf_env >= 2 && nbuf_env <= 4) ? nbuf_env : (ctas == 1 && !half_sm ? 4 : 2);
    if (ctas > 1 || half_sm) nbuf = 2;
    int stages = (int)((232448 / ctas - 1024 * ctas - 2048 - nbuf * 16384) / stage_bytes);
    static const int k1_stages_env = env_int("RBN_TC_K1_STAGES", 0);
    const int max_stages = k1_stages_env > 0 ? (k1_stages_env > 8 ? 8 : k1_stages_env) : (pairm ? 8 : 6);
    if (stages > max_stages) stages = max_stages;
    if (stages < 1) { stages = 1; ctas = 1; }
    int nslots = 2;
    if (half_sm) {              // leave half of the SM (shared memory, TMEM, registers) to a co-resident GEMM CTA
        if (NS == 256) stage_bytes = (size_t)(2 + NS / 64) * 8192;      // halves become sub-items (one accumulator)
        stages = (int)((113 * 1024 - 2048 - 32768) / stage_bytes);
        if (stages < 1) stages = 1;
        if (stages > 3) stages = 3;
        nslots = 1;
        ctas = 2;               // selects the register-capped instantiation; the grid stays one CTA per SM
    }
    size_t smem = stages * stage_bytes + (size_t)nbuf * 16384 /*Y staging*/ + 1024 /*barriers, cj*/ + 1024 /*alignment*/;
    auto kern = pairm ? (ctas > 1 ? k_tc_splitsum<2, true> : k_tc_splitsum<1, true>)
                      : (ctas > 1 ? k_tc_splitsum<2, false> : k_tc_splitsum<1, false>);
    RBN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = items < sms * ctas ? items : sms * ctas;
    if (half_sm && grid > sms) grid = sms;
    { ProfScope ps(KC_SPLITSUM, st);
      // measured (round 2, same box, alternating runs): two groups 157 ms of split-sum per C3 step, one group 152 -- the
      // single staging tile per group serialises on the TMA store's smem read.  With two staging tiles per group (what
      // RBN_TC_K1_GROUPS=2 selects now, out of the same four tiles) it is a wash: 153.4 vs 151.2 ms over three alternating
      // runs each on a 1.5 GHz box -- the epilogue's ~960 instructions per span are not what paces this kernel.
      static const int k1_groups = env_int("RBN_TC_K1_GROUPS", 1);
      RBN_CUDA_CHECK(launch_maybe_paired(kern, paired, grid, (ctas > 1 || k1_groups < 2) ? k1_threads<2>() : k1_threads<1>(), smem, st, tmLC, tmRC, tmLC1, tmRC1, tmY, tmYtail, N, L,
                                         c.Lp, w, m0, cnt, stages, c.lengths, (const int*)c.ce, Ydst, Edst, nslots,
                                         (uint64_t)((tc_tune().hints & 2) ? ptx::kEvictFirst : ptx::kEvictNormal), nbuf, j0, ns)); }
    RBN_LAUNCH_CHECK("k_tc_splitsum");
    return RBN_OK;
}

// Spans wider than 129 tokens have more than the 128 splits one split-sum launch holds (two K slabs of 64): the launch is
// repeated per window of 128 splits into a scratch buffer of the workspace (WsLayout::ywin) and the partial sums, each
// under its own scale exponent, are folded into the unit's Y.
static int launch_k1(TcCtx& c, const Unit& u, cudaStream_t st, int sms, bool paired, bool half_sm = false) {
    const int S = u.w - 1;
    if (S <= kMaxWindow) return launch_k1_window(c, u, st, sms, paired, half_sm, 0, S, c.y_of(u), c.e_of(u));
    int rc = launch_k1_window(c, u, st, sms, paired, half_sm, 0, kMaxWindow, c.y_of(u), c.e_of(u));
    __half* yw = (__half*)(c.base + c.ws->ywin);
    int* ew = (int*)(c.base + c.ws->ywin_e);
    for (int j0 = kMaxWindow; j0 < S && !rc; j0 += kMaxWindow) {
        const int ns = S - j0 < kMaxWindow ? S - j0 : kMaxWindow;
        if ((rc = launch_k1_window(c, u, st, sms, paired, half_sm, j0, ns, yw, ew))) break;
        { ProfScope ps(KC_SPLITSUM, st);
          k_y_combine<<<u.cnt, 256, 0, st>>>(c.N, c.L, u.w, u.m0, c.lengths, c.y_of(u), c.e_of(u), (const __half*)yw, (const int*)ew); }
        RBN_LAUNCH_CHECK("k_y_combine");
    }
    return rc;
}

// K-split factor of the partial last wave of a 2-CTA GEMM: r tiles left for `pairs` CTA pairs.  Splitting each into s
// K-ranges gives ceil(r s / pairs) rounds of 1/s of a tile each; pick the s with the shortest tail (ties: fewer
// partial sums to add).
// Cost model in microseconds (constants measured on B200, C3 shapes, round 2): a CTA pair spends ~0.34 us per k-block of a
// 256 x 256 tile; the partial sums of a K-split unit are red.global.add-ed into the fp32 buffer, and the chip's L2 sustains
// only ~2 TB/s of those (256 KB per unit -> 0.122 us each, chip-wide, not hidden by anything); a split launch also needs
// the buffer cleared and the finalize kernel (~8 us).  So: split as little as evens out the rounds.  (First version of this
// model ignored the L2 reduction rate and split T = 64 tiles eight ways: 20 ms per C3 step slower than not splitting.)
// `tiles_total`: all tiles of the launch whose units are partial (for the red traffic), r of them form the tail.
static int tail_splits_for(int r, int pairs, int k_blocks, bool all_tail = false, double* t_best = nullptr) {
    if (t_best) *t_best = 0;
    if (r <= 0) return 1;
    const double t_kb = 0.34, t_red = 0.122, t_fixed = 8.0;
    int best = 1;
    double best_t = (double)((r + pairs - 1) / pairs) * k_blocks * t_kb;
    for (int s2 = 2; s2 <= 16 && k_blocks / s2 >= 16; ++s2) {
        const int rounds = (r * s2 + pairs - 1) / pairs;
        const int per = (k_blocks + s2 - 1) / s2;
        const double epi = (all_tail && per <= kSegBlocks) ? 0.0 : 4.0;       // exposed TMEM -> red issue time per unit
        const double t = (double)rounds * (per * t_kb + epi) + (double)r * s2 * t_red + t_fixed;
        if (t < best_t - 1e-9) { best_t = t; best = s2; }
    }
    if (t_best) *t_best = best_t;
    return best;
}

template <int BN>
static int launch_k2_bn(TcCtx& c, const Unit& u, cudaStream_t st, int sms, bool half_sm) {
    const int w = u.w, m0 = u.m0, cnt = u.cnt;
    size_t NN = (size_t)c.N * c.N;
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d(&tmA, c.y_of(u), NN, (uint64_t)cnt, NN * 2, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmB, c.wt, NN, (uint64_t)c.N, NN * 2, BN);
    if (rc) return rc;
    // (the direct epilogue only ever sees whole tiles: chains of kSegBlocks k-blocks -> the compensated scale row)
    EpiRule epi{c.N, c.L, c.Lp, w, m0, cnt, c.lengths, c.wscale + c.N, c.e_of(u), c.cv, c.ce, c.lc, c.rc, c.acc(u.slot), BN};
    auto chain_kc = [&](int per) { return c.kappa * (float)(per < kSegBlocks ? per : kSegBlocks); };
    const int n_tiles = c.N / BN;                  // 1, or 2 when N = 512
    if (n_tiles > 1 && !use_pairs()) {
        set_error("N = %d needs the CTA-pair kernels (unset RBN_TC_1CTA)", c.N);
        return RBN_E_UNSUPPORTED;
    }
    if (half_sm && use_pairs() && n_tiles == 1) {
        // co-resident schedule: 3 stages (96 KB), ONE 256-column accumulator.  Every tile is cut along K into chains of
        // <= 64 k-blocks (no in-TMEM segmentation needed); the partial sums meet in the fp32 buffer, k_rule_finalize
        // turns them into chart cells.
        CUtensorMap tmB2;
        rc = make_tmap_2d(&tmB2, c.wt, NN, (uint64_t)c.N, NN * 2, BN / 2);
        if (rc) return rc;
        const int kb = (int)(NN / 64);
        GemmShape shape2{(cnt + 255) / 256, 1, kb, 0, (kb + kSegBlocks - 1) / kSegBlocks, 1};
        if (shape2.tail_splits < 2) shape2.tail_splits = 1;
        RBN_CUDA_CHECK(cudaMemsetAsync(c.acc(u.slot), 0, (size_t)cnt * c.N * sizeof(float), st));
        rc = launch_tc_gemm2<false, false, BN, 3, 0, EpiRule>(tmA, tmB2, shape2, epi, sms, st, "k_tc_gemm2<rule,half>", KC_RULE);
        if (rc) return rc;
        { ProfScope ps(KC_RULE_FIN, st);
          launch_pdl(k_rule_finalize, dim3((cnt + kFinWarps - 1) / kFinWarps), dim3(kFinWarps * 32), 0, st, c.N, c.L, c.Lp, w, m0, 0, cnt, c.lengths, c.wscale, c.e_of(u), c.acc(u.slot), c.cv, c.ce, c.lc, c.rc, c.wdens, chain_kc(kSegBlocks)); }
        RBN_LAUNCH_CHECK("k_rule_finalize");
        return RBN_OK;
    }
    if (use_pairs()) {
        CUtensorMap tmB2;
        rc = make_tmap_2d(&tmB2, c.wt, NN, (uint64_t)c.N, NN * 2, BN / 2);
        if (rc) return rc;
        GemmShape shape2{(cnt + 255) / 256, n_tiles, (int)(NN / 64), 0, 1, 0};
        // Y streams, W is reused.  (N = 512: a Y tile is read by the two parent-half tiles, the second time from L2 -- no hint)
        if ((tc_tune().hints & 4) && n_tiles == 1) { shape2.hint_a = ptx::kEvictFirst; shape2.hint_b = ptx::kEvictLast; }
        // wave quantisation: the last, partial wave of tiles is split along K so that all CTA pairs stay busy
        const int pairs = sms / 2, T = shape2.m_tiles * n_tiles;
        const int full = (T / pairs) * pairs, r = T - full;
        double t_split = 0;
        const int splits = tail_splits_for(r, pairs, shape2.k_blocks, full == 0 && n_tiles == 1, &t_split);
        if (full == 0 && n_tiles == 1 && T >= 1) {
            // fewer tiles than CTA pairs: stream-K (every pair gets the same number of k-blocks) when it beats the best
            // uniform split -- same constants as tail_splits_for; a pair then owns ~2 units with exposed epilogues
            const int per = (int)(((long long)T * shape2.k_blocks + pairs - 1) / pairs);
            const double t_sk = per * 0.34 + (double)(pairs + T) * 0.122 + 8.0 + 2 * 4.0;
            static const int streamk = env_int("RBN_TC_STREAMK", 1);
            if (streamk && per >= 8 && t_sk < t_split) {
                shape2.sk_per = per;
                shape2.short_units = per <= kSegBlocks ? 1 : 0;
                RBN_CUDA_CHECK(cudaMemsetAsync(c.acc(u.slot), 0, (size_t)cnt * c.N * sizeof(float), st));
                rc = launch_tc_gemm2<false, false, BN, kK2Stages, kSegBlocks, EpiRule>(tmA, tmB2, shape2, epi, sms, st, "k_tc_gemm2<rule>", KC_RULE);
                if (rc) return rc;
                { ProfScope ps(KC_RULE_FIN, st);
                  launch_pdl(k_rule_finalize, dim3((cnt + kFinWarps - 1) / kFinWarps), dim3(kFinWarps * 32), 0, st, c.N, c.L, c.Lp, w, m0, 0, cnt, c.lengths, c.wscale, c.e_of(u), c.acc(u.slot), c.cv, c.ce, c.lc, c.rc, c.wdens, chain_kc(per)); }
                RBN_LAUNCH_CHECK("k_rule_finalize");
                return RBN_OK;
            }
        }
        if (n_tiles > 1) {
            // both halves of a row are needed for its max shift: every tile adds into acc, k_rule_finalize does the rest
            shape2.all_partial = 1;
            if (splits > 1) { shape2.full_tiles = full; shape2.tail_splits = splits; }
            RBN_CUDA_CHECK(cudaMemsetAsync(c.acc(u.slot), 0, (size_t)cnt * c.N * sizeof(float), st));
            rc = launch_tc_gemm2<false, false, BN, kK2Stages, kSegBlocks, EpiRule>(tmA, tmB2, shape2, epi, sms, st, "k_tc_gemm2<rule>", KC_RULE);
            if (rc) return rc;
            { ProfScope ps(KC_RULE_FIN, st);
              launch_pdl(k_rule_finalize, dim3((cnt + kFinWarps - 1) / kFinWarps), dim3(kFinWarps * 32), 0, st, c.N, c.L, c.Lp, w, m0, 0, cnt, c.lengths, c.wscale, c.e_of(u), c.acc(u.slot), c.cv, c.ce, c.lc, c.rc, c.wdens, chain_kc(kSegBlocks)); }
            RBN_LAUNCH_CHECK("k_rule_finalize");
            return RBN_OK;
        }
        if (splits > 1) {
            shape2.full_tiles = full;
            shape2.tail_splits = splits;
            shape2.short_units = (shape2.k_blocks <= kSegBlocks || (full == 0 && (shape2.k_blocks + splits - 1) / splits <= kSegBlocks)) ? 1 : 0;
            const int t_begin = full * 256;
            RBN_CUDA_CHECK(cudaMemsetAsync(c.acc(u.slot) + (size_t)t_begin * c.N, 0, (size_t)(cnt - t_begin) * c.N * sizeof(float), st));
            rc = launch_tc_gemm2<false, false, BN, kK2Stages, kSegBlocks, EpiRule>(tmA, tmB2, shape2, epi, sms, st, "k_tc_gemm2<rule>", KC_RULE);
            if (rc) return rc;
            { ProfScope ps(KC_RULE_FIN, st);
              launch_pdl(k_rule_finalize, dim3((cnt - t_begin + kFinWarps - 1) / kFinWarps), dim3(kFinWarps * 32), 0, st, c.N, c.L, c.Lp, w, m0, t_begin, cnt, c.lengths, c.wscale, c.e_of(u), c.acc(u.slot), c.cv, c.ce, c.lc, c.rc, c.wdens, chain_kc((shape2.k_blocks + splits - 1) / splits)); }
            RBN_LAUNCH_CHECK("k_rule_finalize");
            return RBN_OK;
        }
        // N = 64: a whole tile is ONE accumulation segment (64 k-blocks), so consecutive tiles alternate between the two
        // accumulators and a tile's max-shift epilogue runs behind the next tile's MMAs
        shape2.short_units = (shape2.k_blocks <= kSegBlocks) ? 1 : 0;
        return launch_tc_gemm2<false, false, BN, kK2Stages, kSegBlocks, EpiRule>(tmA, tmB2, shape2, epi, sms, st, "k_tc_gemm2<rule>", KC_RULE);
    }
    GemmShape shape{(cnt + 127) / 128, 1, (int)(NN / 64), 0, 1, 0};
    return launch_tc_gemm<false, false, BN, (BN > 128 ? 4 : 6), kSegBlocks, EpiRule>(tmA, tmB, shape, epi, sms, st, "k_tc_gemm<rule>", KC_RULE);
}
static int launch_k2(TcCtx& c, const Unit& u, cudaStream_t st, int sms, bool half_sm = false) {
    switch (c.N) {
        case 64: return launch_k2_bn<64>(c, u, st, sms, false);
        case 128: return launch_k2_bn<128>(c, u, st, sms, false);
        case 192: return launch_k2_bn<192>(c, u, st, sms, false);
        default: return launch_k2_bn<256>(c, u, st, sms, half_sm);
    }
}

static int launch_k3(TcCtx& c, const Unit& u, cudaStream_t st, int sms) {
    const int cnt = u.cnt;
    size_t NN = (size_t)c.N * c.N;
    CUtensorMap tmG, tmW, tmOut;
    int rc = make_tmap_2d(&tmG, c.ghat(u.slot), (uint64_t)c.N, (uint64_t)cnt, (uint64_t)c.N * 2, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmW, c.wk, (uint64_t)c.N, NN, (uint64_t)c.N * 2, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmOut, c.gy(u.slot), NN, (uint64_t)cnt, NN * 2, 128);
    if (rc) return rc;
    int supertiles = (cnt + 255) / 256;
    static const int gy_variant = env_int("RBN_TC_GY", 2);        // 3: gradient rows in TMEM, 2: in shared memory
    if (use_pairs() && gy_variant == 3) {
        CUtensorMap tmW64;
        rc = make_tmap_2d(&tmW64, c.wk, (uint64_t)c.N, NN, (uint64_t)c.N * 2, 64);
        if (rc) return rc;
        int ntiles3 = (int)(NN / 128), pairs = sms / 2;
        int ranges3 = 1;
        while (ranges3 < 16 && supertiles * ranges3 < 2 * pairs && ntiles3 % (ranges3 * 2) == 0 && ntiles3 / (ranges3 * 2) >= 16) ranges3 *= 2;
        size_t smem3 = (size_t)kGY3Stages * 8192 + 65536 + 512 + 1024;
        RBN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gy3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        int units3 = supertiles * ranges3;
        int np = units3 < pairs ? units3 : pairs;
        { ProfScope ps(KC_GY, st); launch_pdl(k_tc_gy3, dim3(2 * np), dim3(384), smem3, st, tmW64, tmOut, c.ghat(u.slot), cnt, c.N, supertiles, ranges3); }
        RBN_LAUNCH_CHECK("k_tc_gy3");
        return RBN_OK;
    }
    if (use_pairs()) {
        int ntiles2 = (int)(NN / 256), pairs = sms / 2;
        // A unit = one 256-span supertile x one range of Wk tiles; it loads its gradient rows once (about one Wk tile's
        // worth of bytes) and streams its tiles.  Pick the number of ranges that minimises rounds x (tiles + 1): enough units
        // to fill the CTA pairs evenly when a diagonal has few supertiles, without reloading the rows needlessly.
        int ranges2 = 1;
        {
            double best = 1e30;
            for (int r = 1; r <= 64 && ntiles2 / r >= 4; ++r) {
                const long long units = (long long)supertiles * r;
                const double cost = (double)((units + pairs - 1) / pairs) * ((ntiles2 + r - 1) / r + 1.0);
                if (cost < best * 0.995) { best = cost; ranges2 = r; }
            }
        }
        static const int deep = env_int("RBN_TC_GY_STAGES", 6);       // N <= 256: 6 ring stages (4: the former setting)
        const bool big = c.N > 256;
        static const int gy_nbuf = env_int("RBN_TC_GY_NBUF", 2);      // 3: two stores in flight per epilogue group, 4-stage Wk ring
        const bool three = !big && gy_nbuf == 3;
        const int nbuf = big ? 1 : (three ? 3 : 2);
        const int stg = big ? 4 : (three ? 4 : (deep >= 6 ? 6 : 4));
        size_t smem2 = (size_t)(c.N / 64) * 16384 + (size_t)stg * 16384 + (size_t)2 * nbuf * 16384 + 256 + 1024;
        auto kern = big ? k_tc_gy2<1, 4> : (three ? k_tc_gy2<3, 4> : (stg == 6 ? k_tc_gy2<2, 6> : k_tc_gy2<2, 4>));
        RBN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        int units2 = supertiles * ranges2;
        int np = units2 < pairs ? units2 : pairs;
        { ProfScope ps(KC_GY, st); launch_pdl(kern, dim3(2 * np), dim3(384), smem2, st, tmG, tmW, tmOut, c.N, supertiles, ranges2,
                                                                     (uint64_t)((tc_tune().hints & 2) ? ptx::kEvictFirst : ptx::kEvictNormal),
                                                                     (uint64_t)((tc_tune().hints & 4) ? ptx::kEvictLast : ptx::kEvictNormal)); }
        RBN_LAUNCH_CHECK("k_tc_gy2");
        return RBN_OK;
    }
    int ntiles = (int)(NN / 128);
    int ranges = 1;
    while (ranges < 16 && supertiles * ranges < 2 * sms && ntiles % (ranges * 2) == 0 && ntiles / (ranges * 2) >= 8) ranges *= 2;
    size_t smem = (size_t)2 * (c.N / 64) * 16384 + kGYStages * 16384 + 32768 + 256 + 1024;
    RBN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int units = supertiles * ranges;
    int grid = units < sms ? units : sms;
    { ProfScope ps(KC_GY, st); launch_pdl(k_tc_gy, dim3(grid), dim3(256), smem, st, tmG, tmW, tmOut, c.N, supertiles, ranges); }
    RBN_LAUNCH_CHECK("k_tc_gy");
    return RBN_OK;
}

template <int BN>
static int launch_k4_bn(TcCtx& c, const Unit& u, float* gW, cudaStream_t st, int sms) {
    const int cnt = u.cnt;
    size_t NN = (size_t)c.N * c.N;
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d(&tmA, c.y_of(u), NN, (uint64_t)cnt, NN * 2, 64);            // MN-major A: [K=items][M=(b,c)]
    if (rc) return rc;
    rc = make_tmap_2d(&tmB, c.gabs_of(u), (uint64_t)c.N, (uint64_t)cnt, (uint64_t)c.N * 2, 64);   // MN-major B: [K=items][N=a]
    if (rc) return rc;
    const int gi = u.coff >= 0 ? 0 : 1 + u.slot;      // scale group: all kept spans, or this unit alone
    {
        const long long n = (long long)cnt * c.N;
        int blocks = (int)((n / 4 + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        ProfScope ps(KC_PREP, st);
        k_gabs_convert<<<blocks, 256, 0, st>>>(n, c.g32_of(u), c.gabs_of(u), c.gmax(gi), c.gscale(gi));
    }
    RBN_LAUNCH_CHECK("k_gabs_convert");
    EpiCounts epi{gW, c.N, BN, c.gscale(gi)};
    const int n_tiles = c.N / BN;
    if (n_tiles > 1 && !use_pairs()) {
        set_error("N = %d needs the CTA-pair kernels (unset RBN_TC_1CTA)", c.N);
        return RBN_E_UNSUPPORTED;
    }
    if (use_pairs() && BN % 128 == 0) {
        GemmShape shape2{(int)(NN / 256), n_tiles, (cnt + 63) / 64, 0, 1, 0};
        if ((tc_tune().hints & 4) && n_tiles == 1) shape2.hint_a = ptx::kEvictFirst;                     // last read of Y
        {   // K-split so that the work units fill whole waves of CTA pairs (256 tiles on 74 pairs = 3.46 waves otherwise)
            const int pairs = sms / 2, tiles = (int)(NN / 256) * n_tiles;
            int best = 1;
            double best_eff = 0;
            for (int sp = 1; sp <= 4; ++sp) {
                if (sp > 1 && shape2.k_blocks / sp < 16) break;
                const int units = tiles * sp;
                const double eff = (double)units / (((units + pairs - 1) / pairs) * (double)pairs);
                if (eff > best_eff + 0.02) { best_eff = eff; best = sp; }
            }
            shape2.tail_splits = best;
        }
        return launch_tc_gemm2<true, true, (BN % 128 == 0 ? BN : 128), kK2Stages, kSegBlocks, EpiCounts>(tmA, tmB, shape2, epi, sms, st, "k_tc_gemm2<counts>", KC_COUNTS);
    }
    GemmShape shape{(int)(NN / 128), 1, (cnt + 63) / 64, 0, 1, 0};
    return launch_tc_gemm<true, true, BN, (BN > 128 ? 4 : 6), kSegBlocks, EpiCounts>(tmA, tmB, shape, epi, sms, st, "k_tc_gemm<counts>", KC_COUNTS);
}
static int launch_k4(TcCtx& c, const Unit& u, float* gW, cudaStream_t st, int sms) {
    switch (c.N) {
        case 64: return launch_k4_bn<64>(c, u, gW, st, sms);
        case 128: return launch_k4_bn<128>(c, u, gW, st, sms);
        case 192: return launch_k4_bn<192>(c, u, gW, st, sms);
        default: return launch_k4_bn<256>(c, u, gW, st, sms);
    }
}

static int launch_k5_window(TcCtx& c, const Unit& u, cudaStream_t st, int sms, bool paired, int j0, int ns) {
    int N = c.N, L = c.L, B = c.B;
    const int w = u.w, m0 = u.m0, cnt = u.cnt;
    int Npad = (ns + 15) & ~15;
    CUtensorMap tmGYk, tmGYm, tmRC, tmLC;
    int rc = make_tmap_2d(&tmGYk, c.gy(u.slot), (uint64_t)N, (uint64_t)cnt * N, (uint64_t)N * 2, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmGYm, c.gy(u.slot), (uint64_t)N, (uint64_t)cnt * N, (uint64_t)N * 2, 64);
    if (rc) return rc;
    rc = make_tmap_2d(&tmRC, c.rc, (uint64_t)N, (uint64_t)B * (L + 1) * c.Lp, (uint64_t)N * 2, (uint32_t)Npad);
    if (rc) return rc;
    rc = make_tmap_2d(&tmLC, c.lc, (uint64_t)N, (uint64_t)B * L * c.Lp, (uint64_t)N * 2, (uint32_t)Npad);
    if (rc) return rc;
    CUtensorMap tmAL, tmAR, tmALt, tmARt;
    const uint32_t tail_rows = (uint32_t)((ns % 32) ? (ns % 32) : 32);
    rc = make_tmap_2d_f32(&tmAL, c.alpha, (uint64_t)N, (uint64_t)B * L * c.Lp, (uint64_t)N * 4, 128, 32);
    if (rc) return rc;
    rc = make_tmap_3d_f32(&tmAR, c.alpha, (uint64_t)N, (uint64_t)c.Lp, (uint64_t)B * L, 128, 32);
    if (rc) return rc;
    rc = make_tmap_2d_f32(&tmALt, c.alpha, (uint64_t)N, (uint64_t)B * L * c.Lp, (uint64_t)N * 4, 128, tail_rows);
    if (rc) return rc;
    rc = make_tmap_3d_f32(&tmARt, c.alpha, (uint64_t)N, (uint64_t)c.Lp, (uint64_t)B * L, 128, tail_rows);
    if (rc) return rc;
    const size_t stage_bytes = 16384 + (size_t)Npad * 128;
    const size_t fixed = 65536 + 512 + 2048 + 1024;          // reduce staging, barriers, split factors, alignment
    int stages = (int)((232448 - fixed) / stage_bytes);
    static const int stage_cap = env_int("RBN_TC_K5_STAGES", kK5MaxStages);
    if (stages > kK5MaxStages) stages = kK5MaxStages;
    if (stages > stage_cap && stage_cap >= 2) stages = stage_cap;
    size_t smem = (size_t)stages * stage_bytes + fixed;
    RBN_CUDA_CHECK(cudaFuncSetAttribute(k_tc_childgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = cnt < sms ? cnt : sms;
    { ProfScope ps(KC_CHILD, st);
      RBN_CUDA_CHECK(launch_maybe_paired(k_tc_childgrad, paired, grid, 384, smem, st, tmGYk, tmGYm, tmRC, tmLC, N, L, c.Lp, w, m0, cnt,
                                         c.lengths, tmAL, tmAR, tmALt, tmARt, (const int*)c.ce, (const int*)c.e_of(u),
                                         (const int*)c.gexp(u.slot),
                                         (uint64_t)((tc_tune().hints & 1) ? ptx::kEvictLast : ptx::kEvictNormal),
                                         (uint64_t)((tc_tune().hints & 8) ? ptx::kEvictFirst : ptx::kEvictNormal), stages, j0, ns)); }
    RBN_LAUNCH_CHECK("k_tc_childgrad");
    return RBN_OK;
}

static int launch_k5(TcCtx& c, const Unit& u, cudaStream_t st, int sms, bool paired) {
    const int S = u.w - 1;
    int rc = RBN_OK;
    for (int j0 = 0; j0 < S && !rc; j0 += kMaxWindow)
        rc = launch_k5_window(c, u, st, sms, paired, j0, S - j0 < kMaxWindow ? S - j0 : kMaxWindow);
    return rc;
}

// Inside pass.  Stream 0 runs the split-sums (HBM-write bound), stream 1 the rule GEMMs (tensor / HBM-read bound).
// A split-sum of diagonal w needs every narrower cell of ITS sequences, i.e. the rule GEMMs of the same sequence
// group on diagonal w - 1; the other group's kernels fill the gap.
int tc_forward(const RbnDesc* d, const WsLayout& ws, char* base, const float* W, const float* T, const float* prior,
               const int* tokens, const int* lengths, float* logZ, float* zroot, float* rootexp, const int* active,
               cudaStream_t st) {
    TcCtx c;
    fill_ctx(c, d, ws, base, lengths);
    int N = c.N, L = c.L, B = c.B;
    const TcTune& tune = tc_tune();
    int rc = prep_weights(c, W, st);
    if (rc) return rc;
    // (the split-sum buffers need no clearing: K1 writes every span of a unit, zeros for spans beyond their sequence)
    // TMA slabs over-read rows of cells that are never written (start >= end, beyond a sequence's length): keep them 0
    RBN_CUDA_CHECK(cudaMemsetAsync(c.lc, 0, (size_t)B * L * c.Lp * N * sizeof(__half), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(c.rc, 0, (size_t)B * (L + 1) * c.Lp * N * sizeof(__half), st));
    { ProfScope ps(KC_EMISSION, st); k_emission<<<B * L, 256, 0, st>>>(N, L, d->n_tvars, T, tokens, lengths, c.cv, c.ce); }
    RBN_LAUNCH_CHECK("k_emission");
    { ProfScope ps(KC_EMISSION, st); k_copy_width1<<<B * L, 128, 0, st>>>(N, L, c.Lp, lengths, c.cv, c.lc, c.rc); }
    RBN_LAUNCH_CHECK("k_copy_width1");

    // (spans wider than 129 tokens share ONE scratch buffer for their split windows: single-stream schedule only)
    const int mode = (tune.overlap && ws.nslot >= 2 && B >= 2 * tune.groups && L <= kMaxWindow + 1) ? tune.overlap : 0;
    const bool overlap = mode != 0, half_sm = (mode == 2);
    const int groups = overlap ? tune.groups : 1;
    Sched S;
    if ((rc = S.begin(st, mode, tune.fwd_sms0))) return rc;
    const int wave = (S.sms[1] / 2) * 256;                     // rule-GEMM rows per wave of CTA pairs
    std::vector<Unit> units;
    std::vector<cudaEvent_t> evK2;
    std::vector<int> last_prev(groups, -1), last_cur(groups, -1);
    int counter = 0;
    // sorted ragged batch: only the first active[w] sequences reach width w (not combined with sequence groups)
    std::vector<std::vector<Unit>> plan;
    plan_units(plan, L, B, overlap ? nullptr : active, groups, c.chunk, wave, overlap ? 0 : ws.cache_spans);
    for (int w = 2; w <= L && !rc; ++w) {
        const size_t first = units.size();
        for (Unit u : plan[w]) { u.slot = counter++ % ws.nslot; units.push_back(u); }
        evK2.resize(units.size(), nullptr);
        for (size_t ui = first; ui < units.size() && !rc; ++ui) {
            const Unit& u = units[ui];
            if (last_prev[u.g] >= 0) S.wait(0, evK2[last_prev[u.g]]);            // narrower cells of this group are written
            if ((int)ui - ws.nslot >= 0) S.wait(0, evK2[ui - ws.nslot]);         // the slot's previous reader is done
            if ((rc = launch_k1(c, u, S.st[0], S.sms[0], S.pair_launch, half_sm))) break;
            S.wait(1, S.record(0));
            if ((rc = launch_k2(c, u, S.st[1], S.sms[1], half_sm))) break;
            evK2[ui] = S.record(1);
            last_cur[u.g] = (int)ui;
        }
        last_prev = last_cur;
    }
    int rc2 = S.end();
    if (rc) return rc;
    if (rc2) return rc2;
    { ProfScope ps(KC_OTHER, st); k_root<<<B, 128, 0, st>>>(N, L, prior, lengths, c.cv, c.ce, logZ, zroot, rootexp, (float*)(base + ws.zroot),
                              (int*)(base + ws.rootexp)); }
    RBN_LAUNCH_CHECK("k_root");
    return RBN_OK;
}

// Outside pass.  Stream 0: split-sum recompute, upstream-gradient prep, gY GEMM (HBM-write bound); stream 1: count GEMM
// and child gradients (HBM-read / reduce bound).  The prep of diagonal w needs the child gradients of every wider
// diagonal of ITS sequence group.
int tc_backward(const RbnDesc* d, const WsLayout& ws, char* base, const float* W, const float* T, const float* prior,
                const int* tokens, const int* lengths, const float* coef, float* gW, float* gT, float* gPrior,
                const int* active, cudaStream_t st) {
    (void)W; (void)T;
    TcCtx c;
    fill_ctx(c, d, ws, base, lengths);
    int N = c.N, L = c.L, B = c.B;
    const TcTune& tune = tc_tune();
    RBN_CUDA_CHECK(cudaMemsetAsync(c.alpha, 0, (size_t)B * L * c.Lp * N * sizeof(float), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(gW, 0, (size_t)N * N * N * sizeof(float), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(gT, 0, (size_t)d->V * N * sizeof(float), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(gPrior, 0, (size_t)N * sizeof(float), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(c.gmax(0), 0, 32 * sizeof(int), st));
    { ProfScope ps(KC_OTHER, st); k_tc_outside_seed<<<B, 128, 0, st>>>(N, L, c.Lp, prior, lengths, c.cv, (const float*)(base + ws.zroot), coef, c.alpha, gPrior); }
    RBN_LAUNCH_CHECK("k_tc_outside_seed");

    // co-resident mode (2): inside pass only so far
    const int mode = ((tune.overlap == 1 || tune.overlap == 3) && ws.nslot >= 2 && B >= 2 * tune.groups && L <= kMaxWindow + 1) ? tune.overlap : 0;
    const bool overlap = mode != 0;
    const int groups = overlap ? tune.groups : 1;
    Sched S;
    int rc;
    if ((rc = S.begin(st, mode, tune.bwd_sms0))) return rc;
    const int s3 = (overlap && tune.k3_stream) ? 1 : 0;        // stream of the gY GEMM
    const int wave = (S.sms[s3] / 2) * 256;
    std::vector<Unit> units;
    std::vector<cudaEvent_t> evK5;
    std::vector<int> last_prev(groups, -1), last_cur(groups, -1);
    int counter = 0;
    std::vector<std::vector<Unit>> plan;
    const long long kept = plan_units(plan, L, B, overlap ? nullptr : active, groups, c.chunk, wave, overlap ? 0 : ws.cache_spans);
    for (int w = L; w >= 2 && !rc; --w) {
        const size_t first = units.size();
        for (Unit u : plan[w]) { u.slot = counter++ % ws.nslot; units.push_back(u); }
        evK5.resize(units.size(), nullptr);
        for (size_t ui = first; ui < units.size() && !rc; ++ui) {
            const Unit& u = units[ui];
            const int cnt = u.cnt;
            if ((int)ui - ws.nslot >= 0) S.wait(0, evK5[ui - ws.nslot]);         // every reader of the slot's buffers is done
            // split-sum: read back from the cache when the inside pass kept it, recomputed otherwise
            if (u.coff < 0 && (rc = launch_k1(c, u, S.st[0], S.sms[0], S.pair_launch))) break;
            if (last_prev[u.g] >= 0) S.wait(0, evK5[last_prev[u.g]]);            // alpha of this diagonal is complete
            int pg = (cnt + 63) / 64 * 64;      // pad rows are written as zeros
            if (u.coff < 0) RBN_CUDA_CHECK(cudaMemsetAsync(c.gmax(1 + u.slot), 0, sizeof(int), S.st[0]));
            { ProfScope ps(KC_PREP, S.st[0]);
              launch_pdl(k_prep_g, dim3((pg + kPrepWarps - 1) / kPrepWarps), dim3(kPrepWarps * 32), 0, S.st[0], N, L, c.Lp, u.w, u.m0, cnt, pg, u.coff >= 0 ? cnt : pg, lengths, c.ce, c.e_of(u), c.wexp, c.alpha,
                                             c.ghat(u.slot), c.g32_of(u), c.gmax(u.coff >= 0 ? 0 : 1 + u.slot), c.gexp(u.slot)); }
            RBN_LAUNCH_CHECK("k_prep_g");
            S.wait(1, S.record(0));
            // expected rule counts: spans whose Y sits in the cache wait for ONE count GEMM over all of them after the
            // last diagonal (K = all kept spans instead of one diagonal's: whole waves of tiles, one pass over gW)
            if (s3 == 0) {
                if ((rc = launch_k3(c, u, S.st[0], S.sms[0]))) break;
                cudaEvent_t e3 = S.record(0);
                if (u.coff < 0 && (rc = launch_k4(c, u, gW, S.st[1], S.sms[1]))) break;
                S.wait(1, e3);
            } else {
                if ((rc = launch_k3(c, u, S.st[1], S.sms[1]))) break;
                if (u.coff < 0 && (rc = launch_k4(c, u, gW, S.st[1], S.sms[1]))) break;
            }
            if ((rc = launch_k5(c, u, S.st[1], S.sms[1], S.pair_launch))) break;
            evK5[ui] = S.record(1);
            last_cur[u.g] = (int)ui;
        }
        last_prev = last_cur;
    }
    int rc2 = S.end();
    if (rc) return rc;
    if (rc2) return rc2;
    if (kept > 0) {
        Unit all;
        all.w = 0; all.g = 0; all.m0 = 0; all.slot = 0; all.coff = 0;
        all.cnt = (int)kept;
        if ((rc = launch_k4(c, all, gW, st, sm_count()))) return rc;
    }
    { ProfScope ps(KC_EMISSION, st); k_tc_emission_bwd<<<B * L, 128, 0, st>>>(N, L, c.Lp, d->n_tvars, tokens, lengths, c.ce, c.alpha, gT); }
    RBN_LAUNCH_CHECK("k_emission_bwd");
    return RBN_OK;
}

// ---- diagnostics: plain GEMM through the same kernel template (tests/test_gpu_tc.py) -----------------------------------
int debug_gemm(int M, int N, int K, const __half* A, const __half* Bm, float* D, int a_mn, int b_mn, int seg, cudaStream_t st) {
    if (N != 256 || K % 64 != 0) {
        set_error("debug gemm: N must be 256 and K a multiple of 64");
        return RBN_E_BADARG;
    }
    CUtensorMap tmA, tmB;
    int rc;
    if (!a_mn) rc = make_tmap_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, 128);
    else rc = make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)M * 2, 64);
    if (rc) return rc;
    if (!b_mn) rc = make_tmap_2d(&tmB, Bm, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, 256);
    else rc = make_tmap_2d(&tmB, Bm, (uint64_t)N, (uint64_t)K, (uint64_t)N * 2, 64);
    if (rc) return rc;
    EpiStoreF32 epi{D, M, N, 256};
    GemmShape shape{(M + 127) / 128, 1, K / 64, 0, 1, 0};
    int sms = sm_count();
    if (seg & 2) {       // 2-CTA variant (tests): K-major/K-major and MN-major/MN-major only
        CUtensorMap tmB2;
        if (!b_mn) rc = make_tmap_2d(&tmB2, Bm, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, 128);
        else rc = make_tmap_2d(&tmB2, Bm, (uint64_t)N, (uint64_t)K, (uint64_t)N * 2, 64);
        if (rc) return rc;
        GemmShape shape2{(M + 255) / 256, 1, K / 64, 0, 1, 0};
        if (!a_mn && !b_mn) {
            if (seg & 1) return launch_tc_gemm2<false, false, 256, 6, kSegBlocks, EpiStoreF32>(tmA, tmB2, shape2, epi, sms, st, "debug_gemm2", KC_OTHER);
            return launch_tc_gemm2<false, false, 256, 6, 0, EpiStoreF32>(tmA, tmB2, shape2, epi, sms, st, "debug_gemm2", KC_OTHER);
        }
        if (a_mn && b_mn) {
            if (seg & 1) return launch_tc_gemm2<true, true, 256, 6, kSegBlocks, EpiStoreF32>(tmA, tmB2, shape2, epi, sms, st, "debug_gemm2", KC_OTHER);
            return launch_tc_gemm2<true, true, 256, 6, 0, EpiStoreF32>(tmA, tmB2, shape2, epi, sms, st, "debug_gemm2", KC_OTHER);
        }
        set_error("debug gemm: the 2-CTA variant is instantiated for equal operand layouts only");
        return RBN_E_BADARG;
    }
    seg &= 1;
#define RBN_DBG(AM, BM_, SG) return launch_tc_gemm<AM, BM_, 256, 4, SG, EpiStoreF32>(tmA, tmB, shape, epi, sms, st, "debug_gemm", KC_OTHER)
    if (seg) {
        if (!a_mn && !b_mn) RBN_DBG(false, false, kSegBlocks);
        if (a_mn && b_mn) RBN_DBG(true, true, kSegBlocks);
        if (a_mn) RBN_DBG(true, false, kSegBlocks);
        RBN_DBG(false, true, kSegBlocks);
    }
    if (!a_mn && !b_mn) RBN_DBG(false, false, 0);
    if (a_mn && b_mn) RBN_DBG(true, true, 0);
    if (a_mn) RBN_DBG(true, false, 0);
    RBN_DBG(false, true, 0);
#undef RBN_DBG
}

}  // namespace rbn
