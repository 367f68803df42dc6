// Shared definitions for librbn_b200: error handling, chart indexing, workspace layout.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "../../include/rbn_b200.h"

namespace rbn {

// ---- error handling (C ABI never throws) --------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define RBN_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess) {                                                                          \
            rbn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);   \
            return RBN_E_CUDA;                                                                            \
        }                                                                                                 \
    } while (0)
#define RBN_LAUNCH_CHECK(name)                                                                            \
    do {                                                                                                  \
        cudaError_t _e = cudaGetLastError();                                                              \
        if (_e != cudaSuccess) {                                                                          \
            rbn::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));                      \
            return RBN_E_CUDA;                                                                            \
        }                                                                                                 \
    } while (0)

// ---- chart indexing ---------------------------------------------------------------------------------------------
// Span-major ("diagonal-major") compact chart: all cells of width 1, then width 2, ...  Within a sequence the cell
// of span (i, i+w) sits at cell_off(w, L) + i; a sequence has L(L+1)/2 cells.  The reference's root-first TMap
// order (pcfg.py:215, test_pcfg.py:14-19) is produced on export only.
__host__ __device__ inline int cell_off(int w, int L) { return (w - 1) * (L + 1) - ((w - 1) * w) / 2; }
__host__ __device__ inline int cells_per_seq(int L) { return L * (L + 1) / 2; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- workspace layout -------------------------------------------------------------------------------------------
struct WsLayout {
    // common (both paths)
    size_t chart_val;   // float  [B][C][N]   mantissas, largest entry of a cell in [0.5, 1) (power-of-two scaled => exact)
    size_t chart_exp;   // int32  [B][C]      beta = mantissa * 2^exp
    size_t alpha;       // float  scaled outside values (backward only): CUDA-core path [B][C][N]; tensor-core path
                        //        start-major [B][L][Lp][N] (root seed + left- and right-child sums), cell (i,j) at row (b*L+i)*Lp + j
    size_t zroot;       // float  [B]
    size_t rootexp;     // int32  [B]
    // tensor-core path
    size_t lc;          // half [B][L][Lp][N]    start-major operand copy: cell (i, j) at row (b*L + i)*Lp + j
    size_t rc;          // half [B][L+1][Lp][N]  end-major operand copy:   cell (j, k) at row (b*(L+1) + k)*Lp + j
    size_t wt;          // half [N][N*N]         Wt[a][b*N+c] = W[b][c][a] * 2^-wexp[a]
    size_t wk;          // half [N*N][N]         Wk[b*N+c][a] = W[b][c][a] * 2^-wexp[a]
    size_t wexp;        // int32 [N]   W = W16 * 2^wexp[a]
    size_t wscale;      // float [2][N]  2^wexp[a]; the same times (1 + truncation compensation of a whole-tile chain)
    size_t wdens;       // float [N]   fraction of nonzero rules of parent a
    size_t wbits;       // int32 [2][N]  scratch: float bits of max_{b,c} W[b][c][a]; number of nonzero rules
    // per-chunk buffers: `nslot` copies each (slot s at offset + s * chunk * row bytes), so that chunks in flight on
    // the two launch streams of the overlapped schedule never share a buffer
    size_t y;           // half [nslot][chunk][N*N]
    size_t gy;          // half [nslot][chunk][N*N]
    size_t item_e;      // int32 [nslot][chunk]         E_m = max_j (e_ij + e_jk)
    size_t ghat;        // half [nslot][chunk][N]       item-normalised upstream gradient
    size_t gabs;        // half [nslot][chunk][N]       upstream gradient under one scale per launch (operand of the count GEMM)
    size_t g32;         // float [nslot][chunk][N]      upstream gradient G, fp32 (k_prep_g -> k_gabs_convert)
    size_t gmax;        // int32 [16] running max |G| (float bits) per scale group + float [16] the groups' inverse scales
    size_t gexp;        // int32 [nslot][chunk]
    size_t acc;         // float [nslot][chunk][N]  K-split partial sums of the rule GEMM's tail tiles
    // spans wider than 129 tokens (more than 128 splits): partial split-sum of one window of splits, before k_y_combine
    size_t ywin;        // half [wide_items][N*N]   (wide_items = most spans a unit of a diagonal w > 129 can have; 0 if L <= 129)
    size_t ywin_e;      // int32 [wide_items]
    // split-sum cache (RbnDesc.cache_bytes): the inside pass writes the Y of `cache_spans` spans here (widest diagonals
    // first) and the outside pass reads them back instead of recomputing them; E_m of the same spans beside it
    size_t ycache;      // half [cache_spans][N*N]
    size_t ycache_e;    // int32 [cache_spans]
    size_t gabs_all;    // half [cache_spans][N]  upstream-gradient rows of the kept spans (the count GEMM over them runs once, last)
    size_t g32_all;     // float [cache_spans][N] the same rows in fp32, before their common scale is known
    long long cache_spans;
    size_t total;
    int Lp;
    int chunk;
    int nslot;
};

// ---- schedule of the tensor-core path (read once from the environment) -----------------------
// overlap = 1: two launch streams with disjoint SM partitions.  Write-bound kernels (split-sum, gY GEMM) run beside
// read-/tensor-bound ones (rule GEMM, count GEMM, child gradients) of ANOTHER group of sequences, so that both
// directions of the HBM interface and the tensor pipe are busy at the same time.
struct TcTune {
    int overlap;        // RBN_TC_OVERLAP   0: one stream; 1: two streams, disjoint SM partitions; 2: two streams, co-resident
                        //                  'half SM' kernel variants (inside pass); 3: two streams, ordinary kernels (tail filling)
    int groups;         // RBN_TC_GROUPS    sequence groups that pipeline against each other (default 2)
    int fwd_sms0;       // RBN_TC_FWD_SMS0  SMs of the split-sum partition in the inside pass (even)
    int bwd_sms0;       // RBN_TC_BWD_SMS0  SMs of the {split-sum, prep, gY} partition in the outside pass (even)
    int nslot;          // RBN_TC_SLOTS     per-chunk buffer copies
    int k3_stream;      // RBN_TC_K3_STREAM 0: gY GEMM beside the split-sum, 1: beside the child-gradient kernel
    int pdl;            // RBN_PDL          1: the per-diagonal kernels of the tensor-core path are launched with programmatic stream
                        //                  serialisation (tc_ptx.cuh grid_dep_sync): a kernel's prologue overlaps its predecessor's tail
    float kappa;        // RBN_TC_KAPPA     tcgen05 accumulator-truncation bias per k-block of a chain (compensated in the rule GEMM)
    int hints;          // RBN_TC_HINTS     L2 eviction-priority hints on the bulk tensor copies (bit mask): 1 outside-chart
                        //                  reduce-adds evict_last (the second contribution to a cell of a diagonal should find the
                        //                  line in L2); 2 streamed stores (Y, gY) evict_first; 4 rule / count GEMM operand streams
                        //                  (Y evict_first, W evict_last); 8 second read of gY in the child-gradient kernel evict_first
};
const TcTune& tc_tune();
int device_sm_count();

// kern<<<grid, block, smem, st>>>(args...) with, under RBN_PDL, the programmatic-stream-serialisation attribute.  Only for
// kernels that call ptx::grid_dep_sync() before they touch global memory.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    if (tc_tune().pdl) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- per-kernel-class launch accounting (always) and CUDA-event timing (when enabled) ------------------------------
enum KernelClass { KC_EMISSION = 0, KC_SPLITSUM, KC_RULE, KC_GY, KC_COUNTS, KC_CHILD, KC_PREP, KC_SIMT_INSIDE,
                   KC_SIMT_OUTSIDE, KC_OTHER, KC_RULE_FIN, KC_NUM };
struct ProfScope {          // brackets ONE kernel launch on `st`
    int cls;
    cudaStream_t st;
    int slot;
    ProfScope(int cls, cudaStream_t st);
    ~ProfScope();
};

int resolve_path(const RbnDesc* d);
bool make_layout(const RbnDesc* d, WsLayout* out);

}  // namespace rbn
