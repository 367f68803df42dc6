/*
 * rbn_b200.h -- C ABI of the B200-native inside / outside chart pass (librbn_b200.so).
 *
 * This is the drop-in boundary for ONE path of robert-lieck/RBN (`rbnet`): RBN.inside()
 * (/root/reference/rbnet/base.py:93-121) with the discrete bricks of /root/reference/rbnet/pcfg.py, and the
 * outside pass that the reference obtains implicitly from torch.autograd.
 * The reference has no FFI (it is pure Python); the Python binding a maintainer adds is shown in INTEGRATION.md
 * and shipped as rbn_b200/_lib.py (ctypes).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; every pointer except RbnDesc* / const char* is a DEVICE pointer.
 *   - The caller allocates everything (including the workspace) and passes a CUDA stream handle (cudaStream_t
 *     as void*; NULL = legacy default stream).  The library never allocates or frees device memory, keeps no
 *     pointer after return and never synchronises the host: all work is stream-ordered.
 *   - Return value 0 = success, negative = error (see RBN_E_*); rbn_last_error() gives a thread-local message.
 *     Numeric "ungrammatical input" is not an error: log Z = -inf, Z = 0 (tests/test_base.py:293 in the reference).
 *
 * Grammar representation ("flat grammar"): a multi-cell RBN is collapsed host-side into one block-structured
 * grammar with the structural probabilities folded in (rbn_b200/engine.py, restated in oracle/inside_oracle.py
 * flatten_bricks):
 *   W     [N][N][N] float32  W[b][c][a] = p(left=b, right=c | parent=a) * p_S(binary | a)   -- parent LAST,
 *                            as DiscreteBinaryNonTerminalTransition stores it (pcfg.py:91-93, 289-290)
 *   T     [V][N]    float32  T[y][a]    = p(y | a) * p_S(terminal | a)                       (pcfg.py:339, 351)
 *   prior [N]       float32  prior[a]   = p_S(var) * p(a)                                    (pcfg.py:270-273)
 *   tokens  [B][L][n_tvars] int32, index into the rows of T, -1 = None (pcfg.py:346-349) / padding
 *   lengths [B] int32, 1 <= lengths[b] <= L; the root of sequence b is span (0, lengths[b]) (sequential.py:28-30)
 */
#ifndef RBN_B200_H
#define RBN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBN_ABI_VERSION 3

/* error codes */
#define RBN_OK            0
#define RBN_E_BADARG     -1   /* null pointer / non-positive size / unsupported combination */
#define RBN_E_WORKSPACE  -2   /* workspace smaller than rbn_workspace_bytes() */
#define RBN_E_CUDA       -3   /* a CUDA call or kernel launch failed (message in rbn_last_error) */
#define RBN_E_UNSUPPORTED -4  /* shape outside what the selected path supports */

/* RbnDesc.path */
#define RBN_PATH_AUTO     0   /* tensor-core path when N is 64, 128, 192, 256 or 512 and 2 <= L <= 1025, else the shared-memory path */
#define RBN_PATH_SIMT     1   /* shared-memory / CUDA-core path, fp32 end to end (small nonterminal counts) */
#define RBN_PATH_TCGEN05  2   /* span-diagonal tcgen05 tensor-core path (fp16 operands, fp32 accumulate); L <= 1025 (spans wider
                                 than 129 tokens: one split-sum / child-gradient launch per window of 128 splits); longer
                                 inputs: RBN_E_UNSUPPORTED here, the shared-memory path under AUTO */

typedef struct RbnDesc {
    int32_t N;            /* number of (flattened) nonterminal values */
    int32_t V;            /* number of rows of T (flattened terminal alphabet) */
    int32_t B;            /* sequences in the batch */
    int32_t L;            /* padded sequence length (max over the batch) */
    int32_t n_tvars;      /* terminal variables per position (AbstractedPCFG: 1) */
    int32_t path;         /* RBN_PATH_* */
    int32_t chunk_items;  /* tensor-core path: spans processed per launch group (0 = default) */
    int32_t flags;        /* reserved, must be 0 */
    int64_t cache_bytes;  /* tensor-core path: bytes of EXTRA workspace the inside pass may fill with the per-span
                             split-sums Y (N*N fp16 each) so that the outside pass reads them back instead of
                             recomputing them (widest diagonals first; clamped to what the batch can use).
                             0 = keep nothing.  rbn_workspace_bytes() includes it; forward and backward must be
                             given the same value.  Pure speed/space trade: results do not depend on it. */
} RbnDesc;

/* ABI version of the loaded library (== RBN_ABI_VERSION of the header it was built from). */
int rbn_abi_version(void);

/* Thread-local description of the last error returned on this thread ("" if none). */
const char* rbn_last_error(void);

/* Which path RBN_PATH_AUTO resolves to for this descriptor (RBN_PATH_SIMT or RBN_PATH_TCGEN05), or <0. */
int rbn_resolve_path(const RbnDesc* desc);

/* Bytes of caller-allocated device workspace for one forward (+ its backward).  The workspace carries the
 * chart between rbn_inside_forward and rbn_inside_backward / rbn_export_chart; distinct concurrent passes need
 * distinct workspaces.  Base address must be 1024-byte aligned.  Returns 0 on a bad descriptor. */
size_t rbn_workspace_bytes(const RbnDesc* desc);

/* Inside pass.  Replaces the loop nest of RBN.inside (base.py:93-121): the width-major schedule
 * (sequential.py:23-26), DiscreteTerminalTransition.inside_marginals (pcfg.py:344-352),
 * DiscreteBinaryNonTerminalTransition.inside_marginals (pcfg.py:305-314), DiscreteCell.inside_mixture
 * (pcfg.py:387-398) and DiscretePrior.marginal_likelihood (pcfg.py:270-273), batched over B sequences.
 * Outputs (device, float32 [B] each):
 *   logZ      natural-log marginal likelihood, -inf when Z == 0
 *   zroot     sum_a prior[a] * mantissa(root)[a]
 *   rootexp   power-of-two exponent of the root cell:  Z = zroot * 2^rootexp  (exact linear-space value)
 */
int rbn_inside_forward(const RbnDesc* desc,
                       const float* W, const float* T, const float* prior,
                       const int32_t* tokens, const int32_t* lengths,
                       void* workspace, size_t workspace_bytes,
                       float* logZ, float* zroot, float* rootexp,
                       void* stream);

/* Outside pass / expected rule counts.  Replaces the torch.autograd traversal of the reference's chart writes
 * (sequential.py:36 CopySlices chain).  Must follow rbn_inside_forward on the same
 * workspace and inputs.  With R_b = rootexp[b] + log2(zroot[b]) (or rootexp[b] when Z_b == 0) it accumulates
 *   gW[b][c][a] = sum_seq coef[seq] * 2^-R_seq * dZ_seq/dW[b][c][a]      (likewise gT [V][N], gPrior [N])
 * i.e. coef = dLoss/dlogZ gives dLoss/dW for grammatical sequences.  gW/gT/gPrior are OVERWRITTEN. */
int rbn_inside_backward(const RbnDesc* desc,
                        const float* W, const float* T, const float* prior,
                        const int32_t* tokens, const int32_t* lengths,
                        void* workspace, size_t workspace_bytes,
                        const float* coef,
                        float* gW, float* gT, float* gPrior,
                        void* stream);

/* Ragged batches.  The same two passes with one more argument: a HOST array active[0..L], active[w] = number of
 * sequences whose length is >= w.  It may only be given when the batch is sorted by length, longest first
 * (lengths[0] >= lengths[1] >= ...); diagonal w then visits only the first active[w] sequences instead of padding
 * every sequence to L (the reference parses one sequence at a time and never pads, sequential.py:23-26).
 * host_active == NULL is rbn_inside_forward / rbn_inside_backward.  Forward and backward must get the same array. */
int rbn_inside_forward_sorted(const RbnDesc* desc,
                              const float* W, const float* T, const float* prior,
                              const int32_t* tokens, const int32_t* lengths,
                              void* workspace, size_t workspace_bytes,
                              float* logZ, float* zroot, float* rootexp,
                              const int32_t* host_active, void* stream);
int rbn_inside_backward_sorted(const RbnDesc* desc,
                               const float* W, const float* T, const float* prior,
                               const int32_t* tokens, const int32_t* lengths,
                               void* workspace, size_t workspace_bytes,
                               const float* coef,
                               float* gW, float* gT, float* gPrior,
                               const int32_t* host_active, void* stream);

/* Most probable derivation (max-semiring variant of the inside pass; NOT in the reference -- SURVEY.md 8(f).4).
 * CUDA-core kernels, any N <= 1024, L <= 2047; workspace = rbn_workspace_bytes() of the same descriptor with
 * path = RBN_PATH_SIMT.  Outputs (device): logP [B] natural log of the best derivation's probability (-inf: none),
 * root_symbol [B] its root nonterminal (-1: none), backptr int32 [B][L(L+1)/2][N]: for the cell of span (i, i+w) at
 * index cell = (w-1)(L+1) - (w-1)w/2 + i and nonterminal a, (u << 20) | (b << 10) | c = best split i+u with left
 * symbol b and right symbol c, or -1 (width-1 cell / no derivation). */
int rbn_viterbi(const RbnDesc* desc,
                const float* W, const float* T, const float* prior,
                const int32_t* tokens, const int32_t* lengths,
                void* workspace, size_t workspace_bytes,
                float* logP, int32_t* backptr, int32_t* root_symbol,
                void* stream);

/* Chart of one sequence in the reference's layout: rows ordered as triangularmap.TMap orders them (root first:
 * row = d(d+1)/2 + start with d = n - (end - start); pinned by /root/reference/tests/test_pcfg.py:14-19),
 * linear probability space, float64 [n(n+1)/2][N] with n = lengths[seq].  Serves SequentialRBN.inside_chart
 * (sequential.py:38-40) and PCFG.map_inside_chart (pcfg.py:30-41). */
int rbn_export_chart(const RbnDesc* desc, const int32_t* lengths,
                     const void* workspace, size_t workspace_bytes,
                     int32_t seq, double* out, void* stream);

/* Fused constrained-parameter maintenance (the step right after the path in a training loop):
 * LogProb.enforce_constraints (util.py:193-198): log_p[g][k] -= logsumexp_k log_p[g][k] over `groups` groups of
 * `group_size` contiguous-with-stride entries.  Layout: element (g, k) at log_p[k * groups + g] (the reference
 * normalises over the LEADING dims, parent last).  Returns RBN_E_BADARG if a group's normaliser is -inf
 * is NOT detected on device; `bad` (device int32, may be NULL) is set to the number of such groups. */
int rbn_lognormalize(float* log_p, int64_t group_size, int64_t groups, int32_t* bad, void* stream);
/* The same for float64 parameters (the reference's Prob / LogProb are fp64 when built from numpy weights, util.py:128-135). */
int rbn_lognormalize_f64(double* log_p, int64_t group_size, int64_t groups, int32_t* bad, void* stream);

/* ---- continuous RBN with multivariate-normal transitions (BASELINE config 4) ------------------------------------
 * beta_{i:k}(x) = exp(logc) N(x; mu, Sigma).  Terminal p(y|x) = N(y; x, cov_term); binary transition
 * p(x_l, x_r | x) = N(x_l; x, cov_left) N(x_r; x, cov_right); per split the children are multiplied as
 * rbnet.multivariate_normal.PairwiseProduct does (/root/reference/rbnet/multivariate_normal.py:339-396), the splits
 * of a span are moment-matched as ApproximateMixture does (:502-550); prior N(x; prior_mean, prior_cov).  The
 * reference ships only these primitives, not the RBN bricks; the composition is the one of
 * oracle/gauss_oracle.py.  All arrays float64 on the device; y [B][L][D]; covariances [D][D] row-major symmetric;
 * log_struct[2] = {log p_S(terminal), log p_S(binary)}; D in {1,2,3,4,8}.
 *
 * flags: RBN_GAUSS_F32 (ABI 3, D = 8 only) selects the _f32 entry points below: observations, the chart, the adjoint
 * chart and g_y are float32 -- the reference's default dtype (torch.get_default_dtype(), multivariate_normal.py:150) --
 * and the per-split algebra runs in fp32; the log weights (log c per cell), log Z, every parameter and every parameter
 * gradient stay float64.  A descriptor and the entry point it is passed to must agree (RBN_E_BADARG otherwise). */
#define RBN_GAUSS_F32 1
typedef struct RbnGaussDesc {
    int32_t D, B, L, flags;
} RbnGaussDesc;
size_t rbn_gauss_workspace_bytes(const RbnGaussDesc* desc);
int rbn_gauss_inside_forward(const RbnGaussDesc* desc, const double* y, const int32_t* lengths, const double* cov_term,
                             const double* cov_left, const double* cov_right, const double* prior_mean,
                             const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                             double* logZ, void* stream);

/* Outside pass of the Gaussian RBN (reverse mode through the moment matching and the pairwise products; the reference
 * would obtain it from torch.autograd).  Must follow rbn_gauss_inside_forward on the same workspace.  Outputs are
 * OVERWRITTEN with d(sum_b coef[b] log Z_b)/d(argument); covariance gradients are symmetrised. g_y is [B][L][D]. */
int rbn_gauss_inside_backward(const RbnGaussDesc* desc, const double* y, const int32_t* lengths, const double* cov_term,
                              const double* cov_left, const double* cov_right, const double* prior_mean,
                              const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                              const double* coef, double* g_y, double* g_cov_term, double* g_cov_left,
                              double* g_cov_right, double* g_prior_mean, double* g_prior_cov, double* g_log_struct,
                              void* stream);
/* The same two passes over float32 observations / charts (desc->flags & RBN_GAUSS_F32; workspace sized by
 * rbn_gauss_workspace_bytes for that descriptor: 2 x B x L(L+1)/2 cells of 296 B). */
int rbn_gauss_inside_forward_f32(const RbnGaussDesc* desc, const float* y, const int32_t* lengths, const double* cov_term,
                                 const double* cov_left, const double* cov_right, const double* prior_mean,
                                 const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                                 double* logZ, void* stream);
int rbn_gauss_inside_backward_f32(const RbnGaussDesc* desc, const float* y, const int32_t* lengths, const double* cov_term,
                                  const double* cov_left, const double* cov_right, const double* prior_mean,
                                  const double* prior_cov, const double* log_struct, void* workspace, size_t workspace_bytes,
                                  const double* coef, float* g_y, double* g_cov_term, double* g_cov_left,
                                  double* g_cov_right, double* g_prior_mean, double* g_prior_cov, double* g_log_struct,
                                  void* stream);

/* Launch accounting and optional per-kernel timing.  The library counts every kernel it launches, by class; with
 * rbn_profile_enable(1) it also brackets each launch with CUDA events on the caller's stream.  rbn_profile_read
 * synchronises the device and returns, per class, the summed event time (ms) and the launch count since the last
 * reset.  Classes (RBN_KC_*): 0 emission, 1 split-sum, 2 rule GEMM, 3 gY GEMM, 4 count GEMM, 5 child-gradient,
 * 6 gradient prep, 7 CUDA-core inside, 8 CUDA-core outside, 9 other.  n must be >= RBN_KC_NUM. */
#define RBN_KC_NUM 10
int rbn_profile_enable(int32_t on);
int rbn_profile_read(double* ms, int64_t* launches, int32_t n, int32_t reset);

/* Diagnostics: D[M][N] (float32) = A * B through the same tcgen05 GEMM kernel template the chart pass uses
 * (N must be 256, K a multiple of 64; fp16 operands).  a_mn / b_mn select the operand layout: 0 = K-major
 * (A[M][K], B[N][K]), 1 = MN-major (A[K][M], B[K][N]).  segmented != 0 selects the segmented-accumulation variant
 * (K chains of 4096 folded with round-to-nearest adds; DESIGN.md "accumulator truncation").  Tests only. */
int rbn_debug_gemm(int32_t M, int32_t N, int32_t K, const void* A, const void* B, float* D, int32_t a_mn, int32_t b_mn,
                   int32_t segmented, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RBN_B200_H */
