// Shared-memory / CUDA-core path of the inside and outside passes (fp32 end to end).
//
// Serves small nonterminal counts (N not a multiple of 64, e.g. the reference's own N <= 10 grammars) and is the
// on-device cross-check for the tensor-core path.  One CTA per span (b, i, i+w); spans of one width are one
// launch ("span diagonal"), widths ascend in the inside pass and descend in the outside pass.
//
// Arithmetic restated from the reference (never copied):
//   emission   beta[i,i+1,a] = sum_t T[y_i^t, a]                       pcfg.py:344-352 (+ structural weight, folded)
//   binary     beta[i,k,a]   = sum_j sum_{b,c} W[b,c,a] beta[i,j,b] beta[j,k,c]   pcfg.py:305-314, 387-398
//   root       Z = sum_a prior[a] beta[0,n,a]                          pcfg.py:270-273
// in per-cell power-of-two scaled form: beta = m * 2^e with max_a m = [0.5, 1); an all-zero cell keeps the finite
// exponent of its best split so that derivatives w.r.t. zero-probability entries survive (oracle: flat_forward).
#include "common.cuh"
#include <cooperative_groups.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>

namespace rbn {

static constexpr int kThreads = 256;
static constexpr int kMaxAPerThread = 4;   // N <= 1024 on this path

__device__ __forceinline__ float block_max(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int k = 1; k < (blockDim.x >> 5); ++k) r = fmaxf(r, red[k]);
    return r;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) r += red[k];
    return r;
}

// exponent e with mx * 2^-e in [0.5, 1); 0 for mx == 0
__device__ __forceinline__ int norm_exp(float mx) {
    if (!(mx > 0.f)) return 0;
    int e;
    frexpf(mx, &e);
    return e;
}

// ---- width-1 cells ---------------------------------------------------------------------------------------------
__global__ void k_emission(int N, int L, int nt, const float* __restrict__ T, const int* __restrict__ tokens,
                           const int* __restrict__ lengths, float* __restrict__ cv, int* __restrict__ ce) {
    __shared__ float red[32];
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    size_t cell = (size_t)b * cells_per_seq(L) + i;
    float v[kMaxAPerThread];
    float mx = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) {
        int a = threadIdx.x + q * kThreads;
        float s = 0.f;
        if (a < N)
            for (int t = 0; t < nt; ++t) {
                int y = tokens[((size_t)b * L + i) * nt + t];
                if (y >= 0) s += T[(size_t)y * N + a];
            }
        v[q] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_max(mx, red);
    int e = norm_exp(mx);
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) {
        int a = threadIdx.x + q * kThreads;
        if (a < N) cv[cell * N + a] = ldexpf(v[q], -e);
    }
    if (threadIdx.x == 0) ce[cell] = e;
}

// split coefficients of span (i, i+w): c_u = 2^(e_left(u) + e_right(u) - E), E = max_u (...)
__device__ __forceinline__ int split_coefs(int L, int w, int i, const int* ce_seq, float* cj, int* sE) {
    if (threadIdx.x == 0) *sE = INT_MIN;
    __syncthreads();
    int mine = INT_MIN;
    for (int u = 1 + threadIdx.x; u < w; u += blockDim.x) {
        int s = ce_seq[cell_off(u, L) + i] + ce_seq[cell_off(w - u, L) + i + u];
        mine = max(mine, s);
    }
    if (mine != INT_MIN) atomicMax(sE, mine);
    __syncthreads();
    int E = *sE;
    for (int u = 1 + threadIdx.x; u < w; u += blockDim.x) {
        int s = ce_seq[cell_off(u, L) + i] + ce_seq[cell_off(w - u, L) + i + u];
        cj[u] = ldexpf(1.0f, max(s - E, -200));
    }
    __syncthreads();
    return E;
}

// Y[bb][c] = sum_u c_u * left_u[b0+bb] * right_u[c]   (the split-sum, taken before the rule contraction)
__device__ __forceinline__ void split_sum_block(int N, int L, int w, int i, int b0, int tb,
                                                const float* cv_seq, const float* cj, float* Y) {
    for (int idx = threadIdx.x; idx < tb * N; idx += blockDim.x) {
        int bb = idx / N, c = idx - bb * N;
        float acc = 0.f;
        for (int u = 1; u < w; ++u) {
            const float* l = cv_seq + (size_t)(cell_off(u, L) + i) * N;
            const float* r = cv_seq + (size_t)(cell_off(w - u, L) + i + u) * N;
            acc += cj[u] * l[b0 + bb] * r[c];
        }
        Y[idx] = acc;
    }
}

// ---- inside: one span diagonal ----------------------------------------------------------------------------------
// body of one span (CTA-uniform control flow); `blk` = b * (L - w + 1) + i
__device__ __forceinline__ void inside_span(int N, int L, int w, int TB, const float* __restrict__ W,
                                            const int* __restrict__ lengths, float* cv, int* ce, int blk, float* sm) {
    __shared__ float red[32];
    __shared__ int sE;
    int nI = L - w + 1;
    int b = blk / nI, i = blk % nI;
    if (i + w > lengths[b]) return;
    size_t seq = (size_t)b * cells_per_seq(L);
    const float* cv_seq = cv + seq * N;
    const int* ce_seq = ce + seq;
    float* cj = sm;               // [L+1]
    float* Y = sm + (L + 1);      // [TB][N]
    float* oacc = Y + (size_t)TB * N;   // [N]  (small N only)
    int E = split_coefs(L, w, i, ce_seq, cj, &sE);

    // small grammars: kThreads / N groups of N threads share the (b, c) range, partial sums meet in shared memory
    const bool grouped = (N * 4 <= kThreads);
    const int ng = grouped ? kThreads / N : 1;
    const int grp = threadIdx.x / N, ga = threadIdx.x - grp * N;
    if (grouped && threadIdx.x < N) oacc[threadIdx.x] = 0.f;

    float out[kMaxAPerThread];
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) out[q] = 0.f;
    for (int b0 = 0; b0 < N; b0 += TB) {
        int tb = min(TB, N - b0);
        split_sum_block(N, L, w, i, b0, tb, cv_seq, cj, Y);
        __syncthreads();
        if (grouped) {
            if (grp < ng) {
                float acc = 0.f;
                const float* Wp = W + (size_t)b0 * N * N + ga;
                for (int idx = grp; idx < tb * N; idx += ng) acc = fmaf(Y[idx], Wp[(size_t)idx * N], acc);
                atomicAdd(&oacc[ga], acc);
            }
        } else {
#pragma unroll
            for (int q = 0; q < kMaxAPerThread; ++q) {
                int a = threadIdx.x + q * kThreads;
                if (a < N) {
                    float acc = out[q];
                    const float* Wp = W + (size_t)b0 * N * N + a;
                    for (int idx = 0; idx < tb * N; ++idx) acc = fmaf(Y[idx], Wp[(size_t)idx * N], acc);
                    out[q] = acc;
                }
            }
        }
        __syncthreads();
    }
    if (grouped && threadIdx.x < N) out[0] = oacc[threadIdx.x];
    float mx = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) mx = fmaxf(mx, out[q]);
    mx = block_max(mx, red);
    int e = norm_exp(mx);
    size_t cell = seq + cell_off(w, L) + i;
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) {
        int a = threadIdx.x + q * kThreads;
        if (a < N) cv[cell * N + a] = ldexpf(out[q], -e);
    }
    if (threadIdx.x == 0) ce[cell] = E + e;
}

__global__ void k_inside_diag(int N, int L, int w, int TB, const float* __restrict__ W,
                              const int* __restrict__ lengths, float* cv, int* ce) {
    extern __shared__ float sm[];
    inside_span(N, L, w, TB, W, lengths, cv, ce, blockIdx.x, sm);
}

// Small problems (a handful of short sequences, the reference's own use): ALL diagonals in one cooperative launch,
// grid-wide barrier between widths instead of L - 1 kernel launches.
__global__ void k_inside_all(int N, int L, int B, int TB, const float* __restrict__ W, const int* __restrict__ lengths,
                             float* cv, int* ce) {
    extern __shared__ float sm[];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int w = 2; w <= L; ++w) {
        const int items = B * (L - w + 1);
        for (int blk = blockIdx.x; blk < items; blk += gridDim.x) {
            inside_span(N, L, w, TB, W, lengths, cv, ce, blk, sm);
            __syncthreads();
        }
        grid.sync();
    }
}

// ---- Viterbi (max-semiring) variant of the inside pass -------------------------------------------------------------
// best[i,k,a] = max_{j,b,c} W[b,c,a] best[i,j,b] best[j,k,c]  with the same per-cell power-of-two scaling; the
// argmax (split, left, right) is packed as (u << 20) | (b << 10) | c into back[cell][a] (-1: no derivation / width 1).
// Not part of the reference; needs N <= 1024 and L <= 2048 for the packing.
__global__ void k_viterbi_diag(int N, int L, int w, int TB, const float* __restrict__ W, const int* __restrict__ lengths,
                               float* __restrict__ cv, int* __restrict__ ce, int* __restrict__ back) {
    extern __shared__ float sm[];
    __shared__ float red[32];
    __shared__ int sE;
    const int nI = L - w + 1;
    const int b = blockIdx.x / nI, i = blockIdx.x % nI;
    if (i + w > lengths[b]) return;
    const size_t seq = (size_t)b * cells_per_seq(L);
    const float* cv_seq = cv + seq * N;
    const int* ce_seq = ce + seq;
    float* cj = sm;                          // [L+1]
    float* Y = sm + (L + 1);                 // [TB][N]  max over splits
    int* Yu = (int*)(Y + (size_t)TB * N);    // [TB][N]  arg max split
    const int E = split_coefs(L, w, i, ce_seq, cj, &sE);

    float best[kMaxAPerThread];
    int arg[kMaxAPerThread];
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) { best[q] = 0.f; arg[q] = -1; }
    for (int b0 = 0; b0 < N; b0 += TB) {
        const int tb = min(TB, N - b0);
        for (int idx = threadIdx.x; idx < tb * N; idx += blockDim.x) {
            const int bb = idx / N, c = idx - bb * N;
            float m = 0.f;
            int mu = -1;
            for (int u = 1; u < w; ++u) {
                const float* l = cv_seq + (size_t)(cell_off(u, L) + i) * N;
                const float* r = cv_seq + (size_t)(cell_off(w - u, L) + i + u) * N;
                const float v = cj[u] * l[b0 + bb] * r[c];
                if (v > m) { m = v; mu = u; }
            }
            Y[idx] = m;
            Yu[idx] = mu;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kMaxAPerThread; ++q) {
            const int a = threadIdx.x + q * kThreads;
            if (a < N) {
                const float* Wp = W + (size_t)b0 * N * N + a;
                for (int idx = 0; idx < tb * N; ++idx) {
                    const float v = Y[idx] * Wp[(size_t)idx * N];
                    if (v > best[q]) {
                        best[q] = v;
                        const int bb = idx / N, c = idx - bb * N;
                        arg[q] = (Yu[idx] << 20) | ((b0 + bb) << 10) | c;
                    }
                }
            }
        }
        __syncthreads();
    }
    float mx = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) mx = fmaxf(mx, best[q]);
    mx = block_max(mx, red);
    const int e = norm_exp(mx);
    const size_t cell = seq + cell_off(w, L) + i;
#pragma unroll
    for (int q = 0; q < kMaxAPerThread; ++q) {
        const int a = threadIdx.x + q * kThreads;
        if (a < N) {
            cv[cell * N + a] = ldexpf(best[q], -e);
            back[cell * N + a] = best[q] > 0.f ? arg[q] : -1;
        }
    }
    if (threadIdx.x == 0) ce[cell] = E + e;
}

// best root symbol and the natural log of the best derivation's probability (-inf if there is none)
__global__ void k_viterbi_root(int N, int L, const float* __restrict__ prior, const int* __restrict__ lengths,
                               const float* __restrict__ cv, const int* __restrict__ ce, float* __restrict__ logP,
                               int* __restrict__ root_sym) {
    __shared__ float sv[kThreads];
    __shared__ int sa[kThreads];
    const int b = blockIdx.x;
    const int n = lengths[b];
    const size_t cell = (size_t)b * cells_per_seq(L) + cell_off(n, L);
    float m = 0.f;
    int ma = -1;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        const float v = prior[a] * cv[cell * N + a];
        if (v > m) { m = v; ma = a; }
    }
    sv[threadIdx.x] = m;
    sa[threadIdx.x] = ma;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < blockDim.x; ++t)
            if (sv[t] > m || (sv[t] == m && sa[t] >= 0 && (ma < 0 || sa[t] < ma))) { m = sv[t]; ma = sa[t]; }
        logP[b] = (m > 0.f) ? logf(m) + (float)ce[cell] * 0.69314718055994531f : -INFINITY;
        root_sym[b] = ma;
    }
}

// ---- root / prior -----------------------------------------------------------------------------------------------
__global__ void k_root(int N, int L, const float* __restrict__ prior, const int* __restrict__ lengths,
                       const float* __restrict__ cv, const int* __restrict__ ce,
                       float* __restrict__ logZ, float* __restrict__ zroot, float* __restrict__ rootexp,
                       float* __restrict__ ws_zroot, int* __restrict__ ws_rootexp) {
    __shared__ float red[32];
    int b = blockIdx.x;
    int n = lengths[b];
    size_t cell = (size_t)b * cells_per_seq(L) + cell_off(n, L);
    float s = 0.f;
    for (int a = threadIdx.x; a < N; a += blockDim.x) s += prior[a] * cv[cell * N + a];
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        int e = ce[cell];
        zroot[b] = s;
        rootexp[b] = (float)e;
        ws_zroot[b] = s;
        ws_rootexp[b] = e;
        logZ[b] = (s > 0.f) ? (logf(s) + (float)e * 0.69314718055994530942f) : -INFINITY;
    }
}

// ---- outside ------------------------------------------------------------------------------------------------------
// scaled adjoint: alpha[cell][a] = dZ/dbeta[cell][a] * 2^(e_cell - R_seq) * coef_seq   ("scaled outside")
__global__ void k_outside_seed(int N, int L, const float* __restrict__ prior, const int* __restrict__ lengths,
                               const float* __restrict__ cv, const float* __restrict__ ws_zroot,
                               const float* __restrict__ coef, float* __restrict__ alpha,
                               float* __restrict__ gPrior) {
    int b = blockIdx.x;
    int n = lengths[b];
    size_t cell = (size_t)b * cells_per_seq(L) + cell_off(n, L);
    float z = ws_zroot[b];
    float f = coef[b] / (z > 0.f ? z : 1.0f);
    if (f == 0.f) return;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        alpha[cell * N + a] = f * prior[a];
        float m = cv[cell * N + a];
        if (m != 0.f) atomicAdd(&gPrior[a], f * m);
    }
}

__device__ __forceinline__ void outside_span(int N, int L, int w, int TB, const float* __restrict__ W,
                                             const int* __restrict__ lengths, const float* __restrict__ cv,
                                             const int* __restrict__ ce, float* alpha, float* gW, int blk, float* sm) {
    __shared__ int sE;
    __shared__ int sAny;
    int nI = L - w + 1;
    int b = blk / nI, i = blk % nI;
    if (i + w > lengths[b]) return;
    size_t seq = (size_t)b * cells_per_seq(L);
    const float* cv_seq = cv + seq * N;
    const int* ce_seq = ce + seq;
    float* al_seq = alpha + seq * N;
    float* cj = sm;                     // [L+1]
    float* G = cj + (L + 1);            // [N]
    float* Y = G + N;                   // [TB][N]
    float* gY = Y + (size_t)TB * N;     // [TB][N]
    int cellw = cell_off(w, L) + i;
    int E = split_coefs(L, w, i, ce_seq, cj, &sE);
    int e_new = ce_seq[cellw] - E;
    if (threadIdx.x == 0) sAny = 0;
    __syncthreads();
    int any = 0;
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        float g = ldexpf(al_seq[(size_t)cellw * N + a], -e_new);
        G[a] = g;
        any |= (g != 0.f);
    }
    if (any) sAny = 1;
    __syncthreads();
    if (!sAny) return;   // span not reachable from the root (or coef == 0): contributes nothing

    for (int b0 = 0; b0 < N; b0 += TB) {
        int tb = min(TB, N - b0);
        split_sum_block(N, L, w, i, b0, tb, cv_seq, cj, Y);
        // gY[bb][c] = sum_a W[b0+bb][c][a] G[a]
        for (int idx = threadIdx.x; idx < tb * N; idx += blockDim.x) {
            const float* Wp = W + ((size_t)b0 * N + idx) * N;
            float acc = 0.f;
            for (int a = 0; a < N; ++a) acc = fmaf(Wp[a], G[a], acc);
            gY[idx] = acc;
        }
        __syncthreads();
        // expected rule counts: gW[b][c][a] += Y[b][c] * G[a]   (all threads busy for any N: flat (b, c, a) index, a fastest)
        for (int idx2 = threadIdx.x; idx2 < tb * N * N; idx2 += blockDim.x) {
            const int idx = idx2 / N, a = idx2 - idx * N;
            const float v = Y[idx] * G[a];
            if (v != 0.f) atomicAdd(&gW[((size_t)b0 * N + idx) * N + a], v);
        }
        // children: alpha[i,j][b] += c_u sum_c gY[b][c] right_u[c];  alpha[j,k][c] += c_u sum_b gY[b][c] left_u[b]
        // flat (split, nonterminal) index so that short grammars still use every thread
        for (int idx = threadIdx.x; idx < (w - 1) * tb; idx += blockDim.x) {
            const int u = 1 + idx / tb, bb = idx - (u - 1) * tb;
            const float cu = cj[u];
            if (cu == 0.f) continue;
            const int lcell = cell_off(u, L) + i, rcell = cell_off(w - u, L) + i + u;
            const float* r = cv_seq + (size_t)rcell * N;
            float acc = 0.f;
            for (int c = 0; c < N; ++c) acc = fmaf(gY[bb * N + c], r[c], acc);
            if (acc != 0.f) atomicAdd(&al_seq[(size_t)lcell * N + b0 + bb], cu * acc);
        }
        for (int idx = threadIdx.x; idx < (w - 1) * N; idx += blockDim.x) {
            const int u = 1 + idx / N, c = idx - (u - 1) * N;
            const float cu = cj[u];
            if (cu == 0.f) continue;
            const int lcell = cell_off(u, L) + i, rcell = cell_off(w - u, L) + i + u;
            const float* l = cv_seq + (size_t)lcell * N;
            float acc = 0.f;
            for (int bb = 0; bb < tb; ++bb) acc = fmaf(gY[bb * N + c], l[b0 + bb], acc);
            if (acc != 0.f) atomicAdd(&al_seq[(size_t)rcell * N + c], cu * acc);
        }
        __syncthreads();
    }
}

__global__ void k_outside_diag(int N, int L, int w, int TB, const float* __restrict__ W,
                               const int* __restrict__ lengths, const float* __restrict__ cv,
                               const int* __restrict__ ce, float* alpha, float* gW) {
    extern __shared__ float sm[];
    outside_span(N, L, w, TB, W, lengths, cv, ce, alpha, gW, blockIdx.x, sm);
}

__global__ void k_outside_all(int N, int L, int B, int TB, const float* __restrict__ W, const int* __restrict__ lengths,
                              const float* __restrict__ cv, const int* __restrict__ ce, float* alpha, float* gW) {
    extern __shared__ float sm[];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int w = L; w >= 2; --w) {
        const int items = B * (L - w + 1);
        for (int blk = blockIdx.x; blk < items; blk += gridDim.x) {
            outside_span(N, L, w, TB, W, lengths, cv, ce, alpha, gW, blk, sm);
            __syncthreads();
        }
        grid.sync();
    }
}

__global__ void k_emission_bwd(int N, int L, int nt, const int* __restrict__ tokens, const int* __restrict__ lengths,
                               const int* __restrict__ ce, const float* __restrict__ alpha,
                               const float* __restrict__ alpha2, float* __restrict__ gT) {
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    size_t cell = (size_t)b * cells_per_seq(L) + i;
    int e = ce[cell];
    for (int a = threadIdx.x; a < N; a += blockDim.x) {
        float g = ldexpf(alpha[cell * N + a] + (alpha2 ? alpha2[cell * N + a] : 0.f), -e);
        if (g == 0.f) continue;
        for (int t = 0; t < nt; ++t) {
            int y = tokens[((size_t)b * L + i) * nt + t];
            if (y >= 0) atomicAdd(&gT[(size_t)y * N + a], g);
        }
    }
}

// ---- chart export in the reference's TMap order, linear space, float64 ---------------------------------------------
__global__ void k_export_chart(int N, int L, const int* __restrict__ lengths, int seq, const float* __restrict__ cv,
                               const int* __restrict__ ce, double* __restrict__ out) {
    // blockIdx.x = row in TMap order: row = d(d+1)/2 + start, d = n - width
    int row = blockIdx.x;
    int n = lengths[seq];
    if (row >= n * (n + 1) / 2) return;
    int d = (int)((sqrtf(8.f * row + 1.f) - 1.f) * 0.5f);
    while ((d + 1) * (d + 2) / 2 <= row) ++d;
    while (d * (d + 1) / 2 > row) --d;
    int start = row - d * (d + 1) / 2;
    int w = n - d;
    size_t cell = (size_t)seq * cells_per_seq(L) + cell_off(w, L) + start;
    int e = ce[cell];
    for (int a = threadIdx.x; a < N; a += blockDim.x) out[(size_t)row * N + a] = ldexp((double)cv[cell * N + a], e);
}

// ---- log-space renormalisation (LogProb.enforce_constraints, util.py:193-198) --------------------------------------
template <class F>      // float (fast exp) or double (the reference's parameters are fp64)
__global__ void k_lognormalize(F* __restrict__ lp, long long gs, long long groups, int* __restrict__ bad) {
    // one thread per group; element (g, k) at lp[k * groups + g] => coalesced across the warp
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    F mx = -INFINITY;
    for (long long k = 0; k < gs; ++k) mx = lp[k * groups + g] > mx ? lp[k * groups + g] : mx;
    if (mx == -INFINITY) {
        if (bad) atomicAdd(bad, 1);
        return;
    }
    F s = 0;
    for (long long k = 0; k < gs; ++k) {
        if (sizeof(F) == 4) s += __expf((float)(lp[k * groups + g] - mx));
        else s += exp((double)(lp[k * groups + g] - mx));
    }
    const F lse = mx + (sizeof(F) == 4 ? (F)logf((float)s) : (F)log((double)s));
    for (long long k = 0; k < gs; ++k) lp[k * groups + g] -= lse;
}

// ---- host-side drivers ----------------------------------------------------------------------------------------------
// Grid of the single-launch (cooperative) kernels, or 0 when the problem is not small: the widest diagonal must fit
// the co-resident CTAs with at most 4 spans per CTA (beyond that per-diagonal launches keep all SMs busier than a
// fixed grid with barriers would).  RBN_SIMT_COOP=0 disables the path.
static int coop_grid_for(const void* kern, size_t smem, int widest_items) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("RBN_SIMT_COOP");
        enabled = (e && e[0] == '0') ? 0 : 1;
        int dev = 0, coop = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop)
            enabled = 0;
    }
    if (!enabled || widest_items < 1) return 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return 0; }
    const int resident = per_sm * device_sm_count();
    if (widest_items > 4 * resident) return 0;
    return widest_items < resident ? widest_items : resident;
}

static int pick_tb(int N, int n_arrays) {
    // rows of Y (and gY) kept in shared memory at once; <= 64 KB in total for the tiles
    int tb = (64 * 1024 / 4) / (N * n_arrays);
    if (tb < 1) tb = 1;
    if (tb > N) tb = N;
    return tb;
}

int simt_forward(const RbnDesc* d, const WsLayout& ws, char* base, const float* W, const float* T, const float* prior,
                 const int* tokens, const int* lengths, float* logZ, float* zroot, float* rootexp, const int* active,
                 cudaStream_t st) {
    if (d->N > kThreads * kMaxAPerThread) {
        set_error("shared-memory path supports N <= %d (got %d)", kThreads * kMaxAPerThread, d->N);
        return RBN_E_UNSUPPORTED;
    }
    float* cv = (float*)(base + ws.chart_val);
    int* ce = (int*)(base + ws.chart_exp);
    int N = d->N, L = d->L, B = d->B;
    { ProfScope ps(KC_EMISSION, st); k_emission<<<B * L, kThreads, 0, st>>>(N, L, d->n_tvars, T, tokens, lengths, cv, ce); }
    RBN_LAUNCH_CHECK("k_emission");
    int TB = pick_tb(N, 1);
    size_t smem = ((size_t)(L + 1) + (size_t)TB * N + N) * sizeof(float);
    RBN_CUDA_CHECK(cudaFuncSetAttribute(k_inside_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int coop_grid = coop_grid_for((const void*)k_inside_all, smem, B * (L - 1));
    if (coop_grid > 0 && L >= 2) {
        RBN_CUDA_CHECK(cudaFuncSetAttribute(k_inside_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        void* args[] = {&N, &L, &B, &TB, (void*)&W, (void*)&lengths, &cv, &ce};
        { ProfScope ps(KC_SIMT_INSIDE, st);
          RBN_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)k_inside_all, dim3(coop_grid), dim3(kThreads), args, smem, st)); }
    } else {
        for (int w = 2; w <= L; ++w) {
            const int Bw = active ? active[w] : B;          // sorted ragged batch: sequences that reach this width
            if (Bw == 0) break;
            { ProfScope ps(KC_SIMT_INSIDE, st); k_inside_diag<<<Bw * (L - w + 1), kThreads, smem, st>>>(N, L, w, TB, W, lengths, cv, ce); }
            RBN_LAUNCH_CHECK("k_inside_diag");
        }
    }
    { ProfScope ps(KC_OTHER, st); k_root<<<B, 128, 0, st>>>(N, L, prior, lengths, cv, ce, logZ, zroot, rootexp, (float*)(base + ws.zroot),
                              (int*)(base + ws.rootexp)); }
    RBN_LAUNCH_CHECK("k_root");
    return RBN_OK;
}

int simt_backward(const RbnDesc* d, const WsLayout& ws, char* base, const float* W, const float* T, const float* prior,
                  const int* tokens, const int* lengths, const float* coef, float* gW, float* gT, float* gPrior,
                  const int* active, cudaStream_t st) {
    (void)T;
    float* cv = (float*)(base + ws.chart_val);
    int* ce = (int*)(base + ws.chart_exp);
    float* alpha = (float*)(base + ws.alpha);
    int N = d->N, L = d->L, B = d->B;
    size_t C = cells_per_seq(L);
    RBN_CUDA_CHECK(cudaMemsetAsync(alpha, 0, (size_t)B * C * N * sizeof(float), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(gW, 0, (size_t)N * N * N * sizeof(float), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(gT, 0, (size_t)d->V * N * sizeof(float), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(gPrior, 0, (size_t)N * sizeof(float), st));
    { ProfScope ps(KC_OTHER, st); k_outside_seed<<<B, 128, 0, st>>>(N, L, prior, lengths, cv, (const float*)(base + ws.zroot), coef, alpha, gPrior); }
    RBN_LAUNCH_CHECK("k_outside_seed");
    int TB = pick_tb(N, 2);
    size_t smem = ((size_t)(L + 1) + N + 2 * (size_t)TB * N) * sizeof(float);
    RBN_CUDA_CHECK(cudaFuncSetAttribute(k_outside_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int coop_grid = coop_grid_for((const void*)k_outside_all, smem, B * (L - 1));
    if (coop_grid > 0 && L >= 2) {
        RBN_CUDA_CHECK(cudaFuncSetAttribute(k_outside_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const float* cvc = cv;
        const int* cec = ce;
        void* args[] = {&N, &L, &B, &TB, (void*)&W, (void*)&lengths, (void*)&cvc, (void*)&cec, &alpha, &gW};
        { ProfScope ps(KC_SIMT_OUTSIDE, st);
          RBN_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)k_outside_all, dim3(coop_grid), dim3(kThreads), args, smem, st)); }
    } else {
        for (int w = L; w >= 2; --w) {
            const int Bw = active ? active[w] : B;
            if (Bw == 0) continue;
            { ProfScope ps(KC_SIMT_OUTSIDE, st); k_outside_diag<<<Bw * (L - w + 1), kThreads, smem, st>>>(N, L, w, TB, W, lengths, cv, ce, alpha, gW); }
            RBN_LAUNCH_CHECK("k_outside_diag");
        }
    }
    { ProfScope ps(KC_EMISSION, st); k_emission_bwd<<<B * L, 128, 0, st>>>(N, L, d->n_tvars, tokens, lengths, ce, alpha, nullptr, gT); }
    RBN_LAUNCH_CHECK("k_emission_bwd");
    return RBN_OK;
}

int simt_viterbi(const RbnDesc* d, const WsLayout& ws, char* base, const float* W, const float* T, const float* prior,
                 const int* tokens, const int* lengths, float* logP, int* back, int* root_sym, cudaStream_t st) {
    if (d->N > 1024 || d->L > 2047) {
        set_error("viterbi needs N <= 1024 and L <= 2047 (back-pointer packing); got N=%d L=%d", d->N, d->L);
        return RBN_E_UNSUPPORTED;
    }
    float* cv = (float*)(base + ws.chart_val);
    int* ce = (int*)(base + ws.chart_exp);
    int N = d->N, L = d->L, B = d->B;
    RBN_CUDA_CHECK(cudaMemsetAsync(back, 0xff, (size_t)B * cells_per_seq(L) * N * sizeof(int), st));
    { ProfScope ps(KC_EMISSION, st); k_emission<<<B * L, kThreads, 0, st>>>(N, L, d->n_tvars, T, tokens, lengths, cv, ce); }
    RBN_LAUNCH_CHECK("k_emission");
    int TB = pick_tb(N, 2);
    size_t smem = ((size_t)(L + 1) + 2 * (size_t)TB * N) * sizeof(float);
    RBN_CUDA_CHECK(cudaFuncSetAttribute(k_viterbi_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int w = 2; w <= L; ++w) {
        { ProfScope ps(KC_SIMT_INSIDE, st); k_viterbi_diag<<<B * (L - w + 1), kThreads, smem, st>>>(N, L, w, TB, W, lengths, cv, ce, back); }
        RBN_LAUNCH_CHECK("k_viterbi_diag");
    }
    { ProfScope ps(KC_OTHER, st); k_viterbi_root<<<B, kThreads, 0, st>>>(N, L, prior, lengths, cv, ce, logP, root_sym); }
    RBN_LAUNCH_CHECK("k_viterbi_root");
    return RBN_OK;
}

int export_chart(const RbnDesc* d, const WsLayout& ws, const char* base, const int* lengths, int seq, double* out,
                 cudaStream_t st) {
    const float* cv = (const float*)(base + ws.chart_val);
    const int* ce = (const int*)(base + ws.chart_exp);
    k_export_chart<<<cells_per_seq(d->L), 128, 0, st>>>(d->N, d->L, lengths, seq, cv, ce, out);
    RBN_LAUNCH_CHECK("k_export_chart");
    return RBN_OK;
}

int lognormalize(void* lp, int is_f64, long long gs, long long groups, int* bad, cudaStream_t st) {
    if (bad) RBN_CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    int threads = 128;
    long long blocks = (groups + threads - 1) / threads;
    if (is_f64) k_lognormalize<double><<<(unsigned)blocks, threads, 0, st>>>((double*)lp, gs, groups, bad);
    else k_lognormalize<float><<<(unsigned)blocks, threads, 0, st>>>((float*)lp, gs, groups, bad);
    RBN_LAUNCH_CHECK("k_lognormalize");
    return RBN_OK;
}

}  // namespace rbn
