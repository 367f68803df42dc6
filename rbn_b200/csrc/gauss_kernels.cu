// Continuous (multivariate-normal) RBN: inside and outside passes over span diagonals, fp64 charts (default, the
// parity-pinned mode) or, for D = 8, float32 charts with fp64 log weights (RBN_GAUSS_F32: the reference's default dtype).
//
// Every chart cell is a scaled Gaussian  beta(x) = exp(logc) * N(x; mu, Sigma)  stored as [logc | mu[D] | Sigma[D][D]].
// Per split the two children are multiplied as in rbnet.multivariate_normal.PairwiseProduct
// (/root/reference/rbnet/multivariate_normal.py:339-396) after adding the transition noise, and the splits of a span
// are collapsed by moment matching as in ApproximateMixture (multivariate_normal.py:502-550).  The composition is
// the one of oracle/gauss_oracle.py (the reference has no Gaussian RBN bricks).
//
// D <= 4: one CTA per span, one thread per split; D x D Cholesky factorisations live in registers, fp64 (parity with the
// fp64 oracle is then limited by summation order only).  D = 8: eight lanes per split (namespace g8 below).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace rbn {

// Cell layout in units of the storage type T: [log c (ALWAYS one fp64, i.e. kHead slots of T) | mu[D] | Sigma[D][D]].
// T = double: 1 + D + D^2 doubles (584 B at D = 8).  T = float (D = 8 only): 2 + 8 + 64 floats = 296 B, 8-byte aligned, so
// the log weight keeps 53 bits while means and covariances are stored (and worked on) in fp32.
template <int D, typename T = double>
struct GCell {
    static constexpr int kHead = (int)(sizeof(double) / sizeof(T));
    static constexpr int kStride = kHead + D + D * D;
};
template <typename T> __device__ __forceinline__ double ld_logc(const T* cell) { return *reinterpret_cast<const double*>(cell); }
template <typename T> __device__ __forceinline__ void st_logc(T* cell, double v) { *reinterpret_cast<double*>(cell) = v; }
__device__ __forceinline__ float tsqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double tsqrt(double v) { return sqrt(v); }
__device__ __forceinline__ float tlog(float v) { return logf(v); }
__device__ __forceinline__ double tlog(double v) { return log(v); }
__device__ __forceinline__ float texp(float v) { return expf(v); }
__device__ __forceinline__ double texp(double v) { return exp(v); }

// Register budget of the D = 8 row kernels.  Measured (C4, round 2, fp64): inside pass 21.3 ms at 252 registers (1-2 CTAs per
// SM), 15.6 ms at 128 (4 CTAs per SM: the kernel is latency-bound, more warps help); outside pass 38.6 ms at 255 registers,
// 39.1 ms at 168 (spills eat the gain).  fp32 charts: inside 8.4 / 7.6 / 7.35 ms at 114 / 96 / 80 registers, outside 25.2 / 23.9 /
// 23.4 ms at 229 / 168 / 128.  Default (RBN_GAUSS_OCC=1): fp64 inside 128, outside 255; fp32 inside 80 (6 CTAs per SM, no
// spills), outside 128 (4 CTAs per SM, 84 B of spills).  RBN_GAUSS_OCC=0: no cap; 2: the other capped variant (G8Budget).
static int gauss_occ() {
    const char* e = getenv("RBN_GAUSS_OCC");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
}
// RBN_GAUSS_ROWS=0 selects the one-thread-per-split kernels for D = 8 too (A/B runs)
static bool gauss_rows8() {
    const char* e = getenv("RBN_GAUSS_ROWS");
    return !(e && e[0] == '0');
}

// lower Cholesky factor of the symmetric matrix S (full storage), in place in the lower triangle; returns log det S
template <int D>
__device__ __forceinline__ double chol_inplace(double (&S)[D][D]) {
    double logdet = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double d = S[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= S[j][k] * S[j][k];
        d = sqrt(d);
        S[j][j] = d;
        logdet += log(d);
        const double inv = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double s = S[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= S[i][k] * S[j][k];
            S[i][j] = s * inv;
        }
    }
    return 2.0 * logdet;
}

// x <- (G G^T)^-1 x
template <int D>
__device__ __forceinline__ void chol_solve(const double (&G)[D][D], double (&x)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= G[i][k] * x[k];
        x[i] = s / G[i][i];
    }
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll
        for (int k = i + 1; k < D; ++k) s -= G[k][i] * x[k];
        x[i] = s / G[i][i];
    }
}

// one split: children cells l, r -> log weight, and (optionally) mean / covariance of the product
template <int D, bool FULL>
__device__ __forceinline__ double split_product(const double* __restrict__ l, const double* __restrict__ r,
                                                const double* __restrict__ covL, const double* __restrict__ covR,
                                                double (&mean)[D], double (&A)[D][D], double (&cov)[D][D]) {
    double S[D][D];
    double d[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        d[a] = r[1 + a] - l[1 + a];
#pragma unroll
        for (int b = 0; b < D; ++b) {
            A[a][b] = l[1 + D + a * D + b] + covL[a * D + b];
            S[a][b] = A[a][b] + r[1 + D + a * D + b] + covR[a * D + b];
        }
    }
    const double logdet = chol_inplace<D>(S);
    double z[D];
#pragma unroll
    for (int a = 0; a < D; ++a) z[a] = d[a];
    chol_solve<D>(S, z);                       // z = S^-1 (mu_r - mu_l)
    double quad = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) quad += d[a] * z[a];
    const double lw = l[0] + r[0] - 0.5 * (quad + logdet + D * 1.8378770664093454836);   // log(2 pi)
    if (FULL) {
#pragma unroll
        for (int a = 0; a < D; ++a) {          // mean = mu_l + A z
            double s = l[1 + a];
#pragma unroll
            for (int b = 0; b < D; ++b) s += A[a][b] * z[b];
            mean[a] = s;
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {          // cov[:, c] = A[:, c] - A S^-1 A[:, c]
            double x[D];
#pragma unroll
            for (int a = 0; a < D; ++a) x[a] = A[a][c];
            chol_solve<D>(S, x);
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double s = A[a][c];
#pragma unroll
                for (int b = 0; b < D; ++b) s -= A[a][b] * x[b];
                cov[a][c] = s;
            }
        }
    }
    return lw;
}

template <int D, typename T>
__global__ void k_gauss_width1(int L, const T* __restrict__ y, const int* __restrict__ lengths,
                               const double* __restrict__ cov_term, const double* __restrict__ log_struct,
                               T* __restrict__ chart) {
    constexpr int H = GCell<D, T>::kHead, G = GCell<D, T>::kStride;
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    T* cell = chart + ((size_t)b * cells_per_seq(L) + i) * G;
    for (int t = threadIdx.x; t < 1 + D + D * D; t += blockDim.x) {
        if (t == 0) st_logc(cell, log_struct[0]);
        else if (t <= D) cell[H + t - 1] = y[((size_t)b * L + i) * D + (t - 1)];
        else cell[H + t - 1] = (T)cov_term[t - 1 - D];
    }
}

template <int D>
__global__ void __launch_bounds__(64)
k_gauss_diag(int L, int w, const int* __restrict__ lengths, const double* __restrict__ cov_left,
             const double* __restrict__ cov_right, const double* __restrict__ log_struct, double* __restrict__ chart) {
    constexpr int G = GCell<D>::kStride;
    extern __shared__ double sm[];
    double* lw = sm;                    // [L]
    double* acc = lw + L;               // [D + D*D]  weighted first and second moments
    double* cL = acc + D + D * D;       // [D*D]
    double* cR = cL + D * D;            // [D*D]
    __shared__ double red[2];
    const int nI = L - w + 1;
    const int b = blockIdx.x / nI, i = blockIdx.x % nI;
    if (i + w > lengths[b]) return;
    const double* seq = chart + (size_t)b * cells_per_seq(L) * G;
    for (int t = threadIdx.x; t < D * D; t += blockDim.x) {
        cL[t] = cov_left[t];
        cR[t] = cov_right[t];
    }
    for (int t = threadIdx.x; t < D + D * D; t += blockDim.x) acc[t] = 0.0;
    __syncthreads();
    double mean[D], A[D][D], cov[D][D];
    double mx = -INFINITY;
    for (int u = 1 + threadIdx.x; u < w; u += blockDim.x) {
        const double* l = seq + (size_t)(cell_off(u, L) + i) * G;
        const double* r = seq + (size_t)(cell_off(w - u, L) + i + u) * G;
        const double v = split_product<D, false>(l, r, cL, cR, mean, A, cov);
        lw[u] = v;
        mx = fmax(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmax(red[0], red[1]);
    __syncthreads();
    double se = 0.0;
    for (int u = 1 + threadIdx.x; u < w; u += blockDim.x) se += exp(lw[u] - mx);
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
    __syncthreads();
    const double log_norm = mx + log(red[0] + red[1]);
    // weighted first and second moments: every quantity is summed over the lanes of a warp with shuffles and added to
    // shared memory once per warp (64-way contended fp64 shared-memory atomics were the bottleneck of this kernel)
    const int lane = threadIdx.x & 31;
    for (int base = 1; base < w; base += blockDim.x) {
        if (base + (int)(threadIdx.x & ~31u) >= w) continue;          // no split left for this warp
        const int u = base + threadIdx.x;
        const bool active = u < w;
        double wt = 0.0;
        if (active) {
            const double* l = seq + (size_t)(cell_off(u, L) + i) * G;
            const double* r = seq + (size_t)(cell_off(w - u, L) + i + u) * G;
            split_product<D, true>(l, r, cL, cR, mean, A, cov);
            wt = exp(lw[u] - log_norm);
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double v = active ? wt * mean[a] : 0.0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(&acc[a], v);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double q = active ? wt * (cov[a][c] + mean[a] * mean[c]) : 0.0;
                for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                if (lane == 0) atomicAdd(&acc[D + a * D + c], q);
            }
        }
    }
    __syncthreads();
    double* out = chart + ((size_t)b * cells_per_seq(L) + cell_off(w, L) + i) * G;
    for (int t = threadIdx.x; t < G; t += blockDim.x) {
        double v;
        if (t == 0) v = log_norm + log_struct[1];
        else if (t <= D) v = acc[t - 1];
        else {
            const int a = (t - 1 - D) / D, c = (t - 1 - D) % D;
            v = acc[D + a * D + c] - acc[a] * acc[c];
        }
        out[t] = v;
    }
}


// =====================================================================================================================
// D = 8 (BASELINE config 4): eight lanes per split, lane r holds ROW r of every 8 x 8 matrix
// =====================================================================================================================
// The one-thread-per-split kernels above keep three to six 8 x 8 fp64 matrices per thread: 400-800 registers, i.e.
// local-memory traffic for every entry (3.4 % of the HBM roofline in round 1).  Here a split is worked on by a group of 8
// lanes; a matrix is 8 doubles per lane, rows move between lanes with warp shuffles (width 8), and a warp owns one span
// (4 splits at a time).  Everything that is symmetric (A, S^-1, the moments) needs no transposes in this layout; the
// products that involve a non-symmetric factor take the wanted COLUMN element out of a broadcast row (sel8).
//   S = L L^T row-distributed Cholesky, Y = L^-1 A by forward substitution (no back substitution anywhere):
//   quad = |L^-1 d|^2,  cov = A - Y^T Y,  mean = mu_l + Y^T (L^-1 d)      (PairwiseProduct, multivariate_normal.py:332-396)
// The moment matching over the splits of a span (ApproximateMixture, :502-550) is accumulated ONLINE (running maximum of
// the log weights with rescaling), so every split is visited once.
namespace g8 {

static constexpr int D8 = 8;
static constexpr double kLog2Pi = 1.8378770664093454836;

template <typename T> __device__ __forceinline__ T gshfl(unsigned mask, T v, int src) { return __shfl_sync(mask, v, src, 8); }
template <typename T> __device__ __forceinline__ T gsum(unsigned mask, T v) {
    v += __shfl_xor_sync(mask, v, 1, 8);
    v += __shfl_xor_sync(mask, v, 2, 8);
    v += __shfl_xor_sync(mask, v, 4, 8);
    return v;
}
// element r (lane-dependent) of a register array
template <typename T> __device__ __forceinline__ T sel8(const T (&v)[8], int r) {
    T x = v[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) x = (r == k) ? v[k] : x;
    return x;
}

// In place: row r of the symmetric S -> row r of its lower Cholesky factor (entries k <= r).  Returns L[r][r].
template <typename T> __device__ __forceinline__ T chol_rows(unsigned mask, int r, T (&s)[8]) {
    T mydiag = T(1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        T t = s[j];
#pragma unroll
        for (int k = 0; k < j; ++k) t -= s[k] * s[k];
        const T ljj = gshfl(mask, tsqrt(t), j);               // lane j's value is the meaningful one
        T acc = s[j];
#pragma unroll
        for (int k = 0; k < j; ++k) acc -= s[k] * gshfl(mask, s[k], j);
        if (r > j) s[j] = acc / ljj;
        else if (r == j) { s[j] = ljj; mydiag = ljj; }
    }
    return mydiag;
}

// Forward substitution L Y = X for 8 right-hand-side columns and one extra vector: lane r holds row r of X in x[] and
// entry r of the vector in xv; on return row r of L^-1 X and entry r of L^-1 v.
template <typename T>
__device__ __forceinline__ void fwd_subst(unsigned mask, int r, const T (&s)[8], T mydiag, T (&x)[8], T& xv) {
    const T inv = T(1) / mydiag;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (r == k) {
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] *= inv;
            xv *= inv;
        }
        const T lrk = s[k];                                     // L[r][k] (meaningful for r > k)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const T yk = gshfl(mask, x[c], k);
            if (r > k) x[c] -= lrk * yk;
        }
        const T yv = gshfl(mask, xv, k);
        if (r > k) xv -= lrk * yv;
    }
}

// T = storage and working type of means / covariances (double, or float = the reference's default dtype); log weights
// are fp64 in both (they reach -1e3 at L = 64: a float there is 6e-5 coarse, which the softmax over splits turns into a
// 6e-5 relative error per level).
template <typename T, int MINB>      // MINB: CTAs per SM the register budget is cut for
__global__ void __launch_bounds__(128, MINB)
k_gauss_diag8(int L, int w, int B, const int* __restrict__ lengths, const double* __restrict__ cov_left,
              const double* __restrict__ cov_right, const double* __restrict__ log_struct, T* __restrict__ chart) {
    constexpr int H = GCell<D8, T>::kHead, G = GCell<D8, T>::kStride;
    const int nI = L - w + 1;
    const int span = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (span >= B * nI) return;
    const int b = span / nI, i = span - b * nI;
    if (i + w > lengths[b]) return;
    const int lane = threadIdx.x & 31, g = lane >> 3, r = lane & 7;
    const unsigned mask = 0xFFu << (lane & 24);
    const T* seq = chart + (size_t)b * cells_per_seq(L) * G;
    T cl[8], cr[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { cl[c] = (T)cov_left[r * 8 + c]; cr[c] = (T)cov_right[r * 8 + c]; }
    // running moment-matching state of this group: max log weight, sum of weights, first moment (entry r), second moment (row r)
    double mx = -INFINITY;
    T se = T(0), a1 = T(0), a2[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a2[c] = T(0);
    for (int u = 1 + g; u < w; u += 4) {
        const T* lc = seq + (size_t)(cell_off(u, L) + i) * G;
        const T* rc = seq + (size_t)(cell_off(w - u, L) + i + u) * G;
        T A[8], s[8], x[8];
        const T mul = lc[H + r];
        T dv = rc[H + r] - mul;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            A[c] = lc[H + D8 + r * 8 + c] + cl[c];
            s[c] = A[c] + rc[H + D8 + r * 8 + c] + cr[c];
            x[c] = A[c];
        }
        const T diag = chol_rows(mask, r, s);
        fwd_subst(mask, r, s, diag, x, dv);                     // x = row r of Y = L^-1 A, dv = (L^-1 d)[r]
        const T quad = gsum(mask, dv * dv);
        const T logdet = T(2) * gsum(mask, tlog(diag));
        const double lw = ld_logc(lc) + ld_logc(rc) - 0.5 * ((double)quad + (double)logdet + D8 * kLog2Pi);
        // cov = A - Y^T Y (row r), mean = mu_l + Y^T (L^-1 d) (entry r)
        T C[8], m = mul;
#pragma unroll
        for (int c = 0; c < 8; ++c) C[c] = A[c];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            T yk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) yk[c] = gshfl(mask, x[c], k);
            const T ya = sel8(yk, r);                            // Y[k][r]
            const T dk = gshfl(mask, dv, k);
#pragma unroll
            for (int c = 0; c < 8; ++c) C[c] -= ya * yk[c];
            m += ya * dk;
        }
        // online moment matching
        const double nmx = fmax(mx, lw);
        const T sc = (mx == -INFINITY) ? T(0) : texp((T)(mx - nmx));
        const T wt = texp((T)(lw - nmx));
        se = se * sc + wt;
        a1 = a1 * sc + wt * m;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const T mc = gshfl(mask, m, c);
            a2[c] = a2[c] * sc + wt * (C[c] + m * mc);
        }
        mx = nmx;
    }
    // combine the four groups of the warp (same rescaling), every lane ends with the span's totals
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        const double omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const T ose = __shfl_xor_sync(0xffffffffu, se, o);
        const T oa1 = __shfl_xor_sync(0xffffffffu, a1, o);
        const double nmx = fmax(mx, omx);
        const T s0 = (mx == -INFINITY) ? T(0) : texp((T)(mx - nmx)), s1 = (omx == -INFINITY) ? T(0) : texp((T)(omx - nmx));
        se = se * s0 + ose * s1;
        a1 = a1 * s0 + oa1 * s1;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const T oa = __shfl_xor_sync(0xffffffffu, a2[c], o);
            a2[c] = a2[c] * s0 + oa * s1;
        }
        mx = nmx;
    }
    const double log_norm = mx + log((double)se);
    const T inv = T(1) / se;
    const T mean = a1 * inv;
    T* out = chart + ((size_t)b * cells_per_seq(L) + cell_off(w, L) + i) * G;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const T mc = __shfl_sync(0xffffffffu, mean, c, 8);
        if (g == 0) out[H + D8 + r * 8 + c] = a2[c] * inv - mean * mc;
    }
    if (g == 0) {
        out[H + r] = mean;
        if (r == 0) st_logc(out, log_norm + log_struct[1]);
    }
}

}  // namespace g8

template <int D, typename T>
__global__ void k_gauss_root(int L, int B, const int* __restrict__ lengths, const double* __restrict__ prior_mean,
                             const double* __restrict__ prior_cov, const T* __restrict__ chart,
                             double* __restrict__ logZ) {
    constexpr int H = GCell<D, T>::kHead, G = GCell<D, T>::kStride;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int n = lengths[b];
    const T* cell = chart + ((size_t)b * cells_per_seq(L) + cell_off(n, L)) * G;
    double S[D][D], d[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        d[a] = (double)cell[H + a] - prior_mean[a];
#pragma unroll
        for (int c = 0; c < D; ++c) S[a][c] = (double)cell[H + D + a * D + c] + prior_cov[a * D + c];
    }
    const double logdet = chol_inplace<D>(S);
    double z[D];
#pragma unroll
    for (int a = 0; a < D; ++a) z[a] = d[a];
    chol_solve<D>(S, z);
    double quad = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) quad += d[a] * z[a];
    logZ[b] = ld_logc(cell) - 0.5 * (quad + logdet + D * 1.8378770664093454836);
}

// Register budgets of the D = 8 row kernels per storage type (see gauss_occ()).
template <typename T> struct G8Budget;
// kFwd / kBwd: CTAs per SM of the default budget; kFwd2 / kBwd2: of the narrow one (RBN_GAUSS_OCC=2).
template <> struct G8Budget<double> { static constexpr int kFwd = 4, kFwd2 = 4, kBwd = 1, kBwd2 = 3; };
template <> struct G8Budget<float> { static constexpr int kFwd = 6, kFwd2 = 5, kBwd = 4, kBwd2 = 3; };

template <int D, typename T>
static int gauss_forward_t(int B, int L, const T* y, const int* lengths, const double* cov_term,
                           const double* cov_left, const double* cov_right, const double* prior_mean,
                           const double* prior_cov, const double* log_struct, T* chart, double* logZ,
                           cudaStream_t st) {
    constexpr bool kF64 = std::is_same<T, double>::value;
    { ProfScope ps(KC_EMISSION, st); k_gauss_width1<D, T><<<B * L, 96, 0, st>>>(L, y, lengths, cov_term, log_struct, chart); }
    RBN_LAUNCH_CHECK("k_gauss_width1");
    static const bool rows8 = gauss_rows8();
    for (int w = 2; w <= L; ++w) {
        { ProfScope ps(KC_SIMT_INSIDE, st);
          if (D == 8 && (rows8 || !kF64)) {
              static const int occ = gauss_occ();
              const int grid = (B * (L - w + 1) + 3) / 4;
              if (occ >= 2) g8::k_gauss_diag8<T, G8Budget<T>::kFwd2><<<grid, 128, 0, st>>>(L, w, B, lengths, cov_left, cov_right, log_struct, chart);
              else if (occ == 1) g8::k_gauss_diag8<T, G8Budget<T>::kFwd><<<grid, 128, 0, st>>>(L, w, B, lengths, cov_left, cov_right, log_struct, chart);
              else g8::k_gauss_diag8<T, 1><<<grid, 128, 0, st>>>(L, w, B, lengths, cov_left, cov_right, log_struct, chart);
          } else if constexpr (kF64) {
              const size_t smem = ((size_t)L + D + 3 * D * D) * sizeof(double);
              k_gauss_diag<D><<<B * (L - w + 1), 64, smem, st>>>(L, w, lengths, cov_left, cov_right, log_struct, chart);
          } }
        RBN_LAUNCH_CHECK("k_gauss_diag");
    }
    { ProfScope ps(KC_OTHER, st); k_gauss_root<D, T><<<(B + 63) / 64, 64, 0, st>>>(L, B, lengths, prior_mean, prior_cov, chart, logZ); }
    RBN_LAUNCH_CHECK("k_gauss_root");
    return RBN_OK;
}

// bytes of ONE chart (the adjoint chart has the same layout); f32 != 0: fp32 storage (D = 8 only)
size_t gauss_workspace_bytes(int D, int B, int L, int f32) {
    const size_t cell = f32 ? (size_t)(2 + D + D * D) * sizeof(float) : (size_t)(1 + D + D * D) * sizeof(double);
    return align_up((size_t)B * cells_per_seq(L) * cell, 1024);
}

int gauss_forward(int D, int B, int L, const double* y, const int* lengths, const double* cov_term,
                  const double* cov_left, const double* cov_right, const double* prior_mean, const double* prior_cov,
                  const double* log_struct, double* chart, double* logZ, cudaStream_t st) {
#define RBN_G(DD) case DD: return gauss_forward_t<DD, double>(B, L, y, lengths, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct, chart, logZ, st)
    switch (D) {
        RBN_G(1); RBN_G(2); RBN_G(3); RBN_G(4); RBN_G(8);
        default:
            set_error("Gaussian path is instantiated for D in {1,2,3,4,8} (got %d)", D);
            return RBN_E_UNSUPPORTED;
    }
#undef RBN_G
}

int gauss_forward_f32(int D, int B, int L, const float* y, const int* lengths, const double* cov_term,
                      const double* cov_left, const double* cov_right, const double* prior_mean, const double* prior_cov,
                      const double* log_struct, float* chart, double* logZ, cudaStream_t st) {
    if (D != 8) {
        set_error("fp32 chart storage is instantiated for D = 8 (got %d); other D run through the fp64 entry points", D);
        return RBN_E_UNSUPPORTED;
    }
    return gauss_forward_t<8, float>(B, L, y, lengths, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct, chart, logZ, st);
}

}  // namespace rbn

// =====================================================================================================================
// outside pass of the Gaussian RBN: reverse-mode through moment matching and the pairwise products
// =====================================================================================================================
// Adjoint chart cell = [a_bar | mu_bar[D] | Sigma_bar[D][D]] of sum_b coef_b log Z_b.  One warp per span, one lane per
// split; the D x D work matrices of a lane are thread-private.
namespace rbn {

#define SMAT(M, r, c) M[(r) * D + (c)]

template <int D, typename T>
__global__ void k_gauss_bwd_root(int L, int B, const int* __restrict__ lengths, const double* __restrict__ prior_mean,
                                 const double* __restrict__ prior_cov, const T* __restrict__ chart,
                                 const double* __restrict__ coef, T* __restrict__ adj,
                                 double* __restrict__ g_prior_mean, double* __restrict__ g_prior_cov) {
    constexpr int H = GCell<D, T>::kHead, G = GCell<D, T>::kStride;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double c = coef[b];
    if (c == 0.0) return;
    const int n = lengths[b];
    const size_t cidx = (size_t)b * cells_per_seq(L) + cell_off(n, L);
    const T* cell = chart + cidx * G;
    T* out = adj + cidx * G;
    double S[D][D], d[D], z[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        d[a] = (double)cell[H + a] - prior_mean[a];
        z[a] = d[a];
#pragma unroll
        for (int e = 0; e < D; ++e) S[a][e] = (double)cell[H + D + a * D + e] + prior_cov[a * D + e];
    }
    chol_inplace<D>(S);
    chol_solve<D>(S, z);
    st_logc(out, ld_logc(out) + c);
#pragma unroll
    for (int a = 0; a < D; ++a) {
        out[H + a] += (T)(-c * z[a]);
        atomicAdd(&g_prior_mean[a], c * z[a]);
    }
#pragma unroll
    for (int e = 0; e < D; ++e) {          // column e of P = S^-1
        double x[D];
#pragma unroll
        for (int a = 0; a < D; ++a) x[a] = (a == e) ? 1.0 : 0.0;
        chol_solve<D>(S, x);
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const double sb = 0.5 * c * (z[a] * z[e] - x[a]);
            out[H + D + a * D + e] += (T)sb;
            atomicAdd(&g_prior_cov[a * D + e], sb);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(32)
k_gauss_bwd_diag(int L, int w, const int* __restrict__ lengths, const double* __restrict__ cov_left,
                 const double* __restrict__ cov_right, const double* __restrict__ chart, double* __restrict__ adj,
                 double* __restrict__ g_cov_left, double* __restrict__ g_cov_right, double* __restrict__ g_log_struct) {
    constexpr int G = GCell<D>::kStride;
    constexpr int DD = D * D;
    extern __shared__ double sm[];
    // D x D work matrices of this lane's split: thread-private (registers / local memory, which the hardware interleaves
    // per lane and caches in L1).  Keeping them in shared memory cost 98 KB per warp at D = 8, i.e. two warps per SM.
    double mA[DD], mP[DD], mK[DD], mQ[DD], mAb[DD], mSb[DD];
    double* lw = sm;                 // [L]
    double* ob = lw + L;             // [L]  omega_bar
    double* par = ob + L;            // parent: [0] a_bar, [1..D] mu_bar_tot, [1+D..] M2_bar (symmetric)
    double* pmu = par + G;           // parent mean [D]
    double* cL = pmu + D;            // [DD]
    double* cR = cL + DD;            // [DD]
    double* gL = cR + DD;            // [DD] accumulators
    double* gR = gL + DD;            // [DD]
    const int lane = threadIdx.x;
    const int nI = L - w + 1;
    const int b = blockIdx.x / nI, i = blockIdx.x % nI;
    if (i + w > lengths[b]) return;
    const size_t seq_off = (size_t)b * cells_per_seq(L);
    const double* seq = chart + seq_off * G;
    double* aseq = adj + seq_off * G;
    const size_t pcell = cell_off(w, L) + i;
    const double abar = aseq[pcell * G];
    // skip spans that no gradient reaches
    {
        double any = 0.0;
        for (int t = lane; t < G; t += 32) any += fabs(aseq[pcell * G + t]);
        for (int o = 16; o > 0; o >>= 1) any += __shfl_xor_sync(0xffffffffu, any, o);
        if (any == 0.0) return;
    }
    for (int t = lane; t < DD; t += 32) {
        cL[t] = cov_left[t];
        cR[t] = cov_right[t];
        gL[t] = 0.0;
        gR[t] = 0.0;
        const int a = t / D, c = t % D;
        par[1 + D + t] = 0.5 * (aseq[pcell * G + 1 + D + a * D + c] + aseq[pcell * G + 1 + D + c * D + a]);
    }
    if (lane < D) pmu[lane] = seq[pcell * G + 1 + lane];
    __syncwarp();
    if (lane < D) {
        double s = aseq[pcell * G + 1 + lane];
        for (int c = 0; c < D; ++c) s -= 2.0 * par[1 + D + lane * D + c] * pmu[c];
        par[1 + lane] = s;             // mu_bar_tot = mu_bar - 2 Sigma_bar mu
    }
    if (lane == 0) {
        par[0] = abar;
        atomicAdd(&g_log_struct[1], abar);
    }
    __syncwarp();

    double mean[D], Areg[D][D], cov[D][D];
    double tL[DD], tR[DD];           // this lane's share of the covariance gradients
#pragma unroll
    for (int t = 0; t < DD; ++t) tL[t] = tR[t] = 0.0;
    // pass 1: log weights and omega_bar
    double mx = -INFINITY;
    for (int u = 1 + lane; u < w; u += 32) {
        const double* l = seq + (size_t)(cell_off(u, L) + i) * G;
        const double* r = seq + (size_t)(cell_off(w - u, L) + i + u) * G;
        const double v = split_product<D, true>(l, r, cL, cR, mean, Areg, cov);
        lw[u] = v;
        mx = fmax(mx, v);
        double o = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
            o += par[1 + a] * mean[a];
#pragma unroll
            for (int c = 0; c < D; ++c) o += par[1 + D + a * D + c] * (cov[a][c] + mean[a] * mean[c]);
        }
        ob[u] = o;
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double se = 0.0;
    for (int u = 1 + lane; u < w; u += 32) se += exp(lw[u] - mx);
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const double log_norm = mx + log(se);
    double swo = 0.0;
    for (int u = 1 + lane; u < w; u += 32) swo += exp(lw[u] - log_norm) * ob[u];
    for (int o = 16; o > 0; o >>= 1) swo += __shfl_xor_sync(0xffffffffu, swo, o);

    // pass 2: adjoints of every split
    for (int u = 1 + lane; u < w; u += 32) {
        const size_t lcell = cell_off(u, L) + i, rcell = cell_off(w - u, L) + i + u;
        const double* l = seq + lcell * G;
        const double* r = seq + rcell * G;
        const double om = exp(lw[u] - log_norm);
        const double lwb = om * (ob[u] - swo + abar);
        // recompute A, S -> P = S^-1, z, m
        double S[D][D], d[D], z[D], m[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            d[a] = r[1 + a] - l[1 + a];
            z[a] = d[a];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const double av = l[1 + D + a * D + c] + cL[a * D + c];
                SMAT(mA, a, c) = av;
                S[a][c] = av + r[1 + D + a * D + c] + cR[a * D + c];
            }
        }
        chol_inplace<D>(S);
        chol_solve<D>(S, z);
#pragma unroll
        for (int e = 0; e < D; ++e) {
            double x[D];
#pragma unroll
            for (int a = 0; a < D; ++a) x[a] = (a == e) ? 1.0 : 0.0;
            chol_solve<D>(S, x);
#pragma unroll
            for (int a = 0; a < D; ++a) SMAT(mP, a, e) = x[a];
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double s = l[1 + a];
#pragma unroll
            for (int c = 0; c < D; ++c) s += SMAT(mA, a, c) * z[c];
            m[a] = s;
        }
        // K = A P
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += SMAT(mA, a, k) * SMAT(mP, k, c);
                SMAT(mK, a, c) = s;
            }
        // Q = M2_bar K   (C_bar = om * M2_bar)
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += par[1 + D + a * D + k] * SMAT(mK, k, c);
                SMAT(mQ, a, c) = s;
            }
        // m_bar = om (mu_bar_tot + 2 M2_bar m);  z_bar = A m_bar;  pz = P z_bar;  d_bar = -lwb z + pz
        double mb[D], zb[D], pz[D], db[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double s = par[1 + a];
#pragma unroll
            for (int c = 0; c < D; ++c) s += 2.0 * par[1 + D + a * D + c] * m[c];
            mb[a] = om * s;
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) s += SMAT(mA, a, c) * mb[c];
            zb[a] = s;
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) s += SMAT(mP, a, c) * zb[c];
            pz[a] = s;
            db[a] = -lwb * z[a] + s;
        }
        // S_bar = 1/2 lwb (z z^T - P) - pz z^T + om K^T Q ;  A_bar = mb z^T + om (M2_bar - Q - Q^T) + S_bar
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double ktq = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) ktq += SMAT(mK, k, a) * SMAT(mQ, k, c);
                const double sb = 0.5 * lwb * (z[a] * z[c] - SMAT(mP, a, c)) - pz[a] * z[c] + om * ktq;
                SMAT(mSb, a, c) = sb;
                SMAT(mAb, a, c) = mb[a] * z[c] + om * (par[1 + D + a * D + c] - SMAT(mQ, a, c) - SMAT(mQ, c, a)) + sb;
            }
        // scatter: children adjoints and parameter gradients (symmetrised)
        double* la = aseq + lcell * G;
        double* ra = aseq + rcell * G;
        atomicAdd(&la[0], lwb);
        atomicAdd(&ra[0], lwb);
#pragma unroll
        for (int a = 0; a < D; ++a) {
            atomicAdd(&la[1 + a], mb[a] - db[a]);
            atomicAdd(&ra[1 + a], db[a]);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const double sa = 0.5 * (SMAT(mAb, a, c) + SMAT(mAb, c, a));
                const double ss = 0.5 * (SMAT(mSb, a, c) + SMAT(mSb, c, a));
                atomicAdd(&la[1 + D + a * D + c], sa);
                atomicAdd(&ra[1 + D + a * D + c], ss);
                tL[a * D + c] += sa;
                tR[a * D + c] += ss;
            }
        }
    }
    // sum the lanes' covariance gradients with shuffles (fp64 shared-memory atomics serialise 32-way), then one global
    // atomic per entry and span
#pragma unroll
    for (int t = 0; t < DD; ++t) {
        double vl = tL[t], vr = tR[t];
        for (int o = 16; o > 0; o >>= 1) {
            vl += __shfl_xor_sync(0xffffffffu, vl, o);
            vr += __shfl_xor_sync(0xffffffffu, vr, o);
        }
        if (lane == 0) {
            if (vl != 0.0) atomicAdd(&g_cov_left[t], vl);
            if (vr != 0.0) atomicAdd(&g_cov_right[t], vr);
        }
    }
}


// Outside pass for D = 8 in the same layout (8 lanes per split, rows in registers).  The adjoint of a split needs the
// softmax normaliser of its span and sum_u omega_u ob_u; both are known from the PARENT cell the inside pass wrote
// (log c, and <g1, mu_p> + <M2_bar, Sigma_p + mu_p mu_p^T>), so every split is visited once.  With W = L^-1 (forward
// substitution of the identity):  P = S^-1 = W^T W,  z = W^T (L^-1 d),  K = A P,  Q = M2_bar K,  <M2_bar, C> = <M2_bar - Q, A>.
// Child adjoints and covariance gradients are accumulated UNsymmetrised (every reader of the adjoint chart symmetrises on
// load; gauss_backward_t symmetrises the parameter gradients once at the end).
namespace g8 {

// Storage type T as in k_gauss_diag8: log weights, the adjoint of log c and every PARAMETER gradient stay fp64.
template <typename T, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_gauss_bwd_diag8(int L, int w, int B, const int* __restrict__ lengths, const double* __restrict__ cov_left,
                  const double* __restrict__ cov_right, const double* __restrict__ log_struct,
                  const T* __restrict__ chart, T* __restrict__ adj, double* __restrict__ g_cov_left,
                  double* __restrict__ g_cov_right, double* __restrict__ g_log_struct) {
    constexpr int H = GCell<D8, T>::kHead, G = GCell<D8, T>::kStride;
    const int nI = L - w + 1;
    const int span = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (span >= B * nI) return;
    const int b = span / nI, i = span - b * nI;
    if (i + w > lengths[b]) return;
    const int lane = threadIdx.x & 31, g = lane >> 3, r = lane & 7;
    const unsigned mask = 0xFFu << (lane & 24);
    const size_t seq_off = (size_t)b * cells_per_seq(L);
    const T* seq = chart + seq_off * G;
    T* aseq = adj + seq_off * G;
    const size_t pcell = cell_off(w, L) + i;
    const T* pa = aseq + pcell * G;
    const double abar_d = ld_logc(pa);
    {   // skip spans that no gradient reaches
        T any = (lane == 0) ? (T)fabs(abar_d) : T(0);
        for (int t = H + lane; t < G; t += 32) any += fabs(pa[t]);
        for (int o = 16; o > 0; o >>= 1) any += __shfl_xor_sync(0xffffffffu, any, o);
        if (any == T(0) && abar_d == 0.0) return;
    }
    const T abar = (T)abar_d;
    T M2b[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) M2b[c] = T(0.5) * (pa[H + D8 + r * 8 + c] + pa[H + D8 + c * 8 + r]);
    const T pm = seq[pcell * G + H + r];
    T g1 = pa[H + r], swo;
    {
        T t = T(0), q = T(0);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const T pmc = gshfl(mask, pm, c);
            t += M2b[c] * pmc;                                                   // (M2_bar mu_p)[r]
            q += M2b[c] * seq[pcell * G + H + D8 + r * 8 + c];                  // <M2_bar, Sigma_p>, row r
        }
        g1 -= T(2) * t;                                                          // mu_bar_tot = mu_bar - 2 Sigma_bar mu_p
        swo = gsum(mask, g1 * pm + q + pm * t);
    }
    const double log_norm = ld_logc(seq + pcell * G) - log_struct[1];
    if (lane == 0) atomicAdd(&g_log_struct[1], abar_d);
    T tL[8], tR[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) tL[c] = tR[c] = T(0);

    for (int u = 1 + g; u < w; u += 4) {
        const size_t lcell = cell_off(u, L) + i, rcell = cell_off(w - u, L) + i + u;
        const T* lc = seq + lcell * G;
        const T* rc = seq + rcell * G;
        T A[8], s[8], x[8];
        const T mul = lc[H + r];
        T yd = rc[H + r] - mul;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            A[c] = lc[H + D8 + r * 8 + c] + (T)cov_left[r * 8 + c];
            s[c] = A[c] + rc[H + D8 + r * 8 + c] + (T)cov_right[r * 8 + c];
            x[c] = (c == r) ? T(1) : T(0);
        }
        const T diag = chol_rows(mask, r, s);
        fwd_subst(mask, r, s, diag, x, yd);                      // x = row r of W = L^-1, yd = (L^-1 d)[r]
        const T quad = gsum(mask, yd * yd);
        const T logdet = T(2) * gsum(mask, tlog(diag));
        const double lw = ld_logc(lc) + ld_logc(rc) - 0.5 * ((double)quad + (double)logdet + D8 * kLog2Pi);
        const T om = texp((T)(lw - log_norm));
        // P = W^T W (row r), z = W^T (L^-1 d) (entry r)
        T P[8], z = T(0);
#pragma unroll
        for (int c = 0; c < 8; ++c) P[c] = T(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            T wk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) wk[c] = gshfl(mask, x[c], k);
            const T wa = sel8(wk, r);
            const T ydk = gshfl(mask, yd, k);
#pragma unroll
            for (int c = 0; c < 8; ++c) P[c] += wa * wk[c];
            z += wa * ydk;
        }
        // K = A P (row r)
        T K[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) K[c] = T(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int c = 0; c < 8; ++c) K[c] += A[k] * gshfl(mask, P[c], k);
        }
        T zc[8], m = mul;
#pragma unroll
        for (int c = 0; c < 8; ++c) { zc[c] = gshfl(mask, z, c); m += A[c] * zc[c]; }    // m = mu_l + A z
        // Q = M2_bar K (row r)
        T Q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) Q[c] = T(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int c = 0; c < 8; ++c) Q[c] += M2b[k] * gshfl(mask, K[c], k);
        }
        T t = T(0), part = T(0);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            t += M2b[c] * gshfl(mask, m, c);                                     // (M2_bar m)[r]
            part += (M2b[c] - Q[c]) * A[c];                                      // <M2_bar, C> = <M2_bar - Q, A>
        }
        const T ob = gsum(mask, g1 * m + part + m * t);
        const T lwb = om * (ob - swo + abar);
        const T mb = om * (g1 + T(2) * t);
        // K^T Q (row r) and row r of Q^T
        T KtQ[8], QT[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) KtQ[c] = T(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            T kk[8], qk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { kk[c] = gshfl(mask, K[c], k); qk[c] = gshfl(mask, Q[c], k); }
            const T ka = sel8(kk, r);
            QT[k] = sel8(qk, r);
#pragma unroll
            for (int c = 0; c < 8; ++c) KtQ[c] += ka * qk[c];
        }
        T zb = T(0), pz = T(0);
#pragma unroll
        for (int c = 0; c < 8; ++c) zb += A[c] * gshfl(mask, mb, c);              // z_bar = A m_bar
#pragma unroll
        for (int c = 0; c < 8; ++c) pz += P[c] * gshfl(mask, zb, c);              // P z_bar
        const T db = -lwb * z + pz;
        T* la = aseq + lcell * G;
        T* ra = aseq + rcell * G;
        if (r == 0) {
            atomicAdd(reinterpret_cast<double*>(la), (double)lwb);
            atomicAdd(reinterpret_cast<double*>(ra), (double)lwb);
        }
        atomicAdd(&la[H + r], mb - db);
        atomicAdd(&ra[H + r], db);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const T sb = T(0.5) * lwb * (z * zc[c] - P[c]) - pz * zc[c] + om * KtQ[c];
            const T ab = mb * zc[c] + om * (M2b[c] - Q[c] - QT[c]) + sb;
            atomicAdd(&la[H + D8 + r * 8 + c], ab);
            atomicAdd(&ra[H + D8 + r * 8 + c], sb);
            tL[c] += ab;
            tR[c] += sb;
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        T vl = tL[c], vr = tR[c];
        vl += __shfl_xor_sync(0xffffffffu, vl, 8);  vr += __shfl_xor_sync(0xffffffffu, vr, 8);
        vl += __shfl_xor_sync(0xffffffffu, vl, 16); vr += __shfl_xor_sync(0xffffffffu, vr, 16);
        if (g == 0) {
            if (vl != T(0)) atomicAdd(&g_cov_left[r * 8 + c], (double)vl);
            if (vr != T(0)) atomicAdd(&g_cov_right[r * 8 + c], (double)vr);
        }
    }
}

// g <- (g + g^T) / 2 for the two 8 x 8 covariance gradients
__global__ void k_sym8(double* __restrict__ a, double* __restrict__ b) {
    const int t = threadIdx.x, r = t >> 3, c = t & 7;
    double* m = blockIdx.x ? b : a;
    const double v = 0.5 * (m[r * 8 + c] + m[c * 8 + r]);
    __syncthreads();
    m[r * 8 + c] = v;
}

}  // namespace g8

template <int D, typename T>
__global__ void k_gauss_bwd_width1(int L, const int* __restrict__ lengths, const T* __restrict__ adj,
                                   T* __restrict__ g_y, double* __restrict__ g_cov_term,
                                   double* __restrict__ g_log_struct) {
    constexpr int H = GCell<D, T>::kHead, G = GCell<D, T>::kStride;
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    const T* cell = adj + ((size_t)b * cells_per_seq(L) + i) * G;
    for (int t = threadIdx.x; t < 1 + D + D * D; t += blockDim.x) {
        if (t == 0) { const double v = ld_logc(cell); if (v != 0.0) atomicAdd(&g_log_struct[0], v); }
        else if (t <= D) g_y[((size_t)b * L + i) * D + (t - 1)] = cell[H + t - 1];
        else {
            const int a = (t - 1 - D) / D, c = (t - 1 - D) % D;
            const double s = 0.5 * ((double)cell[H + t - 1] + (double)cell[H + D + c * D + a]);
            if (s != 0.0) atomicAdd(&g_cov_term[a * D + c], s);
        }
    }
}

template <int D, typename T>
static int gauss_backward_t(int B, int L, const int* lengths, const double* cov_left, const double* cov_right,
                            const double* prior_mean, const double* prior_cov, const double* log_struct, const T* chart, T* adj,
                            const double* coef, T* g_y, double* g_cov_term, double* g_cov_left,
                            double* g_cov_right, double* g_prior_mean, double* g_prior_cov, double* g_log_struct,
                            cudaStream_t st) {
    constexpr bool kF64 = std::is_same<T, double>::value;
    constexpr int G = GCell<D, T>::kStride;
    constexpr int DD = D * D;
    RBN_CUDA_CHECK(cudaMemsetAsync(adj, 0, (size_t)B * cells_per_seq(L) * G * sizeof(T), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_y, 0, (size_t)B * L * D * sizeof(T), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_cov_term, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_cov_left, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_cov_right, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_prior_mean, 0, D * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_prior_cov, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_log_struct, 0, 2 * sizeof(double), st));
    { ProfScope ps(KC_OTHER, st);
      k_gauss_bwd_root<D, T><<<(B + 63) / 64, 64, 0, st>>>(L, B, lengths, prior_mean, prior_cov, chart, coef, adj, g_prior_mean, g_prior_cov); }
    RBN_LAUNCH_CHECK("k_gauss_bwd_root");
    static const bool rows8 = gauss_rows8();
    const bool use8 = (D == 8) && (rows8 || !kF64);
    size_t smem = 0;
    if constexpr (kF64) {
        smem = ((size_t)2 * L + G + D + 4 * DD) * sizeof(double);
        RBN_CUDA_CHECK(cudaFuncSetAttribute(k_gauss_bwd_diag<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    for (int w = L; w >= 2; --w) {
        { ProfScope ps(KC_SIMT_OUTSIDE, st);
          if (use8) {
              static const int occ = gauss_occ();
              const int grid = (B * (L - w + 1) + 3) / 4;
              if (occ >= 2) g8::k_gauss_bwd_diag8<T, G8Budget<T>::kBwd2><<<grid, 128, 0, st>>>(L, w, B, lengths, cov_left, cov_right, log_struct, chart, adj,
                                                                                   g_cov_left, g_cov_right, g_log_struct);
              else if (occ == 1) g8::k_gauss_bwd_diag8<T, G8Budget<T>::kBwd><<<grid, 128, 0, st>>>(L, w, B, lengths, cov_left, cov_right, log_struct, chart, adj,
                                                                                       g_cov_left, g_cov_right, g_log_struct);
              else g8::k_gauss_bwd_diag8<T, 1><<<grid, 128, 0, st>>>(L, w, B, lengths, cov_left, cov_right, log_struct, chart, adj,
                                                                    g_cov_left, g_cov_right, g_log_struct);
          } else if constexpr (kF64) {
              k_gauss_bwd_diag<D><<<B * (L - w + 1), 32, smem, st>>>(L, w, lengths, cov_left, cov_right, chart, adj, g_cov_left, g_cov_right, g_log_struct);
          } }
        RBN_LAUNCH_CHECK("k_gauss_bwd_diag");
    }
    if (use8) {
        { ProfScope ps(KC_OTHER, st); g8::k_sym8<<<2, 64, 0, st>>>(g_cov_left, g_cov_right); }
        RBN_LAUNCH_CHECK("k_sym8");
    }
    { ProfScope ps(KC_EMISSION, st); k_gauss_bwd_width1<D, T><<<B * L, 96, 0, st>>>(L, lengths, adj, g_y, g_cov_term, g_log_struct); }
    RBN_LAUNCH_CHECK("k_gauss_bwd_width1");
    return RBN_OK;
}

int gauss_backward(int D, int B, int L, const int* lengths, const double* cov_left, const double* cov_right,
                   const double* prior_mean, const double* prior_cov, const double* log_struct, const double* chart, double* adj,
                   const double* coef, double* g_y, double* g_cov_term, double* g_cov_left, double* g_cov_right,
                   double* g_prior_mean, double* g_prior_cov, double* g_log_struct, cudaStream_t st) {
#define RBN_G(DD_) case DD_: return gauss_backward_t<DD_, double>(B, L, lengths, cov_left, cov_right, prior_mean, prior_cov, log_struct, chart, adj, coef, g_y, g_cov_term, g_cov_left, g_cov_right, g_prior_mean, g_prior_cov, g_log_struct, st)
    switch (D) {
        RBN_G(1); RBN_G(2); RBN_G(3); RBN_G(4); RBN_G(8);
        default:
            set_error("Gaussian path is instantiated for D in {1,2,3,4,8} (got %d)", D);
            return RBN_E_UNSUPPORTED;
    }
#undef RBN_G
}

int gauss_backward_f32(int D, int B, int L, const int* lengths, const double* cov_left, const double* cov_right,
                       const double* prior_mean, const double* prior_cov, const double* log_struct, const float* chart, float* adj,
                       const double* coef, float* g_y, double* g_cov_term, double* g_cov_left, double* g_cov_right,
                       double* g_prior_mean, double* g_prior_cov, double* g_log_struct, cudaStream_t st) {
    if (D != 8) {
        set_error("fp32 chart storage is instantiated for D = 8 (got %d); other D run through the fp64 entry points", D);
        return RBN_E_UNSUPPORTED;
    }
    return gauss_backward_t<8, float>(B, L, lengths, cov_left, cov_right, prior_mean, prior_cov, log_struct, chart, adj, coef, g_y,
                                      g_cov_term, g_cov_left, g_cov_right, g_prior_mean, g_prior_cov, g_log_struct, st);
}

}  // namespace rbn
