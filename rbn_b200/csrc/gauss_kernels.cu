// Continuous (multivariate-normal) RBN: inside pass over span diagonals in fp64.
//
// Every chart cell is a scaled Gaussian  beta(x) = exp(logc) * N(x; mu, Sigma)  stored as [logc | mu[D] | Sigma[D][D]].
// Per split the two children are multiplied as in rbnet.multivariate_normal.PairwiseProduct
// (/root/reference/rbnet/multivariate_normal.py:339-396) after adding the transition noise, and the splits of a span
// are collapsed by moment matching as in ApproximateMixture (multivariate_normal.py:502-550).  The composition is
// the one of oracle/gauss_oracle.py (the reference has no Gaussian RBN bricks).
//
// One CTA per span, one thread per split; D x D (D <= 8) Cholesky factorisations live in registers.  The pass is
// latency-bound (L-1 dependent diagonals of tiny dense algebra), so it runs in fp64: parity with the fp64 oracle is
// then limited by summation order only.
#include "common.cuh"
#include <math.h>

namespace rbn {

template <int D>
struct GCell {
    static constexpr int kStride = 1 + D + D * D;
};

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

template <int D>
__global__ void k_gauss_width1(int L, const double* __restrict__ y, const int* __restrict__ lengths,
                               const double* __restrict__ cov_term, const double* __restrict__ log_struct,
                               double* __restrict__ chart) {
    constexpr int G = GCell<D>::kStride;
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    double* cell = chart + ((size_t)b * cells_per_seq(L) + i) * G;
    for (int t = threadIdx.x; t < G; t += blockDim.x) {
        double v;
        if (t == 0) v = log_struct[0];
        else if (t <= D) v = y[((size_t)b * L + i) * D + (t - 1)];
        else v = cov_term[t - 1 - D];
        cell[t] = v;
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
    for (int u = 1 + threadIdx.x; u < w; u += blockDim.x) {
        const double* l = seq + (size_t)(cell_off(u, L) + i) * G;
        const double* r = seq + (size_t)(cell_off(w - u, L) + i + u) * G;
        split_product<D, true>(l, r, cL, cR, mean, A, cov);
        const double wt = exp(lw[u] - log_norm);
#pragma unroll
        for (int a = 0; a < D; ++a) {
            atomicAdd(&acc[a], wt * mean[a]);
#pragma unroll
            for (int c = 0; c < D; ++c) atomicAdd(&acc[D + a * D + c], wt * (cov[a][c] + mean[a] * mean[c]));
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

template <int D>
__global__ void k_gauss_root(int L, int B, const int* __restrict__ lengths, const double* __restrict__ prior_mean,
                             const double* __restrict__ prior_cov, const double* __restrict__ chart,
                             double* __restrict__ logZ) {
    constexpr int G = GCell<D>::kStride;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int n = lengths[b];
    const double* cell = chart + ((size_t)b * cells_per_seq(L) + cell_off(n, L)) * G;
    double S[D][D], d[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        d[a] = cell[1 + a] - prior_mean[a];
#pragma unroll
        for (int c = 0; c < D; ++c) S[a][c] = cell[1 + D + a * D + c] + prior_cov[a * D + c];
    }
    const double logdet = chol_inplace<D>(S);
    double z[D];
#pragma unroll
    for (int a = 0; a < D; ++a) z[a] = d[a];
    chol_solve<D>(S, z);
    double quad = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) quad += d[a] * z[a];
    logZ[b] = cell[0] - 0.5 * (quad + logdet + D * 1.8378770664093454836);
}

template <int D>
static int gauss_forward_t(int B, int L, const double* y, const int* lengths, const double* cov_term,
                           const double* cov_left, const double* cov_right, const double* prior_mean,
                           const double* prior_cov, const double* log_struct, double* chart, double* logZ,
                           cudaStream_t st) {
    { ProfScope ps(KC_EMISSION, st); k_gauss_width1<D><<<B * L, 96, 0, st>>>(L, y, lengths, cov_term, log_struct, chart); }
    RBN_LAUNCH_CHECK("k_gauss_width1");
    size_t smem = ((size_t)L + D + 3 * D * D) * sizeof(double);
    for (int w = 2; w <= L; ++w) {
        { ProfScope ps(KC_OTHER, st);
          k_gauss_diag<D><<<B * (L - w + 1), 64, smem, st>>>(L, w, lengths, cov_left, cov_right, log_struct, chart); }
        RBN_LAUNCH_CHECK("k_gauss_diag");
    }
    { ProfScope ps(KC_OTHER, st); k_gauss_root<D><<<(B + 63) / 64, 64, 0, st>>>(L, B, lengths, prior_mean, prior_cov, chart, logZ); }
    RBN_LAUNCH_CHECK("k_gauss_root");
    return RBN_OK;
}

size_t gauss_workspace_bytes(int D, int B, int L) {
    return align_up((size_t)B * cells_per_seq(L) * (1 + D + D * D) * sizeof(double), 1024);
}

int gauss_forward(int D, int B, int L, const double* y, const int* lengths, const double* cov_term,
                  const double* cov_left, const double* cov_right, const double* prior_mean, const double* prior_cov,
                  const double* log_struct, double* chart, double* logZ, cudaStream_t st) {
#define RBN_G(DD) case DD: return gauss_forward_t<DD>(B, L, y, lengths, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct, chart, logZ, st)
    switch (D) {
        RBN_G(1); RBN_G(2); RBN_G(3); RBN_G(4); RBN_G(8);
        default:
            set_error("Gaussian path is instantiated for D in {1,2,3,4,8} (got %d)", D);
            return RBN_E_UNSUPPORTED;
    }
#undef RBN_G
}

}  // namespace rbn
