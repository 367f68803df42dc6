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
        { ProfScope ps(KC_SIMT_INSIDE, st);
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

// =====================================================================================================================
// outside pass of the Gaussian RBN: reverse-mode through moment matching and the pairwise products
// =====================================================================================================================
// Adjoint chart cell = [a_bar | mu_bar[D] | Sigma_bar[D][D]] of sum_b coef_b log Z_b.  One warp per span, one lane per
// split; the D x D work matrices of a lane are thread-private.
namespace rbn {

#define SMAT(M, r, c) M[(r) * D + (c)]

template <int D>
__global__ void k_gauss_bwd_root(int L, int B, const int* __restrict__ lengths, const double* __restrict__ prior_mean,
                                 const double* __restrict__ prior_cov, const double* __restrict__ chart,
                                 const double* __restrict__ coef, double* __restrict__ adj,
                                 double* __restrict__ g_prior_mean, double* __restrict__ g_prior_cov) {
    constexpr int G = GCell<D>::kStride;
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double c = coef[b];
    if (c == 0.0) return;
    const int n = lengths[b];
    const size_t cidx = (size_t)b * cells_per_seq(L) + cell_off(n, L);
    const double* cell = chart + cidx * G;
    double* out = adj + cidx * G;
    double S[D][D], d[D], z[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        d[a] = cell[1 + a] - prior_mean[a];
        z[a] = d[a];
#pragma unroll
        for (int e = 0; e < D; ++e) S[a][e] = cell[1 + D + a * D + e] + prior_cov[a * D + e];
    }
    chol_inplace<D>(S);
    chol_solve<D>(S, z);
    out[0] += c;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        out[1 + a] += -c * z[a];
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
            out[1 + D + a * D + e] += sb;
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

template <int D>
__global__ void k_gauss_bwd_width1(int L, const int* __restrict__ lengths, const double* __restrict__ adj,
                                   double* __restrict__ g_y, double* __restrict__ g_cov_term,
                                   double* __restrict__ g_log_struct) {
    constexpr int G = GCell<D>::kStride;
    int b = blockIdx.x / L, i = blockIdx.x % L;
    if (i >= lengths[b]) return;
    const double* cell = adj + ((size_t)b * cells_per_seq(L) + i) * G;
    for (int t = threadIdx.x; t < G; t += blockDim.x) {
        const double v = cell[t];
        if (t == 0) { if (v != 0.0) atomicAdd(&g_log_struct[0], v); }
        else if (t <= D) g_y[((size_t)b * L + i) * D + (t - 1)] = v;
        else {
            const int a = (t - 1 - D) / D, c = (t - 1 - D) % D;
            const double s = 0.5 * (v + cell[1 + D + c * D + a]);
            if (s != 0.0) atomicAdd(&g_cov_term[a * D + c], s);
        }
    }
}

template <int D>
static int gauss_backward_t(int B, int L, const int* lengths, const double* cov_left, const double* cov_right,
                            const double* prior_mean, const double* prior_cov, const double* chart, double* adj,
                            const double* coef, double* g_y, double* g_cov_term, double* g_cov_left,
                            double* g_cov_right, double* g_prior_mean, double* g_prior_cov, double* g_log_struct,
                            cudaStream_t st) {
    constexpr int G = GCell<D>::kStride;
    constexpr int DD = D * D;
    RBN_CUDA_CHECK(cudaMemsetAsync(adj, 0, (size_t)B * cells_per_seq(L) * G * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_y, 0, (size_t)B * L * D * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_cov_term, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_cov_left, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_cov_right, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_prior_mean, 0, D * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_prior_cov, 0, DD * sizeof(double), st));
    RBN_CUDA_CHECK(cudaMemsetAsync(g_log_struct, 0, 2 * sizeof(double), st));
    { ProfScope ps(KC_OTHER, st);
      k_gauss_bwd_root<D><<<(B + 63) / 64, 64, 0, st>>>(L, B, lengths, prior_mean, prior_cov, chart, coef, adj, g_prior_mean, g_prior_cov); }
    RBN_LAUNCH_CHECK("k_gauss_bwd_root");
    size_t smem = ((size_t)2 * L + G + D + 4 * DD) * sizeof(double);
    RBN_CUDA_CHECK(cudaFuncSetAttribute(k_gauss_bwd_diag<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int w = L; w >= 2; --w) {
        { ProfScope ps(KC_SIMT_OUTSIDE, st);
          k_gauss_bwd_diag<D><<<B * (L - w + 1), 32, smem, st>>>(L, w, lengths, cov_left, cov_right, chart, adj, g_cov_left, g_cov_right, g_log_struct); }
        RBN_LAUNCH_CHECK("k_gauss_bwd_diag");
    }
    { ProfScope ps(KC_EMISSION, st); k_gauss_bwd_width1<D><<<B * L, 96, 0, st>>>(L, lengths, adj, g_y, g_cov_term, g_log_struct); }
    RBN_LAUNCH_CHECK("k_gauss_bwd_width1");
    return RBN_OK;
}

int gauss_backward(int D, int B, int L, const int* lengths, const double* cov_left, const double* cov_right,
                   const double* prior_mean, const double* prior_cov, const double* chart, double* adj,
                   const double* coef, double* g_y, double* g_cov_term, double* g_cov_left, double* g_cov_right,
                   double* g_prior_mean, double* g_prior_cov, double* g_log_struct, cudaStream_t st) {
#define RBN_G(DD_) case DD_: return gauss_backward_t<DD_>(B, L, lengths, cov_left, cov_right, prior_mean, prior_cov, chart, adj, coef, g_y, g_cov_term, g_cov_left, g_cov_right, g_prior_mean, g_prior_cov, g_log_struct, st)
    switch (D) {
        RBN_G(1); RBN_G(2); RBN_G(3); RBN_G(4); RBN_G(8);
        default:
            set_error("Gaussian path is instantiated for D in {1,2,3,4,8} (got %d)", D);
            return RBN_E_UNSUPPORTED;
    }
#undef RBN_G
}

}  // namespace rbn
