"""CPU oracle for the continuous (multivariate-normal) RBN of BASELINE config 4.

TEST INFRASTRUCTURE ONLY (same rules as oracle/inside_oracle.py).

Parity status: the Gaussian PRIMITIVES are PINNED -- pairwise_product() and approximate_mixture() restate
/root/reference/rbnet/multivariate_normal.py:339-396 and :502-550 and are checked against the unmodified reference
classes (tests/test_gauss_oracle.py; fixtures tests/golden/gauss_primitives.npz written by
oracle/make_golden_gauss.py).  The RBN-level composition is "PARITY UNPINNED": the reference contains no
Transition / Cell / Prior that uses these primitives, so gauss_inside() below is the
composition SURVEY.md section 8a infers from the primitives, doc/recursive_bayesian_networks.rst:44-54 and the
analogous autoencoder bricks (autoencoder.py:110-131):

  beta_{i:k}(x) = c * N(x; mu, Sigma)
  terminal  p(y | x) = N(y; x, Sigma_T)                       => width-1 cell (log c = log p_S(term), mu = y_i, Sigma = Sigma_T)
  binary    p(x_l, x_r | x) = N(x_l; x, Sigma_L) N(x_r; x, Sigma_R)
            per split: PairwiseProduct(mu_l, Sigma_l + Sigma_L, mu_r, Sigma_r + Sigma_R), weight log c_l + log c_r + log_norm
            over splits: ApproximateMixture (moment matching), times p_S(binary)
  prior     N(x; mu_P, Sigma_P)                               => log Z = log c + log N(mu; mu_P, Sigma + Sigma_P)
"""
import math

import torch


def log_normal_pdf(x, mean, cov):
    """log N(x; mean, cov) for batched [..., D] / [..., D, D] (what TorchMultivariateNormal.log_prob returns in
    multivariate_normal.py:389-396)."""
    D = x.shape[-1]
    chol = torch.linalg.cholesky(cov)
    d = (x - mean).unsqueeze(-1)
    u = torch.linalg.solve_triangular(chol, d, upper=False).squeeze(-1)
    logdet = 2.0 * torch.log(torch.diagonal(chol, dim1=-2, dim2=-1)).sum(-1)
    return -0.5 * ((u * u).sum(-1) + logdet + D * math.log(2.0 * math.pi))


def pairwise_product(mean1, cov1, mean2, cov2):
    """Product of two Gaussians over the same variable (multivariate_normal.py:194-396):
    cov = cov1 (cov1+cov2)^-1 cov2 (:332-333), mean = cov2 (cov1+cov2)^-1 mean1 + cov1 (cov1+cov2)^-1 mean2 (:366-371),
    log_norm = log N(mean2; mean1, cov1+cov2) (:389-396).  Returns (log_norm, mean, cov)."""
    s = cov1 + cov2
    s_inv = torch.linalg.inv(s)
    cov = cov1 @ s_inv @ cov2
    mean = (cov2 @ s_inv @ mean1.unsqueeze(-1) + cov1 @ s_inv @ mean2.unsqueeze(-1)).squeeze(-1)
    return log_normal_pdf(mean2, mean1, s), mean, cov


def approximate_mixture(means, log_weights, covariances):
    """Moment-matched single Gaussian for a weighted mixture along dim 0 (multivariate_normal.py:502-550):
    log_norm = logsumexp(w), mean = sum w~ m, cov = sum w~ (C + (m - mean)(m - mean)^T).  Returns (log_norm, mean, cov)."""
    log_norm = torch.logsumexp(log_weights, dim=0)
    w = torch.exp(log_weights - log_norm)
    mean = (w[..., None] * means).sum(0)
    diff = means - mean[None]
    cov = (w[..., None, None] * (diff[..., :, None] * diff[..., None, :] + covariances)).sum(0)
    return log_norm, mean, cov


def gauss_inside(y, lengths, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct):
    """log Z [B] of the continuous RBN.  y: [B, L, D] float64; lengths [B]; covariances [D, D]; log_struct = (log
    p_S(terminal), log p_S(binary)).  Differentiable in every float argument (torch autograd = the outside pass)."""
    B, L, D = y.shape
    logc = {1: log_struct[0].expand(B, L).clone() if torch.is_tensor(log_struct[0]) else torch.full((B, L), float(log_struct[0]), dtype=y.dtype)}
    mu = {1: y}
    cov = {1: cov_term.expand(B, L, D, D)}
    for w in range(2, L + 1):
        n = L - w + 1
        lw, ms, cs = [], [], []
        for u in range(1, w):
            a = cov[u][:, :n] + cov_left
            b = cov[w - u][:, u:u + n] + cov_right
            ln, m, c = pairwise_product(mu[u][:, :n], a, mu[w - u][:, u:u + n], b)
            lw.append(logc[u][:, :n] + logc[w - u][:, u:u + n] + ln)
            ms.append(m)
            cs.append(c)
        ln, m, c = approximate_mixture(torch.stack(ms), torch.stack(lw), torch.stack(cs))
        logc[w] = ln + log_struct[1]
        mu[w] = m
        cov[w] = c
    out = []
    for b in range(B):
        n = int(lengths[b])
        out.append(logc[n][b, 0] + log_normal_pdf(mu[n][b, 0], prior_mean, cov[n][b, 0] + prior_cov))
    return torch.stack(out)


from rbn_b200.synth import synthetic_gauss  # noqa: E402,F401  (one generator for tests, oracle and bench)
