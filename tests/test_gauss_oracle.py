"""Gaussian oracle primitives against fixtures produced by the unmodified reference classes."""
import os

import numpy as np
import torch

from oracle import gauss_oracle as G

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gauss_primitives.npz")


def test_pairwise_product_and_mixture_match_reference():
    s = np.load(FIX)
    for k in range(int(s["n"])):
        t = lambda name: torch.tensor(s[f"pp{k}/{name}"])
        ln, mean, cov = G.pairwise_product(t("m1"), t("c1"), t("m2"), t("c2"))
        np.testing.assert_allclose(ln.numpy(), s[f"pp{k}/log_norm"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(mean.numpy(), s[f"pp{k}/mean"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(cov.numpy(), s[f"pp{k}/cov"], rtol=1e-9, atol=1e-11)
        # Product.product() of two factors is the same thing (multivariate_normal.py:462-471)
        np.testing.assert_allclose(ln.numpy(), s[f"pp{k}/product_log_norm"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(mean.numpy(), s[f"pp{k}/product_mean"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(cov.numpy(), s[f"pp{k}/product_cov"], rtol=1e-8, atol=1e-10)
        a = lambda name: torch.tensor(s[f"am{k}/{name}"])
        ln, mean, cov = G.approximate_mixture(a("means"), a("lw"), a("covs"))
        np.testing.assert_allclose(ln.numpy(), s[f"am{k}/log_norm"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(mean.numpy(), s[f"am{k}/mean"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(cov.numpy(), s[f"am{k}/cov"], rtol=1e-10, atol=1e-12)


def test_gauss_inside_length_one_and_two_closed_form():
    D = 3
    g = G.synthetic_gauss(D, 2, 2, seed=3)
    T = {k: torch.tensor(v) for k, v in g.items()}
    lz = G.gauss_inside(T["y"], torch.tensor([1, 2]), T["cov_term"], T["cov_left"], T["cov_right"], T["prior_mean"],
                        T["prior_cov"], T["log_struct"])
    # length 1: log p_S(term) + log N(y0; mu_P, Sigma_T + Sigma_P)
    ref0 = T["log_struct"][0] + G.log_normal_pdf(T["y"][0, 0], T["prior_mean"], T["cov_term"] + T["prior_cov"])
    assert abs(float(lz[0] - ref0)) < 1e-12
    # length 2: one split; integrate x analytically
    a, b = T["cov_term"] + T["cov_left"], T["cov_term"] + T["cov_right"]
    ln, m, c = G.pairwise_product(T["y"][1, 0], a, T["y"][1, 1], b)
    ref1 = 2 * T["log_struct"][0] + T["log_struct"][1] + ln + G.log_normal_pdf(m, T["prior_mean"], c + T["prior_cov"])
    assert abs(float(lz[1] - ref1)) < 1e-12
