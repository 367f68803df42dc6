"""Continuous (multivariate-normal) RBN on the GPU against the fp64 oracle composition (config 4 shape, reduced)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(D, B, L, seed, ragged):
    from oracle import gauss_oracle as G
    g = G.synthetic_gauss(D, B, L, seed=seed)
    lengths = np.full(B, L)
    if ragged:
        lengths = np.random.default_rng(seed).integers(1, L + 1, B)
    return g, lengths


@pytest.mark.parametrize("D,B,L,ragged", [(8, 4, 12, True), (8, 3, 64, False), (2, 5, 9, True), (1, 2, 5, False), (4, 2, 130, True)])
def test_gauss_inside_forward_vs_oracle(D, B, L, ragged):
    from oracle import gauss_oracle as G
    from rbn_b200.gaussian import gauss_logZ
    g, lengths = _case(D, B, L, seed=D * 10 + L, ragged=ragged)
    T = {k: torch.tensor(v) for k, v in g.items()}
    ref = G.gauss_inside(T["y"], torch.tensor(lengths), T["cov_term"], T["cov_left"], T["cov_right"], T["prior_mean"],
                         T["prior_cov"], T["log_struct"])
    C = {k: v.cuda() for k, v in T.items()}
    lz = gauss_logZ(C["y"], C["cov_term"], C["cov_left"], C["cov_right"], C["prior_mean"], C["prior_cov"],
                    C["log_struct"], torch.tensor(lengths))
    np.testing.assert_allclose(lz.cpu().numpy(), ref.numpy(), rtol=1e-9, atol=1e-9)
