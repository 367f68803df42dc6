"""`rbn_b200.multivariate_normal.MultivariateNormal`: the constructor contract of
/root/reference/rbnet/multivariate_normal.py:96-163 restated from the reference's own test
(/root/reference/tests/test_multivariate_normal.py:20-66: every combination of scalar / full location, scalar /
diagonal / full matrix, covariance / precision / scale_tril must build the same torch distribution as the expanded
arguments), and `__mul__` against the fixtures the unmodified reference's PairwiseProduct wrote."""
import os

import numpy as np
import pytest
import torch
from torch.distributions import MultivariateNormal as TorchMVN

from rbn_b200.multivariate_normal import MultivariateNormal, pairwise_product

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gauss_primitives.npz")


def test_init_matches_torch_for_every_parametrisation():
    rng = np.random.default_rng(0)
    for _ in range(10):
        batch = tuple(int(x) for x in rng.integers(1, 4, rng.integers(0, 4)))
        dim = int(rng.integers(1, 5))
        for loc_dim in (False, dim):
            if loc_dim:
                loc = torch.tensor(rng.uniform(-1, 1, batch), dtype=torch.float32)
                full_loc = loc[..., None].expand(batch + (dim,))
            else:
                loc = torch.tensor(rng.uniform(-1, 1, batch + (dim,)), dtype=torch.float32)
                full_loc = loc
            for kind in ("FULL", "DIAG", "SCAL"):
                if kind == "SCAL":
                    mat = torch.tensor(rng.uniform(0.1, 1, batch), dtype=torch.float32) ** 2
                    full = MultivariateNormal._expand_scalar(mat, dim)
                elif kind == "DIAG":
                    mat = torch.tensor(rng.uniform(0.1, 1, batch + (dim,)), dtype=torch.float32) ** 2
                    full = MultivariateNormal._expand_diagonal(mat, dim)
                else:
                    a = torch.tensor(rng.uniform(-1, 1, batch + (dim, dim)), dtype=torch.float32)
                    mat = a @ a.transpose(-1, -2) + 0.1 * torch.eye(dim)
                    full = mat
                tri = torch.linalg.cholesky(mat) if kind == "FULL" else torch.sqrt(mat)
                full_tri = (tri if kind == "FULL" else MultivariateNormal._expand_diagonal(tri, dim) if kind == "DIAG"
                            else MultivariateNormal._expand_scalar(tri, dim))
                for name, m, fm in (("covariance_matrix", mat, full), ("precision_matrix", mat, full), ("scale_tril", tri, full_tri)):
                    n = MultivariateNormal(loc=loc, dim=loc_dim, validate_args=True, **{name: m})
                    t = TorchMVN(loc=full_loc, validate_args=True, **{name: fm})
                    assert n.mat_type == {"FULL": MultivariateNormal.FULL, "DIAG": MultivariateNormal.DIAG, "SCAL": MultivariateNormal.SCAL}[kind]
                    np.testing.assert_array_equal(n.torch.loc.numpy(), t.loc.numpy())
                    np.testing.assert_array_equal(n.torch._unbroadcasted_scale_tril.numpy(), t._unbroadcasted_scale_tril.numpy())
                    assert n.event_dim == dim


def test_constructor_errors():
    loc = torch.zeros(3)
    with pytest.raises(ValueError):
        MultivariateNormal(loc=loc)                                             # no matrix
    with pytest.raises(ValueError):
        MultivariateNormal(loc=loc, covariance_matrix=torch.eye(3), scale_tril=torch.eye(3))
    with pytest.raises(ValueError):
        MultivariateNormal(loc=loc, covariance_matrix=torch.eye(3), norm=1., log_norm=0.)
    with pytest.raises(ValueError):
        MultivariateNormal._expand_diagonal(torch.ones(2), 3)
    assert MultivariateNormal(loc=loc, covariance_matrix=torch.eye(3)).__mul__(3.0) is NotImplemented


def test_mul_is_the_references_pairwise_product():
    s = np.load(FIX)
    for k in range(int(s["n"])):
        t = lambda name: torch.tensor(s[f"pp{k}/{name}"])
        a = MultivariateNormal(loc=t("m1"), covariance_matrix=t("c1"), log_norm=torch.tensor(0.25, dtype=torch.float64))
        b = MultivariateNormal(loc=t("m2"), scale_tril=torch.linalg.cholesky(t("c2")), norm=torch.tensor(2.0, dtype=torch.float64))
        p = a * b
        np.testing.assert_allclose(p.loc.numpy(), s[f"pp{k}/mean"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(p.covariance_matrix.numpy(), s[f"pp{k}/cov"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose((p.log_norm - 0.25 - np.log(2.0)).numpy(), s[f"pp{k}/log_norm"], rtol=1e-9, atol=1e-10)
        mean, cov, log_c = pairwise_product(t("m1"), t("c1"), t("m2"), t("c2"))
        np.testing.assert_allclose(log_c.numpy(), s[f"pp{k}/log_norm"], rtol=1e-9, atol=1e-11)
    # scalar x diagonal parametrisations multiply too (densities over the same 3-dimensional variable)
    a = MultivariateNormal(loc=torch.tensor([0., 1., 2.], dtype=torch.float64), covariance_matrix=torch.tensor(0.5, dtype=torch.float64))
    b = MultivariateNormal(loc=torch.tensor(1.0, dtype=torch.float64), covariance_matrix=torch.tensor([1., 2., 3.], dtype=torch.float64), dim=3)
    p = a * b
    x = torch.tensor([0.3, -0.2, 1.1], dtype=torch.float64)
    assert abs(float(p.log_prob(x) - a.log_prob(x) - b.log_prob(x))) < 1e-12


def test_gaussian_rbn_takes_scalar_diagonal_and_full_covariances():
    from rbn_b200.gaussian import GaussianRBN
    D = 3
    full = np.array([[2.0, 0.3, 0.0], [0.3, 1.0, 0.1], [0.0, 0.1, 0.5]])
    m = GaussianRBN(D, cov_term=0.25, cov_left=[1.0, 2.0, 3.0], cov_right=full,
                    prior_cov=MultivariateNormal(loc=torch.zeros(D), covariance_matrix=torch.tensor(4.0)))
    np.testing.assert_allclose(m._cov(m.term_scale_tril).detach().numpy(), 0.25 * np.eye(D), atol=1e-12)
    np.testing.assert_allclose(m._cov(m.left_scale_tril).detach().numpy(), np.diag([1.0, 2.0, 3.0]), atol=1e-12)
    np.testing.assert_allclose(m._cov(m.right_scale_tril).detach().numpy(), full, atol=1e-12)
    np.testing.assert_allclose(m._cov(m.prior_scale_tril).detach().numpy(), 4.0 * np.eye(D), atol=1e-6)
    assert m.term_scale_tril.dim() == 0 and m.left_scale_tril.dim() == 1 and m.right_scale_tril.dim() == 2
    with pytest.raises(ValueError):
        GaussianRBN(D, cov_term=[1.0, 2.0])
