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


@pytest.mark.parametrize("D,B,L,ragged", [(8, 3, 10, True), (2, 4, 7, False), (8, 2, 24, False), (3, 2, 5, True)])
def test_gauss_gradients_vs_oracle_autograd(D, B, L, ragged):
    """d(sum coef log Z)/d(y, covariances, prior mean, structural log-probabilities): CUDA outside pass against
    torch.autograd through the fp64 oracle (covariance gradients compared symmetrised)."""
    from oracle import gauss_oracle as G
    from rbn_b200.gaussian import gauss_logZ
    g, lengths = _case(D, B, L, seed=D * 7 + L, ragged=ragged)
    coef = np.random.default_rng(1).uniform(0.5, 1.5, B)
    names = ["y", "cov_term", "cov_left", "cov_right", "prior_mean", "prior_cov", "log_struct"]
    T = {k: torch.tensor(g[k], requires_grad=True) for k in names}
    ref = G.gauss_inside(T["y"], torch.tensor(lengths), T["cov_term"], T["cov_left"], T["cov_right"], T["prior_mean"],
                         T["prior_cov"], T["log_struct"])
    (ref * torch.tensor(coef)).sum().backward()
    C = {k: torch.tensor(g[k], device="cuda", requires_grad=True) for k in names}
    lz = gauss_logZ(C["y"], C["cov_term"], C["cov_left"], C["cov_right"], C["prior_mean"], C["prior_cov"],
                    C["log_struct"], torch.tensor(lengths))
    (lz * torch.tensor(coef, device="cuda")).sum().backward()
    for k in names:
        got, want = C[k].grad.cpu().numpy(), T[k].grad.numpy()
        if k.startswith("cov") or k == "prior_cov":
            want = 0.5 * (want + want.T)
            got = 0.5 * (got + got.T)
        if k == "y":          # padding positions carry no gradient
            for b in range(B):
                want[b, lengths[b]:] = 0
        scale = max(np.abs(want).max(), 1e-12)
        assert np.abs(got - want).max() <= 1e-7 * scale, (k, np.abs(got - want).max(), scale)


def test_gaussian_rbn_module_end_to_end():
    from oracle import gauss_oracle as G
    from rbn_b200.gaussian import GaussianRBN
    g = G.synthetic_gauss(8, 3, 9, seed=5)
    m = GaussianRBN(8, g["cov_term"], g["cov_left"], g["cov_right"], g["prior_mean"], g["prior_cov"], (0.4, 0.6)).cuda()
    lz = m.log_inside_batch(torch.tensor(g["y"], device="cuda"))
    T = {k: torch.tensor(v) for k, v in g.items()}
    ref = G.gauss_inside(T["y"], torch.tensor([9, 9, 9]), T["cov_term"], T["cov_left"], T["cov_right"], T["prior_mean"],
                         T["prior_cov"], T["log_struct"])
    np.testing.assert_allclose(lz.detach().cpu().numpy(), ref.numpy(), rtol=1e-9)
    (-lz.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    z = m.inside(g["y"][0])
    assert abs(float(torch.log(z)) - float(ref[0])) < 1e-8


def test_full_size_config4_properties():
    """BASELINE configs[3] shape (D = 8, L = 64) through size-independent properties: a directional finite difference of
    sum_b log Z_b along a random perturbation of every parameter agrees with the CUDA gradient, the structural
    log-probability gradients count the rule uses (L terminal, L - 1 binary per sequence), and log Z of a sequence does
    not depend on what else is in the batch."""
    from oracle import gauss_oracle as G
    from rbn_b200.gaussian import gauss_logZ
    D, B, L = 8, 6, 64
    g = G.synthetic_gauss(D, B, L, seed=77)
    names = ["y", "cov_term", "cov_left", "cov_right", "prior_mean", "prior_cov", "log_struct"]
    C = {k: torch.tensor(g[k], device="cuda", requires_grad=True) for k in names}

    def f(t):
        return gauss_logZ(t["y"], t["cov_term"], t["cov_left"], t["cov_right"], t["prior_mean"], t["prior_cov"], t["log_struct"])

    lz = f(C)
    assert torch.isfinite(lz).all()
    lz.sum().backward()
    assert abs(float(C["log_struct"].grad[0]) - B * L) < 1e-6 * B * L
    assert abs(float(C["log_struct"].grad[1]) - B * (L - 1)) < 1e-6 * B * L
    rng = np.random.default_rng(3)
    direction, dot = {}, 0.0
    for k in names:
        d = rng.standard_normal(g[k].shape)
        if k.startswith("cov") or k == "prior_cov":
            d = 0.5 * (d + d.T)                       # stay symmetric
        direction[k] = torch.tensor(d, device="cuda")
        dot += float((C[k].grad * direction[k]).sum())
    eps = 1e-6
    with torch.no_grad():
        plus = f({k: C[k] + eps * direction[k] for k in names}).sum()
        minus = f({k: C[k] - eps * direction[k] for k in names}).sum()
    fd = float(plus - minus) / (2 * eps)
    assert abs(fd - dot) <= 1e-5 * max(abs(dot), 1.0), (fd, dot)
    with torch.no_grad():
        sub = f({k: (C[k][[4, 1]] if k == "y" else C[k]) for k in names})
    np.testing.assert_allclose(sub.cpu().numpy(), lz.detach().cpu().numpy()[[4, 1]], rtol=1e-12)


# ---- float32 charts (RBN_GAUSS_F32): the reference's default dtype ---------------------------------------------------
# Tolerances: the SAME algebra (oracle/gauss_oracle.py, i.e. rbnet.multivariate_normal's formulas) evaluated by torch in
# float32 against float64 differs by 2e-6 relative on log Z and 1.8e-3 (max-norm) on the gradients at D = 8, L = 64
# (4e-7 / 1.7e-4 at L = 24); the fp32 kernels keep the log weights in fp64 and have to stay inside those figures.
@pytest.mark.parametrize("B,L,ragged,tol_lz,tol_g", [(4, 12, True, 2e-6, 2e-4), (3, 24, False, 2e-6, 5e-4), (2, 64, False, 5e-6, 3e-3)])
def test_gauss_float32_charts_vs_fp64_oracle(B, L, ragged, tol_lz, tol_g):
    from oracle import gauss_oracle as G
    from rbn_b200.gaussian import gauss_logZ
    D = 8
    g, lengths = _case(D, B, L, seed=100 + L, ragged=ragged)
    coef = np.random.default_rng(2).uniform(0.5, 1.5, B)
    names = ["y", "cov_term", "cov_left", "cov_right", "prior_mean", "prior_cov", "log_struct"]
    T = {k: torch.tensor(g[k], requires_grad=True) for k in names}
    ref = G.gauss_inside(T["y"], torch.tensor(lengths), T["cov_term"], T["cov_left"], T["cov_right"], T["prior_mean"],
                         T["prior_cov"], T["log_struct"])
    (ref * torch.tensor(coef)).sum().backward()
    C = {k: torch.tensor(g[k], device="cuda", requires_grad=True) for k in names}
    lz = gauss_logZ(C["y"], C["cov_term"], C["cov_left"], C["cov_right"], C["prior_mean"], C["prior_cov"],
                    C["log_struct"], torch.tensor(lengths), chart_dtype=torch.float32)
    assert lz.dtype == torch.float64
    (lz * torch.tensor(coef, device="cuda")).sum().backward()
    rel = np.abs(lz.detach().cpu().numpy() - ref.detach().numpy()) / np.abs(ref.detach().numpy())
    assert rel.max() <= tol_lz, rel
    for k in names:
        got, want = C[k].grad.cpu().numpy(), T[k].grad.numpy()
        assert got.dtype == np.float64
        if k.startswith("cov") or k == "prior_cov":
            want = 0.5 * (want + want.T)
            got = 0.5 * (got + got.T)
        if k == "y":
            for b in range(B):
                want[b, lengths[b]:] = 0
        scale = max(np.abs(want).max(), 1e-12)
        assert np.abs(got - want).max() <= tol_g * scale, (k, np.abs(got - want).max() / scale)
    # the structural gradients count rule uses: exact in any precision up to the softmax normalisation
    n = np.asarray(lengths, dtype=np.float64)
    assert abs(float(C["log_struct"].grad[0]) - float((coef * n).sum())) < 1e-4 * float((coef * n).sum())


def test_gauss_float32_descriptor_and_entry_point_must_agree():
    """RBN_GAUSS_F32 on the fp64 entry point (or the reverse) is refused; D != 8 has no fp32 instantiation and the Python
    layer runs it through the float64 kernels."""
    import ctypes
    from oracle import gauss_oracle as G
    from rbn_b200 import _lib
    from rbn_b200.gaussian import gauss_logZ
    lib = _lib.load()
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    p = ctypes.c_void_p(ws.data_ptr())
    d32 = _lib.GaussDesc(D=8, B=1, L=4, flags=_lib.GAUSS_F32)
    d64 = _lib.GaussDesc(D=8, B=1, L=4, flags=0)
    assert 0 < lib.rbn_gauss_workspace_bytes(ctypes.byref(d32)) < lib.rbn_gauss_workspace_bytes(ctypes.byref(d64))
    assert lib.rbn_gauss_inside_forward(ctypes.byref(d32), *([p] * 9), 1 << 20, p, None) == -1
    assert lib.rbn_gauss_inside_forward_f32(ctypes.byref(d64), *([p] * 9), 1 << 20, p, None) == -1
    assert lib.rbn_gauss_workspace_bytes(ctypes.byref(_lib.GaussDesc(D=4, B=1, L=4, flags=_lib.GAUSS_F32))) == 0
    g = G.synthetic_gauss(4, 2, 6, seed=1)
    T = {k: torch.tensor(v) for k, v in g.items()}
    ref = G.gauss_inside(T["y"], torch.tensor([6, 6]), T["cov_term"], T["cov_left"], T["cov_right"], T["prior_mean"],
                         T["prior_cov"], T["log_struct"])
    lz = gauss_logZ(T["y"].cuda(), T["cov_term"], T["cov_left"], T["cov_right"], T["prior_mean"], T["prior_cov"],
                    T["log_struct"], chart_dtype=torch.float32)
    np.testing.assert_allclose(lz.cpu().numpy(), ref.numpy(), rtol=1e-9)
