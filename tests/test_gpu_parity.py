"""Parity of the CUDA chart pass (through the C ABI) with the reference fixtures and the fp64 oracle.

Tolerances (BASELINE.json north_star): log marginal likelihood within 1e-4 relative; rule gradients / expected
counts within 1e-3 relative (max-norm relative, fp32 accumulation against the fp64 oracle)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGZ_RTOL = 1e-4
GRAD_RTOL = 1e-3


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_reference_golden_cases_Z_chart_and_gradients(golden):
    """Every case produced by the unmodified reference: Z, the TMap-layout chart per variable, and the gradients
    left on the Parameters by Z.backward() (after the projection hooks)."""
    from tests.golden_util import build_model, case_prob_rep
    for key, c in golden.items():
        model = build_model(c, case_prob_rep(c))
        Z = model.inside(sequence=c["sequence"])
        assert Z.dtype == torch.float64 and Z.dim() == 0
        if c["Z"] == 0:
            assert float(Z) == 0.0, key
        else:
            assert abs(float(Z) - c["Z"]) <= 2e-5 * c["Z"], (key, float(Z), c["Z"])
            assert abs(math.log(float(Z)) - math.log(c["Z"])) <= LOGZ_RTOL * abs(math.log(c["Z"])) + 1e-6, key
        for v, ref in enumerate(c["charts"]):
            got = model.inside_chart[v].arr.double().numpy()
            assert got.shape == ref.shape, key
            assert _rel(got, ref) <= 2e-5, (key, v)
            assert np.all((got == 0) == (ref == 0)), (key, "zero pattern")
        if c["grads"]:
            model.zero_grad()
            Z.backward()
            for name, p in model.named_parameters():
                g = p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape))
                assert _rel(g, c["grads"][name]) <= GRAD_RTOL, (key, name)


def test_reference_test_suite_semantics():
    """The reference's own hot-path tests, restated against this package (tests/test_pcfg.py:21-35,
    tests/test_base.py:47-77 and 250-295)."""
    from rbn_b200 import pcfg, sequential, util
    g = pcfg.AbstractedPCFG(non_terminals="SAB", terminals="ab", start="S", rules=[
        ("S --> A B", 1), ("S --> B A", 1), ("A --> B A", 1), ("B --> A B", 1), ("A --> a", 1), ("B --> b", 1)],
        prob_rep=util.Prob)
    assert g.inside(sequence="aaaa") == 0
    assert g.inside(sequence="bbbb") == 0
    assert g.inside(sequence="bbba") == 1 / 2 ** 7
    assert g.inside(sequence="aaab") == 1 / 2 ** 7
    chart = np.array([[1 / 2 ** 7, 0, 1 / 2 ** 7], [0, 0, 0], [1 / 2 ** 5, 0, 1 / 2 ** 5], [0, 0, 0], [0, 0, 0],
                      [1 / 2 ** 3, 0, 1 / 2 ** 3], [0, 1 / 2, 0], [0, 1 / 2, 0], [0, 1 / 2, 0], [0, 0, 1 / 2]])
    np.testing.assert_array_equal(g.inside_chart[0].arr.detach().numpy(), chart)
    assert g.map_inside_chart()[0, 4] == ["S:7.8125e-03", "B:7.8125e-03"]
    # minimal grammar: closed-form levels, exact
    rbn = sequential.SequentialRBN(
        cells=[pcfg.StaticCell(variable=pcfg.DiscreteNonTermVar(2), weights=np.ones(2), transitions=[
            pcfg.DiscreteTerminalTransition(weights=np.ones((2, 2)), prob_rep=util.Prob),
            pcfg.DiscreteBinaryNonTerminalTransition(weights=np.ones((2, 2, 2)), prob_rep=util.Prob)], prob_rep=util.Prob)],
        prior=pcfg.DiscretePrior(struc_weights=np.ones(1), prior_weights=[np.ones(2)], prob_rep=util.Prob))
    n = 5
    Z = rbn.inside(sequence=np.random.default_rng(0).integers(0, 2, (n, 1)))
    lvl = np.zeros(n)
    for k in range(n):
        lvl[k] = 0.25 if k == 0 else (lvl[:k] * np.flip(lvl[:k])).sum() * 0.5
        for s in range(n - k):
            np.testing.assert_array_equal(rbn.inside_chart[0][s, s + k + 1].numpy(), np.ones(2) * lvl[k])
    assert float(Z) == lvl[-1]
    assert rbn.root_location == (0, n) and rbn.n == n and len(rbn.terminal_chart) == n


def _flat_case(N, V, L, B, ragged, seed, path, nt=1):
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ
    W, T, pr = O.synthetic_grammar(N, V, seed=seed)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=seed + 1, ragged=ragged)
    coef = np.random.default_rng(seed + 2).uniform(0.5, 1.5, B)
    lz_ref, gW_ref, gT_ref, gP_ref = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths, grad_logZ=coef)
    Wt = torch.tensor(W, requires_grad=True)
    Tt = torch.tensor(T, requires_grad=True)
    pt = torch.tensor(pr, requires_grad=True)
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=path)
    (lz * torch.tensor(coef)).sum().backward()
    assert np.abs(lz.detach().numpy() - lz_ref.numpy()).max() <= LOGZ_RTOL * np.abs(lz_ref.numpy()).max()
    return (_rel(Wt.grad.numpy(), gW_ref.numpy()), _rel(Tt.grad.numpy(), gT_ref.numpy()),
            _rel(pt.grad.numpy(), gP_ref.numpy()), float(np.abs(lz.detach().numpy() - lz_ref.numpy()).max()))


@pytest.mark.parametrize("N,V,L,B,ragged", [(10, 19, 20, 3, False), (7, 5, 33, 4, True), (33, 8, 12, 2, True),
                                             (64, 16, 9, 2, False), (1, 1, 6, 2, True), (300, 4, 4, 1, False)])
def test_simt_path_vs_oracle(N, V, L, B, ragged):
    from rbn_b200 import _lib
    eW, eT, eP, elz = _flat_case(N, V, L, B, ragged, seed=11 + N, path=_lib.PATH_SIMT)
    assert eW <= GRAD_RTOL and eT <= GRAD_RTOL and eP <= GRAD_RTOL, (eW, eT, eP, elz)


def test_expected_count_identities_long_sequences():
    """Size-independent property at lengths where linear space underflows (L = 128): raw counts sum to L-1
    (binary), L (terminal), 1 (prior) per sequence."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    N, V, L, B = 12, 32, 128, 6
    W, T, pr = O.synthetic_grammar(N, V, seed=5)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=6, ragged=True)
    Wt, Tt, pt = (torch.tensor(x, requires_grad=True) for x in (W, T, pr))
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_SIMT)
    assert torch.isfinite(lz).all() and float(lz.max()) < -100
    lz.sum().backward()
    assert abs(float((Wt.grad * Wt).sum()) - float((lengths - 1).sum())) <= 1e-3 * float(lengths.sum())
    assert abs(float((Tt.grad * Tt).sum()) - float(lengths.sum())) <= 1e-3 * float(lengths.sum())
    assert abs(float((pt.grad * pt).sum()) - B) <= 1e-3 * B


def test_batched_api_and_multivariable_rbn(golden):
    """log_inside_batch over ragged reference-style sequences equals per-sequence inside(); covers several cells,
    several terminal variables and None terminals (tests/test_pcfg.py:37-81 shape)."""
    from tests.golden_util import build_model
    from rbn_b200 import util
    c = golden["randmulti_LogProb/3"]
    model = build_model(c, util.LogProb)
    seq = c["sequence"]
    seqs = [seq, seq[:3], seq[2:], seq[:1]]
    lz = model.log_inside_batch(sequences=seqs)
    assert lz.shape == (4,) and lz.dtype == torch.float64
    for k, s in enumerate(seqs):
        z = float(model.inside(sequence=s))
        assert abs(float(lz[k]) - math.log(z)) <= 1e-5 * abs(math.log(z)) + 1e-6
    assert abs(float(lz[0]) - math.log(c["Z"])) <= LOGZ_RTOL * abs(math.log(c["Z"]))
    lz.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_lognormalize_kernel():
    import ctypes
    from rbn_b200 import _lib
    lib = _lib.load()
    x = torch.randn(5, 7, 33, device="cuda")
    ref = x - torch.logsumexp(x, dim=(0, 1), keepdim=True)
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.rbn_lognormalize(ctypes.c_void_p(x.data_ptr()), 35, 33, ctypes.c_void_p(bad.data_ptr()),
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.allclose(x, ref, atol=1e-5) and int(bad) == 0
