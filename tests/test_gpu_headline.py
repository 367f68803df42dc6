"""Parity AT THE BENCHMARKED SHAPES: BASELINE configs[2] (N = 256, L = 128; one full-length and one ragged sequence)
and configs[4] (N = 512, L = 128) through the tensor-core path against the fp64 oracle
(oracle/inside_oracle.py::flat_inside_with_grads = /root/reference/rbnet/pcfg.py:305-314 + base.py:93-121 with
torch.autograd as the outside pass), entry by entry.

Tolerances (north_star): log Z 1e-4 relative; gradients / expected counts 1e-3 relative, stated three ways:
  max-norm            max|d - ref| / max|ref|
  relative L2         |d - ref|_2 / |ref|_2
  per-entry relative  max over entries with |ref| >= FLOOR * max|ref| of |d - ref| / |ref|
The oracle needs 30 s (C3) to several minutes (C5) per sequence, so the default run compares against fixtures the
oracle wrote in the build container (tests/golden/headline_*.npz, script oracle/make_golden_headline.py: log Z, gT and
gPrior in full, a fixed arithmetic sample of ~130 k entries of gW, its three marginals, its norm and sum);
RBN_FULL_ORACLE=1 re-runs the oracle live and compares all N^3 entries of gW.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOGZ_RTOL = 1e-4
GRAD_RTOL = 1e-3
FLOOR = 1e-4          # per-entry relative error is taken over entries >= FLOOR * max|ref|


def metrics(d, ref):
    d, ref = np.asarray(d, dtype=np.float64).ravel(), np.asarray(ref, dtype=np.float64).ravel()
    mx = np.abs(ref).max()
    big = np.abs(ref) >= FLOOR * mx
    return dict(max_rel=float(np.abs(d - ref).max() / mx),
                l2_rel=float(np.linalg.norm(d - ref) / np.linalg.norm(ref)),
                entry_rel=float((np.abs(d - ref)[big] / np.abs(ref)[big]).max()),
                entries_above_floor=int(big.sum()), entries=int(ref.size))


def _gpu(case):
    from oracle import make_golden_headline as G
    from rbn_b200 import inside_logZ, _lib
    W, T, pr, tok, lengths = G.inputs(case)
    Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_TCGEN05)
    lz.sum().backward()
    torch.cuda.synchronize()
    return (W, T, pr, tok, lengths), lz.detach().cpu().numpy(), Wt.grad.cpu().numpy(), Tt.grad.cpu().numpy(), pt.grad.cpu().numpy()


def _record(name, report):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"parity_{name}.json"), "w") as f:
            json.dump(report, f, indent=1)
    except OSError:
        pass


@pytest.mark.parametrize("name", ["headline_c3", "headline_c5"])
def test_headline_shape_vs_oracle(name):
    from oracle import make_golden_headline as G
    case = G.CASES[name]
    inp, lz, gW, gT, gP = _gpu(case)
    report = {}
    if os.environ.get("RBN_FULL_ORACLE") == "1":
        from oracle import inside_oracle as O
        W, T, pr, tok, lengths = inp
        lz_ref, gW_ref, gT_ref, gP_ref = (x.numpy() for x in O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths))
        report["mode"] = "live oracle, all entries"
        report["gW"] = metrics(gW, gW_ref)
    else:
        fx = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        assert list(fx["lengths"]) == list(case["lengths"]) and list(fx["meta"][:3]) == [case["N"], case["V"], case["L"]]
        lz_ref, gT_ref, gP_ref = fx["logZ"], fx["gT"], fx["gP"]
        st = G.grad_stats(gW, int(fx["stride"]))
        report["mode"] = "fixture (oracle run in the build container)"
        report["gW"] = metrics(st["sample"], fx["gW_sample"])                 # ~130 k entries, fixed arithmetic sample
        report["gW_marginals"] = metrics(st["marg"], fx["gW_marg"])
        report["gW_norm_rel"] = abs(st["l2"] - float(fx["gW_l2"])) / float(fx["gW_l2"])
        report["gW_sum_rel"] = abs(st["total"] - float(fx["gW_total"])) / abs(float(fx["gW_total"]))
        report["gW_amax_rel"] = abs(st["amax"] - float(fx["gW_amax"])) / float(fx["gW_amax"])
    report["logZ"] = [float(x) for x in lz]
    report["logZ_ref"] = [float(x) for x in lz_ref]
    report["logZ_rel"] = float(np.max(np.abs(lz - lz_ref) / np.abs(lz_ref)))
    report["gT"] = metrics(gT, gT_ref)
    report["gPrior"] = metrics(gP, gP_ref)
    _record(name, report)
    print(json.dumps(report))
    assert report["logZ_rel"] <= LOGZ_RTOL, report
    for key in ("gW", "gT", "gPrior") + (("gW_marginals",) if "gW_marginals" in report else ()):
        m = report[key]
        assert m["max_rel"] <= GRAD_RTOL and m["l2_rel"] <= GRAD_RTOL and m["entry_rel"] <= GRAD_RTOL, (key, report)
    for key in ("gW_norm_rel", "gW_sum_rel", "gW_amax_rel"):
        if key in report:
            assert report[key] <= GRAD_RTOL, (key, report)


@pytest.mark.parametrize("N,L,B,chunk", [(64, 24, 5, 128), (256, 12, 3, 0)])
def test_split_sum_cache_does_not_change_results(monkeypatch, N, L, B, chunk):
    """RbnDesc.cache_bytes (the inside pass keeps Y for the outside pass) is a pure speed / space trade: whole, partial
    (only the widest diagonals fit) and no cache give the same gradients (the kept spans' expected counts come from ONE
    count GEMM at the end instead of one per diagonal, so sums are ordered differently); workspace memory is poisoned with
    NaN bytes first, so a span beyond its sequence whose Y was never written would show."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    W, T, pr = O.synthetic_grammar(N, 16, seed=3)
    tok, lengths = O.synthetic_tokens(B, L, 16, seed=4, ragged=True)
    monkeypatch.setenv("RBN_DEBUG_POISON", "1")
    per_span = N * N * 2 + 4
    spans = B * L * (L - 1) // 2
    res = []
    for cache_gb in (0.0, 0.37 * spans * per_span / 1e9, 1.0):
        monkeypatch.setenv("RBN_Y_CACHE_GB", repr(cache_gb))
        Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
        lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_TCGEN05, chunk_items=chunk)
        lz.sum().backward()
        res.append((lz.detach().cpu().numpy(), Wt.grad.cpu().numpy(), Tt.grad.cpu().numpy(), pt.grad.cpu().numpy()))
    assert np.isfinite(res[0][1]).all()
    for r in res[1:]:         # equal up to the order of the atomic adds of K-split partial sums (and of the deferred count GEMM)
        np.testing.assert_allclose(r[0], res[0][0], rtol=1e-6)
        for a, b in zip(res[0][1:], r[1:]):
            assert metrics(b, a)["max_rel"] <= 2e-4 and metrics(b, a)["l2_rel"] <= 2e-4
    ref, gW, gT, gP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
    assert np.abs(res[0][0] - ref.numpy()).max() <= LOGZ_RTOL * np.abs(ref.numpy()).max()
    assert metrics(res[0][1], gW.numpy())["max_rel"] <= GRAD_RTOL


@pytest.mark.parametrize("N,L,B,cache_gb,against", [(64, 130, 2, None, "oracle"), (64, 200, 2, "0", "oracle"),
                                                   (128, 150, 3, None, "simt"), (64, 300, 1, None, "simt"),
                                                   (128, 260, 2, "0.02", "simt"), (64, 513, 1, None, "simt")])
def test_sequences_longer_than_129_tokens_on_the_tensor_core_path(monkeypatch, N, L, B, cache_gb, against):
    """A split-sum / child-gradient launch holds 128 splits of a span; spans wider than 129 tokens run one launch per window
    of 128 splits, the partial split-sums (each under its own scale exponent) are folded by k_y_combine (L >= 130
    used to be refused on this path).  Whole / partial / no split-sum cache, ragged batch, NaN-poisoned workspace; two to
    four windows.  Reference: the fp64 oracle where it is affordable on the GPU box's host, else the CUDA-core path (itself
    checked against the oracle at these lengths, tests/test_gpu_parity.py and the first two cases here)."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    V = 8
    W, T, pr = O.synthetic_grammar(N, V, seed=11)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=12, ragged=B > 2)
    lengths[0] = L
    tok[0] = np.random.default_rng(5).integers(0, V, L)
    monkeypatch.setenv("RBN_DEBUG_POISON", "1")
    if cache_gb is not None:
        monkeypatch.setenv("RBN_Y_CACHE_GB", cache_gb)

    def run(path):
        Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
        lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=path)
        lz.sum().backward()
        return lz.detach().cpu().numpy(), Wt.grad.cpu().numpy(), Tt.grad.cpu().numpy(), pt.grad.cpu().numpy()

    lz, gW_, gT_, gP_ = run(_lib.PATH_TCGEN05)
    if against == "oracle":
        ref, gW, gT, gP = (x.numpy() for x in O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths))
    else:
        ref, gW, gT, gP = run(_lib.PATH_SIMT)
    assert np.abs(lz - ref).max() <= LOGZ_RTOL * np.abs(ref).max()
    for got, want in ((gW_, gW), (gT_, gT), (gP_, gP)):
        m = metrics(got, want)
        assert m["max_rel"] <= GRAD_RTOL and m["l2_rel"] <= GRAD_RTOL, m
    assert abs(float((gW_.astype(np.float64) * W).sum()) - float((lengths - 1).sum())) <= 1e-3 * float((lengths - 1).sum())


def test_tensor_core_path_length_cap():
    """Beyond the cap (1025 tokens) PATH_AUTO runs the CUDA-core path and an explicit PATH_TCGEN05 request is refused."""
    import ctypes
    from rbn_b200 import _lib
    lib = _lib.load()
    d = _lib.RbnDesc(N=64, V=8, B=1, L=1026, n_tvars=1, path=_lib.PATH_AUTO, chunk_items=0, flags=0, cache_bytes=0)
    assert lib.rbn_resolve_path(ctypes.byref(d)) == _lib.PATH_SIMT
    d.path = _lib.PATH_TCGEN05
    assert lib.rbn_resolve_path(ctypes.byref(d)) == -4
    d.L = 1025
    assert lib.rbn_resolve_path(ctypes.byref(d)) == _lib.PATH_TCGEN05


@pytest.mark.parametrize("N,L", [(10, 64), (64, 64), (10, 120)])
def test_linear_space_gradients_of_long_sequences(N, L):
    """SequentialRBN.inside() returns Z in LINEAR space (base.py:120-121) and the reference differentiates that: dZ/dtheta
    for Z ~ 1e-80 .. 1e-150.  The device pass runs in the log-Z normalisation and the host rescales in fp64 (ADVICE r1:
    g * Z used to be cast to fp32 and the gradients went to exactly 0 beyond ~30 tokens)."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ
    V, B = 8, 2
    W, T, pr = O.synthetic_grammar(N, V, seed=21)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=22, ragged=True)
    g = np.array([0.7, -1.3])
    Wr, Tr, pr_r = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (W, T, pr))
    Zr = O.flat_inside_linear(Wr, Tr, pr_r, torch.tensor(tok)[:, :, None], lengths)
    (Zr * torch.tensor(g)).sum().backward()
    Wt, Tt, pt = (torch.tensor(x, dtype=torch.float64, device="cuda", requires_grad=True) for x in (W, T, pr))
    Z = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), linear=True)
    (Z * torch.tensor(g, device="cuda")).sum().backward()
    assert float(Zr.detach().abs().max()) < 1e-38                # beyond fp32: the case that used to break
    # the contract is on log Z (1e-4 relative); Z itself then carries |log Z| * 1e-4 of relative error, and dZ/dtheta =
    # Z * dlogZ/dtheta carries that on top of the gradient tolerance
    np.testing.assert_allclose(np.log(Z.detach().cpu().numpy()), np.log(Zr.detach().numpy()), rtol=LOGZ_RTOL)
    z_rel = float(np.max(np.abs(Z.detach().cpu().numpy() / Zr.detach().numpy() - 1)))
    for a, b in ((Wt, Wr), (Tt, Tr), (pt, pr_r)):
        assert float(b.grad.abs().max()) > 0
        m = metrics(a.grad.cpu().numpy(), b.grad.numpy())
        assert m["max_rel"] <= GRAD_RTOL + z_rel and m["l2_rel"] <= GRAD_RTOL + z_rel, (m, z_rel)
