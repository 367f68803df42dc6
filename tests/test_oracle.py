"""The oracle against the reference's golden vectors and against outputs of the unmodified reference."""
import math

import numpy as np
import pytest
import torch

from oracle import inside_oracle as O

# /root/reference/tests/test_pcfg.py:14-19 -- the reference's own golden chart for "aaab" (columns S, A, B)
BINARY_GRAMMAR_CHART = np.array([
    [1 / 2 ** 7, 0, 1 / 2 ** 7],
    [0, 0, 0], [1 / 2 ** 5, 0, 1 / 2 ** 5],
    [0, 0, 0], [0, 0, 0], [1 / 2 ** 3, 0, 1 / 2 ** 3],
    [0, 1 / 2, 0], [0, 1 / 2, 0], [0, 1 / 2, 0], [0, 0, 1 / 2],
])


def binary_grammar_bricks():
    """AbstractedPCFG("SAB","ab", S->AB|BA, A->BA|a, B->AB|b) as normalised by the reference constructor
    (/root/reference/rbnet/pcfg.py:97-107): struct rows [terminal, binary]."""
    W = np.zeros((3, 3, 3))
    S, A, B = 0, 1, 2
    W[A, B, S] = W[B, A, S] = 0.5
    W[B, A, A] = 1.0
    W[A, B, B] = 1.0
    T = np.zeros((2, 3))
    T[:, S] = 0.5            # zero column made uniform (pcfg.py:99); its structural weight is 0
    T[0, A] = 1.0
    T[1, B] = 1.0
    struct = np.array([[0.0, 0.5, 0.5], [1.0, 0.5, 0.5]])
    cells = [dict(cardinality=3, struct=struct, transitions=[("term", T, 0), ("bin", W, 0, 0)])]
    prior = dict(struct=np.ones(1), dists=[np.array([1.0, 0, 0])])
    return cells, prior


def test_reference_golden_chart_bricks():
    cells, prior = binary_grammar_bricks()
    tok = {"a": 0, "b": 1}
    for s, z in [("aaaa", 0.0), ("bbbb", 0.0), ("bbba", 2.0 ** -7), ("aaab", 2.0 ** -7)]:
        Z, charts = O.bricks_inside(cells, prior, [[tok[c]] for c in s])
        assert Z == z
    np.testing.assert_array_equal(charts[0], BINARY_GRAMMAR_CHART)


def test_reference_golden_chart_flat():
    cells, prior = binary_grammar_bricks()
    W, T, pr, off, toff = O.flatten_bricks(cells, prior)
    tok = {"a": 0, "b": 1}
    seqs = ["aaaa", "bbbb", "bbba", "aaab"]
    tokens = torch.tensor([[[tok[c]] for c in s] for s in seqs])
    logZ, vals, logs = O.flat_inside(W, T, pr, tokens, return_chart=True)
    np.testing.assert_allclose(torch.exp(logZ).numpy(), [0, 0, 2.0 ** -7, 2.0 ** -7], rtol=1e-14)
    np.testing.assert_allclose(O.chart_to_tmap(vals, logs, 3, 4), BINARY_GRAMMAR_CHART, rtol=1e-14)


def test_expanded_pcfg_two_variables():
    """/root/reference/tests/test_pcfg.py:37-81: same language as two cardinality-1 variables, Z = 2^-8."""
    one = np.ones((1, 1, 1))
    cells = [dict(cardinality=1, struct=np.array([[0.5], [0.5]]), transitions=[("bin", one, 1, 0), ("term", np.ones((1, 1)), 0)]),
             dict(cardinality=1, struct=np.array([[0.5], [0.5]]), transitions=[("bin", one, 0, 1), ("term", np.ones((1, 1)), 1)])]
    prior = dict(struct=np.array([0.5, 0.5]), dists=[np.ones(1), np.ones(1)])
    a, b = [0, None], [None, 0]
    for seq, z in [([a] * 4, 0.0), ([b] * 4, 0.0), ([b, b, b, a], 2.0 ** -8), ([a, a, a, b], 2.0 ** -8)]:
        Z, charts = O.bricks_inside(cells, prior, seq)
        assert Z == z
        W, T, pr, off, toff = O.flatten_bricks(cells, prior)
        logZ = O.flat_inside(W, T, pr, torch.as_tensor(O.flatten_sequence(seq, toff))[None])
        assert math.isclose(float(torch.exp(logZ[0])), z, rel_tol=1e-14)
    np.testing.assert_array_equal(charts[0][:, 0], BINARY_GRAMMAR_CHART[:, 1])
    np.testing.assert_array_equal(charts[1][:, 0], BINARY_GRAMMAR_CHART[:, 2])


def test_minimal_grammar_closed_form():
    """/root/reference/tests/test_base.py:47-77: N=2 uniform grammar, closed-form per-level inside values."""
    cells = [dict(cardinality=2, struct=np.full((2, 2), 0.5),
                  transitions=[("term", np.full((2, 2), 0.5), 0), ("bin", np.full((2, 2, 2), 0.25), 0, 0)])]
    prior = dict(struct=np.ones(1), dists=[np.full(2, 0.5)])
    n = 5
    seq = [[int(x)] for x in np.random.default_rng(0).integers(0, 2, n)]
    Z, charts = O.bricks_inside(cells, prior, seq)
    lvl = np.zeros(n)
    for k in range(n):
        lvl[k] = 0.25 if k == 0 else (lvl[:k] * np.flip(lvl[:k])).sum() * 0.5
        for s in range(n - k):
            np.testing.assert_array_equal(charts[0][O.tmap_index(n, s, s + k + 1)], np.ones(2) * lvl[k])
    assert math.isclose(Z, lvl[-1], rel_tol=1e-15)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_bricks_vs_unmodified_reference(golden):
    for key, c in golden.items():
        Z, charts = O.bricks_inside(c["cells"], c["prior"], c["sequence"])
        tol = 1e-6 if c["meta"]["extra"].get("shipped_dtype") else 1e-12
        assert abs(Z - c["Z"]) <= tol * max(abs(c["Z"]), 1e-300), key
        for ch, ref in zip(charts, c["charts"]):
            assert _rel(ch, ref) <= tol, key


def test_flat_vs_unmodified_reference(golden):
    for key, c in golden.items():
        W, T, pr, off, toff = O.flatten_bricks(c["cells"], c["prior"])
        tokens = torch.as_tensor(O.flatten_sequence(c["sequence"], toff))[None]
        logZ, vals, logs = O.flat_inside(W, T, pr, tokens, return_chart=True)
        tol = 1e-6 if c["meta"]["extra"].get("shipped_dtype") else 1e-12
        if c["Z"] == 0:
            assert float(logZ[0]) == -math.inf, key
        else:
            assert abs(float(logZ[0]) - math.log(c["Z"])) <= tol * abs(math.log(c["Z"])) + 1e-13, key
        n = len(c["sequence"])
        flat_ref = np.concatenate(c["charts"], axis=1)
        assert _rel(O.chart_to_tmap(vals, logs, 0, n), flat_ref) <= tol, key


def _project(g, dims):
    """the reference's project_grad hook (/root/reference/rbnet/util.py:147-156,200-202)."""
    return g - g.sum(dim=dims, keepdim=True) / np.prod([g.shape[d] for d in dims])


def test_flat_autograd_vs_reference_gradients(golden):
    """d Z / d parameter through flatten_bricks + flat_inside equals what Z.backward() leaves on the reference's
    Parameters (after the projection hook), for both Prob and LogProb representations."""
    n_checked = 0
    for key, c in golden.items():
        if not c["grads"]:
            continue
        log_rep = any(k.endswith("log_p") for k in c["grads"])
        leaves, cells = {}, []

        def leaf(name, p, dims):
            p = torch.as_tensor(p, dtype=torch.float64)
            x = (torch.log(p) if log_rep else p.clone()).requires_grad_(True)
            leaves[name] = (x, dims)
            return torch.exp(x) if log_rep else x

        suffix = "log_p" if log_rep else "p"
        for v, cell in enumerate(c["cells"]):
            trs = []
            for t, tr in enumerate(cell["transitions"]):
                dims = (0,) if tr[0] == "term" else (0, 1)
                trs.append((tr[0], leaf(f"_cells.{v}._transitions.{t}.transition_probabilities.{suffix}", tr[1], dims)) + tuple(tr[2:]))
            cells.append(dict(cardinality=cell["cardinality"], transitions=trs,
                              struct=leaf(f"_cells.{v}.transition_probabilities.{suffix}", cell["struct"], (0,))))
        prior = dict(struct=leaf(f"_prior.structural_distribution.{suffix}", c["prior"]["struct"], (0,)),
                     dists=[leaf(f"_prior.prior_distributions.{v}.{suffix}", d, (0,)) for v, d in enumerate(c["prior"]["dists"])])
        W, T, pr, off, toff = O.flatten_bricks(cells, prior)
        tokens = torch.as_tensor(O.flatten_sequence(c["sequence"], toff))[None]
        Z = O.flat_inside_linear(W, T, pr, tokens)[0]
        Z.backward()
        assert set(leaves) == set(c["grads"]), key
        for name, (x, dims) in leaves.items():
            g = torch.nan_to_num(x.grad) if x.grad is not None else torch.zeros_like(x)
            ref = c["grads"][name]
            assert _rel(_project(g, dims).numpy(), ref) <= 1e-10, (key, name)
        n_checked += 1
    assert n_checked >= 20


def test_expected_count_identities():
    """sum of d logZ / d log p over binary rules = L-1, terminal rules = L, prior = 1."""
    N, V, L, B = 6, 5, 9, 3
    W, T, pr = O.synthetic_grammar(N, V, seed=3)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=4, ragged=True)
    logZ, gW, gT, gP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
    assert math.isclose(float((gW * W).sum()), float((lengths - 1).sum()), rel_tol=1e-10)
    assert math.isclose(float((gT * T).sum()), float(lengths.sum()), rel_tol=1e-10)
    assert math.isclose(float((gP * pr).sum()), B, rel_tol=1e-10)


def test_flat_equals_bricks_longer_sequences():
    """scaled/vectorised vs literal loops where linear space is still representable (L <= 24)."""
    rng = np.random.default_rng(7)
    for N, V, L in [(4, 3, 16), (7, 5, 24)]:
        W, T, pr = O.synthetic_grammar(N, V, seed=int(rng.integers(1000)))
        tok, _ = O.synthetic_tokens(1, L, V, seed=int(rng.integers(1000)))
        cells = [dict(cardinality=N, struct=np.ones((2, N)), transitions=[("term", T, 0), ("bin", W, 0, 0)])]
        prior = dict(struct=np.ones(1), dists=[pr])
        Z, charts = O.bricks_inside(cells, prior, [[int(t)] for t in tok[0]])
        logZ, vals, logs = O.flat_inside(torch.as_tensor(W), torch.as_tensor(T), torch.as_tensor(pr),
                                         torch.as_tensor(tok)[:, :, None], return_chart=True)
        assert math.isclose(float(logZ[0]), math.log(Z), rel_tol=1e-12)
        assert _rel(O.chart_to_tmap(vals, logs, 0, L), charts[0]) <= 1e-11


def test_viterbi_oracle_equals_brute_force_enumeration():
    """The numpy max-semiring CYK (checker of rbn_viterbi) against enumeration of every derivation on a tiny grammar."""
    import numpy as np
    from oracle import inside_oracle as O

    def all_trees(N, i, k):
        if k - i == 1:
            for a in range(N):
                yield (a, i, k)
            return
        for j in range(i + 1, k):
            for lt in all_trees(N, i, j):
                for rt in all_trees(N, j, k):
                    for a in range(N):
                        yield (a, i, k, lt, rt)

    N, V, L = 2, 3, 4
    W, T, pr = O.synthetic_grammar(N, V, seed=9)
    W[0, 1, 0] = 0.0                                    # a zero-probability rule
    tok = np.array([[2, 0, 1, 1], [1, 1, 0, -1]])
    lengths = np.array([4, 3])
    lp, trees = O.flat_viterbi(W, T, pr, tok, lengths)
    for b in range(2):
        n = lengths[b]
        brute = max(O.tree_log_prob(t, W, T, pr, tok[b, :n]) for t in all_trees(N, 0, n))
        assert abs(lp[b] - brute) < 1e-12
        assert abs(O.tree_log_prob(trees[b], W, T, pr, tok[b, :n]) - brute) < 1e-12
