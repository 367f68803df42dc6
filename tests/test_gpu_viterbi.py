"""Most probable derivation (rbn_viterbi, max-semiring variant of the inside kernels) against a numpy CYK and against
brute-force enumeration of all derivations on a tiny case."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _all_trees(N, i, k):
    """Every derivation shape + labelling over span (i, k) (tiny cases only)."""
    if k - i == 1:
        for a in range(N):
            yield (a, i, k)
        return
    for j in range(i + 1, k):
        for lt in _all_trees(N, i, j):
            for rt in _all_trees(N, j, k):
                for a in range(N):
                    yield (a, i, k, lt, rt)


def test_viterbi_is_the_maximum_over_all_derivations():
    from oracle import inside_oracle as O
    from rbn_b200.engine import viterbi
    N, V, L = 2, 3, 4
    W, T, pr = O.synthetic_grammar(N, V, seed=60)
    tok = np.array([[0, 2, 1, 1]])
    brute = max(O.tree_log_prob(t, W, T, pr, tok[0]) for t in _all_trees(N, 0, L))
    logP, trees = viterbi(torch.tensor(W), torch.tensor(T), torch.tensor(pr), torch.tensor(tok))
    assert abs(float(logP[0]) - brute) < 1e-5 * abs(brute)
    assert abs(O.tree_log_prob(trees[0], W, T, pr, tok[0]) - brute) < 1e-9 * abs(brute)


@pytest.mark.parametrize("N,V,L,B,ragged", [(7, 5, 9, 4, True), (10, 19, 20, 3, False), (40, 6, 12, 2, True)])
def test_viterbi_vs_numpy_cyk(N, V, L, B, ragged):
    from oracle import inside_oracle as O
    from rbn_b200.engine import viterbi
    W, T, pr = O.synthetic_grammar(N, V, seed=61 + N)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=62 + N, ragged=ragged)
    ref_lp, ref_trees = O.flat_viterbi(W, T, pr, tok, lengths)
    logP, trees = viterbi(torch.tensor(W), torch.tensor(T), torch.tensor(pr), torch.tensor(tok), torch.tensor(lengths))
    np.testing.assert_allclose(logP.cpu().numpy(), ref_lp, rtol=1e-5)
    for b in range(B):                      # ties may pick another tree: compare the probability of the tree returned
        lp_tree = O.tree_log_prob(trees[b], W, T, pr, tok[b, :lengths[b]])
        assert abs(lp_tree - ref_lp[b]) <= 1e-5 * abs(ref_lp[b])
        assert trees[b][1] == 0 and trees[b][2] == lengths[b]
    # the best derivation can never be more probable than the whole marginal
    lz = O.flat_inside(torch.tensor(W), torch.tensor(T), torch.tensor(pr), torch.tensor(tok)[:, :, None], lengths)
    assert (logP.cpu().numpy() <= lz.numpy() + 1e-6).all()


def test_viterbi_through_the_model_api_and_ungrammatical_input():
    from rbn_b200 import pcfg, util
    g = pcfg.AbstractedPCFG(non_terminals="SAB", terminals="ab", start="S", rules=[
        ("S --> A B", 1), ("S --> B A", 1), ("A --> B A", 1), ("B --> A B", 1), ("A --> a", 1), ("B --> b", 1)],
        prob_rep=util.Prob)
    logP, trees = g.viterbi(["aaab", "aaaa"])
    assert abs(float(torch.exp(logP[0])) - 2.0 ** -7) < 1e-9       # the grammar is unambiguous for aaab: best = Z (tests/test_pcfg.py:33)
    assert trees[0][0] == (0, 0) and trees[0][1:3] == (0, 4)       # root S over the whole sequence
    assert trees[1] is None and float(logP[1]) == -np.inf
