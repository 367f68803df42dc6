"""Synthetic workloads of SURVEY.md section 8(d): seeded random grammars, token batches and Gaussian observations.

Lives in the product package so that `bench.py`'s timed arms build their inputs without touching `oracle/` (which is
test infrastructure); `oracle/` re-exports these generators so that tests, the oracle and the bench see the same inputs.
"""
import numpy as np


def synthetic_grammar(N, V, seed=0):
    """Random dense grammar normalised the reference's way: W over (left,right) per parent (pcfg.py:299), T over
    terminals per parent (pcfg.py:339), structural [terminal, binary] over transitions (pcfg.py:373), prior
    over symbols; returned FLAT (S folded in) as float64 numpy: W[N,N,N], T[V,N], prior[N]."""
    rng = np.random.default_rng(seed)
    W = rng.random((N, N, N))
    T = rng.random((V, N))
    S = rng.random((2, N))
    pi = rng.random(N)
    W /= W.sum(axis=(0, 1), keepdims=True)
    T /= T.sum(axis=0, keepdims=True)
    S /= S.sum(axis=0, keepdims=True)
    pi /= pi.sum()
    return W * S[1][None, None, :], T * S[0][None, :], pi


def synthetic_tokens(B, L, V, seed=1, ragged=False):
    """tokens[B, L] int64 (padding -1) and lengths[B]; `ragged` draws lengths from [L/2, L]."""
    rng = np.random.default_rng(seed)
    tok = rng.integers(0, V, (B, L)).astype(np.int64)
    lengths = np.full((B,), L, dtype=np.int64)
    if ragged:
        lengths = rng.integers(max(L // 2, 1), L + 1, B).astype(np.int64)
        for b in range(B):
            tok[b, lengths[b]:] = -1
    return tok, lengths


def synthetic_gauss(D, B, L, seed=0):
    """Config C4: y = cumulative-sum random walk of N(0, I) steps; every covariance is A A^T + 0.1 I with
    A ~ U(-1, 1)^{DxD} (the recipe of /root/reference/tests/test_multivariate_normal.py:89-91)."""
    rng = np.random.default_rng(seed)

    def spd():
        a = rng.uniform(-1, 1, (D, D))
        return a @ a.T + 0.1 * np.eye(D)

    y = np.cumsum(rng.standard_normal((B, L, D)), axis=1)
    return dict(y=y, cov_term=spd(), cov_left=spd(), cov_right=spd(), prior_mean=rng.standard_normal(D),
                prior_cov=spd(), log_struct=np.log(np.array([0.4, 0.6])))
