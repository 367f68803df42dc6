"""CPU oracle for the rbnet inside pass (+ autograd-driven outside).

TEST INFRASTRUCTURE ONLY.  Nothing under rbn_b200/ imports this module; it is imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` legs, and there only as the
checker or as the timed CPU arm -- never as the product.

Parity status: PINNED.  tests/test_oracle.py checks every function below against
  (1) the golden vectors held by the reference's own tests (restated in tests/test_oracle.py from
      /root/reference/tests/test_pcfg.py:14-35,37-81 and /root/reference/tests/test_base.py:47-295), and
  (2) outputs of the UNMODIFIED reference run in the build container (tests/golden/*.npz written by
      oracle/make_golden.py, which imports /root/reference through oracle/reference_loader.py).

Three tiers:
  bricks_inside()  -- literal fp64 restatement of the reference loop nest, brick by brick, linear probability
                      space, pure Python/numpy (small cases only).
  flatten_bricks() -- restates how a multi-cell RBN collapses to ONE block-structured grammar (W, T, prior);
                      this is algebra on the reference's formulas, cross-checked against bricks_inside().
  flat_inside()    -- vectorised, per-cell max-shifted fp64 torch restatement of the flat grammar, batched,
                      returning log Z; torch.autograd through it is the outside pass exactly as in the
                      reference.
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------------------
# plain-data description of an RBN ("bricks"), mirroring the reference constructors
#   cell       = dict(cardinality=K, struct=[n_trans, K] probabilities, transitions=[...])
#   transition = ("term", T[V, K], term_idx)  |  ("bin", W[Kl, Kr, K], left_idx, right_idx)
#   prior      = dict(struct=[n_cells], dists=[ [K_v] ... ])
# All arrays hold NORMALISED probabilities (what `.p` returns in the reference).
# --------------------------------------------------------------------------------------------------------------


def tmap_index(n, start, end):
    """Linear index of span (start, end) in the reference's chart; layout pinned by
    /root/reference/tests/test_pcfg.py:14-19 and /root/reference/tests/test_util.py:18-22."""
    depth = n - (end - start)
    return depth * (depth + 1) // 2 + start


def bricks_inside(cells, prior, sequence):
    """Literal restatement of RBN.inside (/root/reference/rbnet/base.py:93-121) for discrete bricks.

    sequence: list over positions of lists over terminal variables (int or None), as in
    /root/reference/tests/test_pcfg.py:59-77.  Returns (Z, [chart_v of shape (n(n+1)/2, K_v)]) in float64.
    """
    n = len(sequence)
    # SequentialRBN.init_inside (/root/reference/rbnet/sequential.py:18-21); chart = zeros (pcfg.py:215) but fp64
    charts = [np.zeros((n * (n + 1) // 2, c["cardinality"])) for c in cells]
    # inside_schedule: width-major, start-minor (/root/reference/rbnet/sequential.py:23-26)
    for span in range(1, n + 1):
        for start in range(n - span + 1):
            end = start + span
            new_values = []
            for cell in cells:
                struct = np.asarray(cell["struct"], dtype=np.float64)
                mixtures = []
                valid = []
                for t_idx, tr in enumerate(cell["transitions"]):
                    marg = []
                    if tr[0] == "term":
                        # DiscreteTerminalTransition.inside_marginals (/root/reference/rbnet/pcfg.py:344-352);
                        # splits exist only for width 1 (/root/reference/rbnet/sequential.py:65-71)
                        if span == 1:
                            _, T, term_idx = tr
                            y = sequence[start][term_idx]
                            if y is not None:
                                marg = [np.asarray(T, dtype=np.float64)[int(y), :]]
                    else:
                        # DiscreteBinaryNonTerminalTransition.inside_marginals (/root/reference/rbnet/pcfg.py:305-314);
                        # splits start+1..end-1 (/root/reference/rbnet/sequential.py:53-60)
                        _, W, li, ri = tr
                        W = np.asarray(W, dtype=np.float64)
                        for split in range(start + 1, end):
                            left = charts[li][tmap_index(n, start, split)]
                            right = charts[ri][tmap_index(n, split, end)]
                            marg.append((W * left[:, None, None] * right[None, :, None]).sum(axis=(0, 1)))
                    if len(marg) > 0:
                        # DiscreteNonTermVar.mixture over splits (/root/reference/rbnet/pcfg.py:219-238)
                        mixtures.append(np.stack(marg).sum(axis=0))
                        valid.append(t_idx)
                # DiscreteCell.inside_mixture (/root/reference/rbnet/pcfg.py:387-398)
                if len(mixtures) == 0:
                    new_values.append(np.zeros(cell["cardinality"]))
                else:
                    new_values.append((struct[valid, :] * np.stack(mixtures)).sum(axis=0))
            # all cells of one location are computed from the chart state BEFORE any of them is written only as
            # far as width-1/-w dependencies go; the reference writes cell by cell (base.py:104-118) which is
            # equivalent because a span never depends on same-width spans.
            for v, vals in enumerate(new_values):
                charts[v][tmap_index(n, start, end)] = vals
    # DiscretePrior.marginal_likelihood (/root/reference/rbnet/pcfg.py:270-273)
    root = tmap_index(n, 0, n)
    Z = float(sum(w * (np.asarray(p, dtype=np.float64) * c[root]).sum()
                  for w, p, c in zip(prior["struct"], prior["dists"], charts)))
    return Z, charts


def flatten_bricks(cells, prior):
    """Collapse a multi-cell RBN into one block-structured grammar.

    With offsets o_v = sum_{u<v} K_u and N = sum K_v:
      W[o_l + b, o_r + c, o_v + a] += S_v[tau, a] * W_tau[b, c, a]      (binary tau of cell v; pcfg.py:312, 395-398)
      T[q_t + y, o_v + a]          += S_v[tau, a] * T_tau[y, a]          (terminal tau with term_idx t; pcfg.py:351)
      prior[o_v + a]                = omega_v * pi_v[a]                  (pcfg.py:270-273)
    where q_t offsets the values of terminal variable t in a flat terminal alphabet.  Returns
    (W[N,N,N], T[Vtot,N], prior[N], var_offsets, term_offsets) as float64 torch tensors; differentiable when
    the brick arrays are torch tensors that require grad.
    """
    cards = [int(c["cardinality"]) for c in cells]
    off = np.concatenate([[0], np.cumsum(cards)]).astype(int)
    N = int(off[-1])
    as_t = lambda x: x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dtype=np.float64))
    # sizes of terminal variables
    tsize = {}
    for c in cells:
        for tr in c["transitions"]:
            if tr[0] == "term":
                tsize[tr[2]] = max(tsize.get(tr[2], 0), int(as_t(tr[1]).shape[0]))
    n_tvars = (max(tsize) + 1) if tsize else 1
    tcards = [tsize.get(t, 0) for t in range(n_tvars)]
    toff = np.concatenate([[0], np.cumsum(tcards)]).astype(int)
    W = torch.zeros((N, N, N), dtype=torch.float64)
    T = torch.zeros((max(int(toff[-1]), 1), N), dtype=torch.float64)
    for v, c in enumerate(cells):
        S = as_t(c["struct"])
        for t_idx, tr in enumerate(c["transitions"]):
            if tr[0] == "term":
                Tt = as_t(tr[1])
                t = tr[2]
                sl = (slice(int(toff[t]), int(toff[t]) + Tt.shape[0]), slice(int(off[v]), int(off[v + 1])))
                T[sl] = T[sl] + S[t_idx][None, :] * Tt
            else:
                Wt = as_t(tr[1])
                li, ri = tr[2], tr[3]
                sl = (slice(int(off[li]), int(off[li + 1])), slice(int(off[ri]), int(off[ri + 1])),
                      slice(int(off[v]), int(off[v + 1])))
                W[sl] = W[sl] + S[t_idx][None, None, :] * Wt
    ps = as_t(prior["struct"])
    pr = torch.cat([ps[v] * as_t(p) for v, p in enumerate(prior["dists"])])
    return W, T, pr, off, toff


def flatten_sequence(sequence, toff):
    """tokens[L, n_tvars] int64 with -1 for None (pcfg.py:346-349), offset into the flat terminal alphabet."""
    n_t = len(toff) - 1
    out = -np.ones((len(sequence), n_t), dtype=np.int64)
    for i, pos in enumerate(sequence):
        pos = list(np.atleast_1d(np.asarray(pos, dtype=object)))
        for t in range(min(n_t, len(pos))):
            if pos[t] is not None:
                out[i, t] = int(pos[t]) + int(toff[t])
    return out


# --------------------------------------------------------------------------------------------------------------
# vectorised, scaled, batched restatement of the flat grammar
# --------------------------------------------------------------------------------------------------------------

def flat_emission(T, tokens):
    """beta[i,i+1,:] = sum_t T[tokens[.., i, t], :], -1 contributes nothing (pcfg.py:344-352)."""
    B, L, nt = tokens.shape
    E = torch.zeros((B, L, T.shape[1]), dtype=T.dtype)
    for t in range(nt):
        tok = tokens[:, :, t]
        valid = tok >= 0
        rows = T[tok.clamp(min=0)]
        E = E + rows * valid[..., None].to(T.dtype)
    return E


def flat_forward(W, T, prior, tokens, lengths=None):
    """Batched inside pass of a flat grammar in per-cell max-shifted form.

    W[b, c, a] (left, right, parent -- parent LAST, /root/reference/rbnet/pcfg.py:91-93,289-290), T[Vtot, N],
    prior[N]: float64 torch tensors (may require grad).  tokens: int64 [B, L, n_tvars] (-1 = None);
    lengths: [B] (default L).  Recurrence restated from pcfg.py:305-314, 387-398, 270-273 with the split-sum
    taken first:  beta[i,k,a] = sum_{b,c} W[b,c,a] * sum_j beta[i,j,b] beta[j,k,c].

    Storage: vals[w] is [B, L-w+1, N] with max 1 per cell, logs[w] is [B, L-w+1] (natural log of the scale).
    Every cell keeps a FINITE scale: an all-zero cell inherits the scale of its best split (its mantissa is 0),
    so derivatives with respect to zero-probability entries survive, as they do in the reference's unscaled
    autograd graph.  The shifts are constants (detached), so autograd through this function differentiates
    exactly what the reference differentiates.
    Returns (zroot[B], rootscale[B], vals, logs) with Z = zroot * exp(rootscale).
    """
    tokens = torch.as_tensor(tokens)
    if tokens.dim() == 2:
        tokens = tokens[:, :, None]
    B, L, _ = tokens.shape
    N = W.shape[0]
    if lengths is None:
        lengths = torch.full((B,), L, dtype=torch.int64)
    lengths = torch.as_tensor(lengths, dtype=torch.int64)
    W2 = W.reshape(N * N, N)

    def shift(x):
        mx = x.detach().max(dim=-1).values
        safe = torch.where(mx > 0, mx, torch.ones_like(mx))
        return x / safe[..., None], safe.log()

    vals = [None] * (L + 1)
    logs = [None] * (L + 1)
    vals[1], logs[1] = shift(flat_emission(T, tokens))
    for w in range(2, L + 1):
        n_items = L - w + 1
        # stack over left width u = 1..w-1
        ls = torch.stack([logs[u][:, :n_items] + logs[w - u][:, u:u + n_items] for u in range(1, w)], dim=-1)
        m = ls.max(dim=-1).values                                   # [B, items]
        c = torch.exp(ls - m[..., None])
        Y = None
        for u in range(1, w):
            l = vals[u][:, :n_items] * c[..., u - 1][..., None]
            r = vals[w - u][:, u:u + n_items]
            y = l[..., :, None] * r[..., None, :]
            Y = y if Y is None else Y + y
        out = Y.reshape(B, n_items, N * N) @ W2                     # [B, items, N]
        v, lg = shift(out)
        vals[w] = v
        logs[w] = lg + m
    # DiscretePrior.marginal_likelihood on the root (0, n_b) (/root/reference/rbnet/pcfg.py:270-273)
    zroot = torch.stack([(prior * vals[int(lengths[b])][b, 0]).sum() for b in range(B)])
    rootscale = torch.stack([logs[int(lengths[b])][b, 0] for b in range(B)])
    return zroot, rootscale, vals, logs


def flat_inside(W, T, prior, tokens, lengths=None, return_chart=False):
    """log Z [B] (=-inf for ungrammatical input) of the flat grammar; see flat_forward()."""
    zroot, rootscale, vals, logs = flat_forward(W, T, prior, tokens, lengths)
    logZ = torch.log(zroot) + rootscale
    if return_chart:
        return logZ, vals, logs
    return logZ


def flat_inside_linear(W, T, prior, tokens, lengths=None):
    """Z [B] in linear probability space (what RBN.inside returns, base.py:120-121); differentiable at Z == 0."""
    zroot, rootscale, _, _ = flat_forward(W, T, prior, tokens, lengths)
    return zroot * torch.exp(rootscale)


def flat_inside_with_grads(W, T, prior, tokens, lengths=None, grad_logZ=None):
    """log Z and d(sum_b grad_logZ[b] * log Z_b)/d{W, T, prior} by autograd (fp64).  Sequences with Z == 0 are
    excluded from the gradient (their log Z is -inf)."""
    W = torch.as_tensor(W, dtype=torch.float64).clone().requires_grad_(True)
    T = torch.as_tensor(T, dtype=torch.float64).clone().requires_grad_(True)
    prior = torch.as_tensor(prior, dtype=torch.float64).clone().requires_grad_(True)
    logZ = flat_inside(W, T, prior, tokens, lengths)
    ok = torch.isfinite(logZ.detach())
    g = torch.ones_like(logZ) if grad_logZ is None else torch.as_tensor(grad_logZ, dtype=torch.float64)
    if bool(ok.any()):
        (logZ[ok] * g[ok]).sum().backward()
    z = lambda t: t.grad if t.grad is not None else torch.zeros_like(t)
    return logZ.detach(), z(W).detach(), z(T).detach(), z(prior).detach()


def chart_to_tmap(vals, logs, b, n):
    """Linear-space float64 chart of sequence b in the reference's TMap layout (root first)."""
    N = vals[1].shape[-1]
    arr = np.zeros((n * (n + 1) // 2, N))
    for w in range(1, n + 1):
        for s in range(n - w + 1):
            arr[tmap_index(n, s, s + w)] = vals[w][b, s].detach().numpy() * math.exp(float(logs[w][b, s]))
    return arr


# --------------------------------------------------------------------------------------------------------------
# synthetic workloads -- shared by tests and bench so every leg sees the same inputs
# --------------------------------------------------------------------------------------------------------------

from rbn_b200.synth import synthetic_grammar, synthetic_tokens  # noqa: E402,F401  (one generator for tests, oracle and bench)


def flat_viterbi(W, T, prior, tokens, lengths=None):
    """Most probable derivation of every sequence under the flat grammar (max-semiring CYK, log space, float64 numpy).

    Not part of the reference (it has no MAP parse); checker for rbn_viterbi.  Returns (logP [B], trees): a tree is
    (a, i, k, left_tree, right_tree) for a binary node over span (i, k) or (a, i, i + 1) for a leaf; None when the
    sequence has no derivation.  The emission of a position is the SUM over its terminal variables, as in the inside
    pass (pcfg.py:344-352)."""
    W, T, prior = (np.asarray(x, dtype=np.float64) for x in (W, T, prior))
    tokens = np.asarray(tokens)
    if tokens.ndim == 2:
        tokens = tokens[:, :, None]
    B, L, nt = tokens.shape
    lengths = np.full(B, L) if lengths is None else np.asarray(lengths)
    N = W.shape[0]
    with np.errstate(divide="ignore"):
        lW, lp = np.log(W), np.log(prior)
    out_lp, out_tree = np.full(B, -np.inf), []
    for b in range(B):
        n = int(lengths[b])
        best, back = {}, {}
        for i in range(n):
            e = np.zeros(N)
            for t in range(nt):
                if tokens[b, i, t] >= 0:
                    e += T[tokens[b, i, t]]
            with np.errstate(divide="ignore"):
                best[i, i + 1] = np.log(e)
        for w in range(2, n + 1):
            for i in range(n - w + 1):
                k = i + w
                cand = np.full((w - 1, N), -np.inf)
                arg = np.zeros((w - 1, N), dtype=np.int64)
                for u in range(1, w):
                    s = lW + best[i, i + u][:, None, None] + best[i + u, k][None, :, None]      # [b, c, a]
                    flat = s.reshape(N * N, N)
                    arg[u - 1] = flat.argmax(0)
                    cand[u - 1] = flat.max(0)
                us = cand.argmax(0)
                best[i, k] = cand.max(0)
                back[i, k] = [(int(us[a]) + 1, int(arg[us[a], a]) // N, int(arg[us[a], a]) % N) for a in range(N)]
        root = best[0, n] + lp
        a0 = int(root.argmax())
        out_lp[b] = root[a0]

        def build(a, i, k):
            if k - i == 1:
                return (a, i, k)
            u, bl, cr = back[i, k][a]
            return (a, i, k, build(bl, i, i + u), build(cr, i + u, k))
        out_tree.append(build(a0, 0, n) if np.isfinite(root[a0]) else None)
    return out_lp, out_tree


def tree_log_prob(tree, W, T, prior, tokens_b):
    """log probability of one derivation tree (as returned by flat_viterbi / rbn_b200's viterbi) under the flat grammar."""
    W, T, prior = (np.asarray(x, dtype=np.float64) for x in (W, T, prior))
    tokens_b = np.asarray(tokens_b)
    if tokens_b.ndim == 1:
        tokens_b = tokens_b[:, None]

    def rec(t):
        if len(t) == 3:
            a, i, _ = t
            return np.log(sum(T[y, a] for y in tokens_b[i] if y >= 0))
        a, _, _, lt, rt = t
        return np.log(W[lt[0], rt[0], a]) + rec(lt) + rec(rt)
    with np.errstate(divide="ignore"):
        return float(np.log(prior[tree[0]]) + rec(tree))
