"""Edge cases of the chart pass on the GPU: length-1 sequences, ragged batches that contain them, positions without a
terminal (None / -1, /root/reference/rbnet/pcfg.py:346-349) which make a sequence ungrammatical (Z = 0, log Z = -inf,
no NaN in any gradient, tests/test_base.py:293 in the reference), and a batch of one."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _grads(N, V, tok, lengths, path, seed=0):
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ
    W, T, pr = O.synthetic_grammar(N, V, seed=seed)
    Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=path)
    finite = torch.isfinite(lz)
    torch.where(finite, lz, torch.zeros_like(lz)).sum().backward()
    return (W, T, pr), lz.detach().cpu().numpy(), [g.grad.cpu().numpy() for g in (Wt, Tt, pt)]


@pytest.mark.parametrize("N", [10, 64])
def test_length_one_sequences(N):
    """Only width-1 cells exist: Z = sum_a prior[a] T[y, a]; binary rules get no count."""
    from rbn_b200 import _lib
    V = 7
    tok = np.array([[3], [0], [6]], dtype=np.int32)
    (W, T, pr), lz, (gW, gT, gP) = _grads(N, V, tok, np.array([1, 1, 1]), _lib.PATH_AUTO)
    ref = np.log((T[tok[:, 0]] * pr[None, :]).sum(1))
    np.testing.assert_allclose(lz, ref, rtol=1e-5)
    assert np.abs(gW).max() == 0.0
    assert abs((gT * T).sum() - 3) < 1e-3 and abs((gP * pr).sum() - 3) < 1e-3


@pytest.mark.parametrize("N,path", [(10, "simt"), (64, "tc"), (128, "tc")])
def test_ragged_batch_with_length_one_and_two_members(N, path):
    from oracle import inside_oracle as O
    from rbn_b200 import _lib
    V, L = 6, 9
    tok, _ = O.synthetic_tokens(5, L, V, seed=3)
    lengths = np.array([1, 9, 2, 5, 1])
    for b in range(5):
        tok[b, lengths[b]:] = -1
    p = _lib.PATH_SIMT if path == "simt" else _lib.PATH_TCGEN05
    (W, T, pr), lz, (gW, gT, gP) = _grads(N, V, tok, lengths, p, seed=4)
    ref, rW, rT, rP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
    np.testing.assert_allclose(lz, ref.numpy(), rtol=1e-4)
    for g, r in ((gW, rW), (gT, rT), (gP, rP)):
        assert np.abs(g - r.numpy()).max() <= 1e-3 * np.abs(r.numpy()).max()


@pytest.mark.parametrize("N,path", [(10, "simt"), (64, "tc")])
def test_missing_terminal_makes_the_sequence_ungrammatical_without_nans(N, path):
    from oracle import inside_oracle as O
    from rbn_b200 import _lib
    V, L = 5, 8
    tok, lengths = O.synthetic_tokens(3, L, V, seed=5)
    tok[1, 4] = -1                                       # None inside sequence 1: its cell (4,5) has no mass
    p = _lib.PATH_SIMT if path == "simt" else _lib.PATH_TCGEN05
    (W, T, pr), lz, grads = _grads(N, V, tok, lengths, p, seed=6)
    assert np.isneginf(lz[1]) and np.isfinite(lz[[0, 2]]).all()
    for g in grads:
        assert np.isfinite(g).all()
    ref, rW, rT, rP = O.flat_inside_with_grads(W, T, pr, tok[[0, 2]][:, :, None], lengths[[0, 2]])
    np.testing.assert_allclose(lz[[0, 2]], ref.numpy(), rtol=1e-4)
    assert np.abs(grads[0] - rW.numpy()).max() <= 1e-3 * np.abs(rW.numpy()).max()      # the dead sequence adds nothing


def test_batch_of_one_on_the_tensor_core_path():
    from oracle import inside_oracle as O
    from rbn_b200 import _lib
    N, V, L = 64, 8, 17
    tok, lengths = O.synthetic_tokens(1, L, V, seed=8)
    (W, T, pr), lz, (gW, gT, gP) = _grads(N, V, tok, lengths, _lib.PATH_TCGEN05, seed=9)
    ref, rW, rT, rP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
    np.testing.assert_allclose(lz, ref.numpy(), rtol=1e-4)
    assert np.abs(gW - rW.numpy()).max() <= 1e-3 * np.abs(rW.numpy()).max()


def test_two_passes_on_two_streams_with_distinct_workspaces():
    """The C ABI is stream-ordered and keeps no state between calls (include/rbn_b200.h): two passes queued on two
    streams at once, each with its own workspace, give the results of running them one after the other."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    N, V, L, B = 64, 8, 24, 16
    W, T, pr = O.synthetic_grammar(N, V, seed=13)
    toks = [O.synthetic_tokens(B, L, V, seed=14 + k, ragged=True) for k in range(2)]
    base = []
    for tok, ln in toks:
        Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
        lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(ln), path=_lib.PATH_TCGEN05)
        lz.sum().backward()
        base.append((lz.detach().cpu().numpy(), Wt.grad.cpu().numpy()))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for (tok, ln), st in zip(toks, streams):
        with torch.cuda.stream(st):
            Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
            lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok).cuda(), torch.tensor(ln).cuda(), path=_lib.PATH_TCGEN05)
            lz.sum().backward()
            outs.append((lz, Wt))
    torch.cuda.synchronize()
    for (lz, Wt), (lz0, g0) in zip(outs, base):
        np.testing.assert_allclose(lz.detach().cpu().numpy(), lz0, rtol=1e-6)
        np.testing.assert_allclose(Wt.grad.cpu().numpy(), g0, rtol=1e-4, atol=1e-7 * np.abs(g0).max())


def test_nonterminal_count_is_padded_onto_the_tensor_core_path():
    """N = 100 with a large enough batch: AUTO zero-pads the grammar to N = 128 (tcgen05 kernels); results and gradient
    shapes are those of the unpadded grammar."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, engine
    N, V, L, B = 100, 9, 13, 64
    assert engine._padded_cardinality(N, B * L * (L - 1) // 2) == 128
    assert engine._padded_cardinality(10, 10 ** 6) == 10 and engine._padded_cardinality(N, 100) == N
    W, T, pr = O.synthetic_grammar(N, V, seed=40)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=41, ragged=True)
    Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
    holder = []
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), holder=holder)
    assert holder[0].desc.N == 128
    lz.sum().backward()
    assert Wt.grad.shape == (N, N, N) and Tt.grad.shape == (V, N) and pt.grad.shape == (N,)
    ref, rW, rT, rP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
    np.testing.assert_allclose(lz.detach().cpu().numpy(), ref.numpy(), rtol=1e-4)
    for g, r in ((Wt.grad, rW), (Tt.grad, rT), (pt.grad, rP)):
        assert float((g.cpu() - r).abs().max() / r.abs().max()) <= 1e-3


@pytest.mark.parametrize("N,path", [(10, "simt"), (64, "tc")])
def test_ragged_batch_is_sorted_internally_and_results_come_back_in_caller_order(N, path):
    """The engine sorts a ragged batch by length and tells the library how many sequences reach each width
    (rbn_inside_forward_sorted); log Z, gradients and exported charts are those of the caller's order."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    V, L, B = 6, 21, 40                       # SIMT: B (L - 1) large enough to leave the single-launch path
    tok, _ = O.synthetic_tokens(B, L, V, seed=50)
    lengths = np.random.default_rng(51).integers(1, L + 1, B)
    lengths[3], lengths[7] = L, 1
    for b in range(B):
        tok[b, lengths[b]:] = -1
    W, T, pr = O.synthetic_grammar(N, V, seed=52)
    coef = np.random.default_rng(53).uniform(0.5, 1.5, B)
    p = _lib.PATH_SIMT if path == "simt" else _lib.PATH_TCGEN05
    Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
    holder = []
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=p, holder=holder)
    assert holder[0].active is not None and list(holder[0].active)[:2] == [B, B] and holder[0].active[L] >= 1
    (lz * torch.tensor(coef, device="cuda")).sum().backward()
    ref, rW, rT, rP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths, grad_logZ=coef)
    np.testing.assert_allclose(lz.detach().cpu().numpy(), ref.numpy(), rtol=1e-4)
    for g, r in ((Wt.grad, rW), (Tt.grad, rT), (pt.grad, rP)):
        assert float((g.cpu() - r).abs().max() / r.abs().max()) <= 1e-3
    # chart of caller-order sequence 5 == chart of that sequence parsed alone
    alone = []
    inside_logZ(Wt.detach(), Tt.detach(), pt.detach(), torch.tensor(tok[5:6, :lengths[5]]), torch.tensor(lengths[5:6]), path=p, holder=alone)
    np.testing.assert_allclose(holder[0].export_chart(5).cpu().numpy(), alone[0].export_chart(0).cpu().numpy(), rtol=1e-5, atol=1e-300)


def test_batched_tokens_with_two_terminal_variables():
    """`log_inside_batch(tokens=...)` takes per-variable terminal values (they used to be read as flat-alphabet
    offsets, so variable 1 silently indexed variable 0's rows of T).  Same log Z and gradients as the `sequences=` path,
    which packs through pack_sequences; an out-of-range value raises IndexError before anything is launched."""
    from rbn_b200 import pcfg, sequential
    rng = np.random.default_rng(60)
    K, V0, V1 = 5, 3, 4

    def model():
        r = np.random.default_rng(61)
        cell = pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(K), weights=r.random((3, K)),
                                 transitions=[pcfg.DiscreteTerminalTransition(weights=r.random((V0, K)), term_idx=0),
                                              pcfg.DiscreteTerminalTransition(weights=r.random((V1, K)), term_idx=1),
                                              pcfg.DiscreteBinaryNonTerminalTransition(weights=r.random((K, K, K)))])
        prior = pcfg.DiscretePrior(struc_weights=np.ones(1), prior_weights=[r.random(K)])
        return sequential.SequentialRBN(cells=[cell], prior=prior).to("cuda")

    B, L = 4, 7
    lengths = np.array([7, 5, 7, 2])
    raw = -np.ones((B, L, 2), dtype=np.int64)
    seqs = []
    for b in range(B):
        seq = []
        for i in range(lengths[b]):
            y0, y1 = int(rng.integers(0, V0)), (int(rng.integers(0, V1)) if rng.random() > 0.3 else None)
            raw[b, i] = [y0, -1 if y1 is None else y1]
            seq.append([y0, y1])
        seqs.append(seq)
    m1, m2 = model(), model()
    lz1 = m1.log_inside_batch(sequences=seqs)
    lz2 = m2.log_inside_batch(tokens=torch.tensor(raw), lengths=torch.tensor(lengths))
    np.testing.assert_allclose(lz2.detach().cpu().numpy(), lz1.detach().cpu().numpy(), rtol=1e-6)
    lz1.sum().backward()
    lz2.sum().backward()
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2
        # same kernels, different batch order (pack_sequences sorts): fp32 atomics add in another order
        np.testing.assert_allclose(p2.grad.cpu().numpy(), p1.grad.cpu().numpy(), rtol=1e-4, atol=1e-6)
    bad = raw.copy()
    bad[0, 0, 0] = V0                              # valid for variable 1's alphabet, not for variable 0's
    with pytest.raises(IndexError):
        m2.log_inside_batch(tokens=torch.tensor(bad), lengths=torch.tensor(lengths))
