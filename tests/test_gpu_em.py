"""EM step on the GPU path: E-step counts against the fp64 oracle's gradient, count identities, monotone likelihood."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(N, V, seed):
    from rbn_b200 import pcfg, sequential
    rng = np.random.default_rng(seed)
    cell = pcfg.DiscreteCell(variable=pcfg.DiscreteNonTermVar(N), weights=rng.random((2, N)),
                             transitions=[pcfg.DiscreteTerminalTransition(weights=rng.random((V, N))),
                                          pcfg.DiscreteBinaryNonTerminalTransition(weights=rng.random((N, N, N)))])
    prior = pcfg.DiscretePrior(struc_weights=np.ones(1), prior_weights=[rng.random(N)])
    return sequential.SequentialRBN(cells=[cell], prior=prior).to("cuda")


def test_e_step_counts_match_oracle_and_identities():
    from oracle import inside_oracle as O
    from rbn_b200 import em
    N, V, L, B = 6, 5, 9, 7
    model = _model(N, V, seed=11)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=12, ragged=True)
    total, n_ok, counts = em.expected_counts(model, torch.tensor(tok), torch.tensor(lengths), micro_batch=3)
    assert int(n_ok) == B
    W, T, pr = (x.detach().cpu().double().numpy() for x in model._get_engine().flatten())
    lz, gW, gT, gP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
    assert abs(float(total) - float(lz.sum())) <= 1e-4 * abs(float(lz.sum()))
    by_shape = {tuple(c.shape): c.cpu().numpy() for c in counts}
    cW, cT, cS, cP = by_shape[(N, N, N)], by_shape[(V, N)], by_shape[(2, N)], by_shape[(N,)]
    np.testing.assert_allclose(cW, gW.numpy() * W, rtol=2e-3, atol=1e-6)          # count = dlogZ/dW * W
    np.testing.assert_allclose(cT, gT.numpy() * T, rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(cP, gP.numpy() * pr, rtol=2e-3, atol=1e-6)
    # binary rules L-1, terminal rules L, structural rows = per-parent marginals, prior 1 per sequence
    assert abs(cW.sum() - (lengths - 1).sum()) < 1e-3 * lengths.sum()
    assert abs(cT.sum() - lengths.sum()) < 1e-3 * lengths.sum()
    np.testing.assert_allclose(cS[0], cT.sum(0), rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(cS[1], cW.sum((0, 1)), rtol=2e-3, atol=1e-6)
    assert abs(cP.sum() - B) < 1e-3 * B


def test_em_iterations_do_not_decrease_the_likelihood():
    from oracle import inside_oracle as O
    from rbn_b200 import em
    N, V, L, B = 8, 6, 12, 24
    model = _model(N, V, seed=21)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=22, ragged=True)
    tok, lengths = torch.tensor(tok), torch.tensor(lengths)
    history = [em.em_step(model, tok, lengths, micro_batch=16) for _ in range(6)]
    final = float(model.log_inside_batch(tokens=tok, lengths=lengths).detach().mean())
    history.append(final)
    for a, b in zip(history[:-1], history[1:]):
        assert b >= a - 1e-4 * abs(a), history
    assert history[-1] > history[0] + 0.05
    for m, p in em.constrained_parameters(model):                  # parameters stayed normalised
        assert not torch.isnan(p).any()
        norm = torch.logsumexp(p, dim=m._dims()) if hasattr(m, "log_p") else torch.log(p.sum(dim=m._dims()))
        assert float(norm.abs().max()) < 1e-9


def test_em_on_the_tensor_core_path_and_fused_lognormalize():
    """N = 64 goes through the tcgen05 kernels; a float32 LogProb tensor is normalised by rbn_lognormalize."""
    from oracle import inside_oracle as O
    from rbn_b200 import em
    N, V, L, B = 64, 8, 10, 6
    model = _model(N, V, seed=31)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=32, ragged=False)
    l0 = em.em_step(model, torch.tensor(tok), torch.tensor(lengths))
    l1 = em.em_step(model, torch.tensor(tok), torch.tensor(lengths))
    assert l1 >= l0 - 1e-4 * abs(l0)
    x = torch.randn(5, 7, 11, device="cuda", dtype=torch.float32)
    ref = x.double() - torch.logsumexp(x.double(), dim=(0, 1), keepdim=True)
    out = em._lognormalize_(x.clone(), (0, 1))
    np.testing.assert_allclose(out.cpu().numpy(), ref.float().cpu().numpy(), rtol=1e-5, atol=1e-5)
    # the API's parameters are float64 (built from numpy weights): same kernel, double arithmetic (rbn_lognormalize_f64);
    # the M-step of the model above went through it (its LogProb tensors are fp64 on the GPU)
    x64 = torch.randn(5, 7, 11, device="cuda", dtype=torch.float64)
    x64[:, :, 3] = -float("inf")
    x64[2, 4, 3] = 0.5                                     # a group with a single finite entry
    ref64 = x64 - torch.logsumexp(x64, dim=(0, 1), keepdim=True)
    out64 = em._lognormalize_(x64.clone(), (0, 1))
    np.testing.assert_allclose(out64.cpu().numpy(), ref64.cpu().numpy(), rtol=1e-13, atol=1e-13)
    assert all(p.dtype == torch.float64 and p.is_cuda for _, p in em.constrained_parameters(model))


def test_length_bucketed_batches_feed_the_tensor_core_em_step():
    """SURVEY 8(f).2 end to end: a ragged corpus -> rbn_b200.data.LengthBucketedBatches (tokenise once, bucket by length,
    pinned int32 batches) -> em.expected_counts on the tcgen05 path (N = 64), batch by batch -> the summed expected
    counts equal the fp64 oracle's over the whole corpus.  The batches stay host tensors: the library copies them to the
    device itself and a pass never synchronises."""
    from oracle import inside_oracle as O
    from rbn_b200 import data, em, engine
    N, V = 64, 8
    model = _model(N, V, seed=41)
    rng = np.random.default_rng(42)
    corpus = [rng.integers(0, V, int(n)).astype(np.int32) for n in rng.integers(3, 15, 37)]
    loader = data.LengthBucketedBatches(corpus, batch_size=8, max_work=8 * 12.0 ** 3, shuffle=True, seed=1, drop_last=False)
    cps = em.constrained_parameters(model)
    sums = [torch.zeros_like(p, dtype=torch.float64) for _, p in cps]
    seen, total = [], 0.0
    for tok, lengths, idx in loader:
        assert not tok.is_cuda and tok.dtype == torch.int32 and (not torch.cuda.is_available() or tok.is_pinned())
        # the tensor-core path is what runs for these batches (N = 64, L <= 129)
        assert engine._padded_cardinality(N, tok.shape[0] * tok.shape[1] * (tok.shape[1] - 1) // 2, tok.shape[1]) == 64
        t, n_ok, counts = em.expected_counts(model, tok, lengths, reduce=False)
        assert int(n_ok) == tok.shape[0]
        total += float(t)
        for s, c in zip(sums, counts):
            s += c.double()
        seen += idx.tolist()
    assert sorted(seen) == list(range(len(corpus)))
    W, T, pr = (x.detach().cpu().double().numpy() for x in model._get_engine().flatten())
    L = max(len(s) for s in corpus)
    tok_all = -np.ones((len(corpus), L), dtype=np.int64)
    for b, s in enumerate(corpus):
        tok_all[b, :len(s)] = s
    lens = np.array([len(s) for s in corpus])
    lz, gW, gT, gP = O.flat_inside_with_grads(W, T, pr, tok_all[:, :, None], lens)
    assert abs(total - float(lz.sum())) <= 1e-4 * abs(float(lz.sum()))
    by_shape = {tuple(c.shape): c.cpu().numpy() for c in sums}
    for got, ref in ((by_shape[(N, N, N)], gW.numpy() * W), (by_shape[(V, N)], gT.numpy() * T), (by_shape[(N,)], gP.numpy() * pr)):
        assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
        assert np.linalg.norm(got - ref) <= 1e-3 * np.linalg.norm(ref)
    assert abs(by_shape[(N, N, N)].sum() - (lens - 1).sum()) < 1e-3 * lens.sum()
