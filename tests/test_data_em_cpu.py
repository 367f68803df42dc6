"""Host logic around the path that needs no GPU: the batched data path (rbn_b200/data.py) and the M-step /
raw-gradient switch of the EM step (rbn_b200/em.py)."""
import numpy as np
import pytest
import torch

from rbn_b200 import data, em, util


def test_tokenise_follows_reference_semantics():
    toks = data.tokenise(["aab", "b"], terminals=["a", "b"])
    assert [t.tolist() for t in toks] == [[0, 0, 1], [1]]
    with pytest.raises(KeyError):                       # unknown symbol, as PCFG.tokenise (pcfg.py:23)
        data.tokenise(["abc"], terminals=["a", "b"])


def test_pad_batch_and_empty_sequence():
    tok, ln = data.pad_batch([np.array([1, 2, 3]), np.array([4])])
    assert tok.tolist() == [[1, 2, 3], [4, -1, -1]] and ln.tolist() == [3, 1]
    with pytest.raises(ValueError):
        data.pad_batch([np.array([], dtype=np.int32)])


def test_length_buckets_cover_every_sequence_once_and_respect_the_work_cap():
    rng = np.random.default_rng(0)
    seqs = [rng.integers(0, 5, rng.integers(1, 40)) for _ in range(203)]
    seen = []
    for rank in range(2):
        it = data.LengthBucketedBatches(seqs, batch_size=16, max_work=16 * 20.0 ** 3, rank=rank, world=2, drop_last=False,
                                        shuffle=True, seed=3, pin_memory=False)
        for tok, ln, idx in it:
            assert tok.shape[0] == ln.shape[0] == idx.shape[0] <= 16
            assert tok.shape[1] == int(ln.max())
            assert tok.shape[0] * float(ln.max()) ** 3 <= 16 * 20.0 ** 3 or tok.shape[0] == 1
            for b, i in enumerate(idx.tolist()):
                assert tok[b, :ln[b]].tolist() == list(seqs[i]) and (tok[b, ln[b]:] == -1).all()
            assert int(ln.max()) - int(ln.min()) <= 20      # a bucket never mixes very short and very long sequences
            seen += idx.tolist()
    assert set(seen) == set(range(203))                     # drop_last=False: everything is visited (tail repeated)
    a = data.LengthBucketedBatches(seqs, batch_size=16, rank=0, world=2, pin_memory=False)
    b = data.LengthBucketedBatches(seqs, batch_size=16, rank=1, world=2, pin_memory=False)
    assert len(a) == len(b)                                 # equal number of steps per rank (collectives line up)


def test_raw_gradients_switches_the_projection_hook_off():
    lp = util.LogProb(p=torch.tensor([[0.2, 0.5], [0.8, 0.5]], dtype=torch.float64), dim=0)
    w = torch.tensor([[1.0, 2.0], [3.0, 4.0]], dtype=torch.float64)
    (lp.log_p * w).sum().backward()
    np.testing.assert_allclose(lp.log_p.grad.numpy(), [[-1, -1], [1, 1]])          # projected (util.py:200-202)
    lp.log_p.grad = None
    with util.raw_gradients():
        (lp.log_p * w).sum().backward()
    np.testing.assert_allclose(lp.log_p.grad.numpy(), w.numpy())                    # raw = expected counts
    lp.log_p.grad = None
    (lp.log_p * w).sum().backward()
    np.testing.assert_allclose(lp.log_p.grad.numpy(), [[-1, -1], [1, 1]])          # restored on exit


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = util.LogProb(p=torch.full((3, 2), 1 / 3, dtype=torch.float64), dim=0)
        self.b = util.Prob(torch.tensor([[0.5, 0.5], [0.25, 0.75]], dtype=torch.float64), dim=1)


def test_m_step_normalises_counts_over_the_constrained_dims():
    m = _Toy()
    old_b_row1 = m.b.p.detach()[1].clone()
    counts = [torch.tensor([[1.0, 0.0], [1.0, 0.0], [2.0, 0.0]], dtype=torch.float64),      # column 1 got no count
              torch.tensor([[3.0, 1.0], [0.0, 0.0]], dtype=torch.float64)]                  # row 1 got no count
    untouched = em.m_step(m, counts)
    assert untouched == 2
    np.testing.assert_allclose(m.a.p.detach().numpy()[:, 0], [0.25, 0.25, 0.5])
    np.testing.assert_allclose(m.a.p.detach().numpy()[:, 1], [1 / 3] * 3)                   # kept
    np.testing.assert_allclose(m.b.p.detach().numpy()[0], [0.75, 0.25])
    np.testing.assert_allclose(m.b.p.detach().numpy()[1], old_b_row1.numpy())               # kept
    em.m_step(m, counts, pseudo_count=1.0)
    np.testing.assert_allclose(m.a.p.detach().numpy()[:, 1], [1 / 3] * 3)
    np.testing.assert_allclose(m.a.p.detach().numpy()[:, 0], [2 / 7, 2 / 7, 3 / 7])


def test_fewer_batches_than_ranks_still_gives_every_rank_the_same_count():
    """drop_last=False with 1 batch and 4 ranks: the batch list wraps around, every rank gets one batch."""
    seqs = [np.array([1, 2, 3]), np.array([1])]
    lens = [len(data.LengthBucketedBatches(seqs, batch_size=16, rank=r, world=4, drop_last=False, pin_memory=False)) for r in range(4)]
    assert lens == [1, 1, 1, 1]
    seqs = [np.arange(1 + (i % 7)) for i in range(80)]
    lens = [len(data.LengthBucketedBatches(seqs, batch_size=16, rank=r, world=4, drop_last=False, pin_memory=False)) for r in range(4)]
    assert len(set(lens)) == 1 and lens[0] == 2
