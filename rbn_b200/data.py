"""Batched data path in front of the chart pass.

The reference feeds `inside()` one Python sequence at a time (`SequenceDataModule`, batch_size = 1,
/root/reference/rbnet/util.py:213-240; `PCFG.tokenise`, pcfg.py:22-28).  To keep B >= 512 sequences per launch
train busy the host side here tokenises once, buckets by length (chart work grows with L^3, and a batch is padded
to its longest member) and hands out padded int32 `tokens[B, L]` + `lengths[B]` in pinned memory.
"""
import numpy as np
import torch


def tokenise(sequences, terminals):
    """Symbols -> indices with the reference's semantics (pcfg.py:22-28): `terminals` is the ordered alphabet, an
    unknown symbol raises KeyError.  Returns a list of int32 arrays."""
    index = {s: i for i, s in enumerate(terminals)}
    return [np.fromiter((index[s] for s in seq), dtype=np.int32, count=len(seq)) for seq in sequences]


def pad_batch(token_lists, pad_to=None):
    """List of int arrays -> (tokens[B, L] int32 with -1 padding, lengths[B] int32)."""
    B = len(token_lists)
    lengths = np.fromiter((len(t) for t in token_lists), dtype=np.int32, count=B)
    if B and lengths.min() < 1:
        raise ValueError("cannot parse an empty sequence")
    L = int(pad_to or (lengths.max() if B else 0))
    tokens = -np.ones((B, L), dtype=np.int32)
    for b, t in enumerate(token_lists):
        tokens[b, :len(t)] = t
    return tokens, lengths


class LengthBucketedBatches:
    """Iterable of (tokens, lengths, indices) batches.

    Sequences are sorted by length and cut into batches whose padded chart work sum_b L_max^3 stays below
    `max_work` (and whose size stays below `batch_size`), so a batch never mixes very short and very long sequences;
    the batch ORDER is shuffled per epoch when `shuffle=True` (content of a batch is fixed, like bucketed samplers).
    `rank` / `world` deal the batches round-robin to the ranks of a data-parallel job (equal counts per rank: the tail
    that does not divide is dropped, or kept and the list wrapped around to a multiple of `world` with `drop_last=False`)."""

    def __init__(self, token_lists, batch_size=512, max_work=None, shuffle=False, seed=0, rank=0, world=1,
                 drop_last=True, pin_memory=True):
        self.items = [np.asarray(t, dtype=np.int32) for t in token_lists]
        if any(len(t) < 1 for t in self.items):
            raise ValueError("cannot parse an empty sequence")
        order = np.argsort([len(t) for t in self.items], kind="stable")
        self.batches = []
        cur = []
        for idx in order:
            L = len(self.items[idx])
            full = len(cur) >= batch_size or (max_work is not None and cur and (len(cur) + 1) * float(L) ** 3 > max_work)
            if full:
                self.batches.append(np.array(cur))
                cur = []
            cur.append(int(idx))
        if cur:
            self.batches.append(np.array(cur))
        self.shuffle, self.seed, self.rank, self.world = shuffle, seed, rank, world
        self.drop_last, self.epoch = drop_last, 0
        self.pin = pin_memory and torch.cuda.is_available()

    def set_epoch(self, epoch):
        self.epoch = epoch

    def _my_batches(self):
        ids = np.arange(len(self.batches))
        if self.shuffle:
            ids = np.random.default_rng(self.seed + self.epoch).permutation(ids)
        per = len(ids) // self.world
        if self.drop_last or len(ids) % self.world == 0:
            ids = ids[:per * self.world]
        else:
            ids = np.resize(ids, (per + 1) * self.world)      # wraps around: also right when there are fewer batches than ranks
        return ids[self.rank::self.world]

    def __len__(self):
        return len(self._my_batches())

    def __iter__(self):
        for bi in self._my_batches():
            idx = self.batches[bi]
            tokens, lengths = pad_batch([self.items[i] for i in idx])
            t, l = torch.from_numpy(tokens), torch.from_numpy(lengths)
            if self.pin:
                t, l = t.pin_memory(), l.pin_memory()
            yield t, l, torch.from_numpy(idx.astype(np.int64))
