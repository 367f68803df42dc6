"""Multi-GPU plumbing of the chart pass: shard the batch of sequences, all-reduce the expected rule counts.

Sequences are independent (the reference parses them one at a time and sums nothing across them,
/root/reference/rbnet/base.py:93-121), so the path shards by sequence: every rank owns a contiguous slice of the
(length-sorted) batch and the full parameter set; the only exchange is one sum all-reduce of the gradient /
expected-count tensors per step plus a gather of log Z.  No collective touches chart data.  One process per GPU,
`torch.distributed` (NCCL on GPUs; the same code runs over gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_order(lengths, world):
    """Permutation that deals sequences to ranks so that every rank gets (almost) the same chart work.

    Work of a sequence of length n is ~ n^3 (number of (span, split) pairs).  Sequences are sorted by length and
    dealt round-robin in a serpentine; rank r then owns positions [shard_bounds(...)] of the returned order."""
    lengths = torch.as_tensor(lengths)
    order = torch.argsort(lengths, descending=True, stable=True)
    per_rank = [[] for _ in range(world)]
    for pos, idx in enumerate(order.tolist()):
        rnd, k = divmod(pos, world)
        per_rank[k if rnd % 2 == 0 else world - 1 - k].append(idx)
    return torch.tensor([i for r in per_rank for i in r], dtype=torch.int64), [len(r) for r in per_rank]


def shard_batch(tokens, lengths, rank=None, world=None, balance=True):
    """This rank's (tokens, lengths, global indices).  tokens [B, L(, n_tvars)], lengths [B]."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    tokens, lengths = torch.as_tensor(tokens), torch.as_tensor(lengths)
    if balance:
        order, counts = balanced_order(lengths, world)
        lo = sum(counts[:rank])
        idx = order[lo:lo + counts[rank]]
    else:
        lo, hi = shard_bounds(tokens.shape[0], rank, world)
        idx = torch.arange(lo, hi)
    tok, ln = tokens[idx], lengths[idx]
    keep = int(ln.max()) if ln.numel() else 0
    return tok[:, :keep], ln, idx


def allreduce_counts(tensors, group=None):
    """Sum the expected-count / gradient tensors over ranks IN PLACE: one all-reduce per tensor, issued together on the
    process group's own stream (NCCL), no staging copy -- the big one (gW, 67 MB at N = 256, 537 MB at N = 512) moves at
    NVLink rate, the small ones cost a launch each."""
    tensors = [t for t in tensors if t is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True) for t in tensors]
    for w in works:
        w.wait()
    return tensors


class CountReducer:
    """Accumulates the count tensors of successive micro-batches; with several ranks every micro-batch's tensors are
    all-reduced asynchronously (in place, on the process group's stream) WHILE the next micro-batch's chart pass runs,
    and folded into the running sum one micro-batch later.  Summing per micro-batch moves a few times more bytes over
    NVLink than one all-reduce at the end (C5: 2 x 537 MB per step against seconds of compute) and removes the
    collective from the critical path.

        red = CountReducer()
        for micro-batch: red.add(torch.autograd.grad(logZ.sum(), (W, T, prior)))
        gW, gT, gPrior = red.result()
    """

    def __init__(self, group=None):
        self.group = group
        self.acc = None
        self.pending = None          # (tensors, works) of the micro-batch whose all-reduce is in flight
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def _fold(self, tensors):
        if self.acc is None:
            self.acc = [t if t.is_contiguous() else t.contiguous() for t in tensors]
        else:
            for a, t in zip(self.acc, tensors):
                a.add_(t)

    def _drain(self):
        if self.pending is not None:
            tensors, works = self.pending
            for w in works:
                w.wait()
            self.pending = None
            self._fold(tensors)

    def add(self, tensors):
        tensors = [t.detach() for t in tensors]
        if not self.distributed:
            self._fold(tensors)
            return
        self._drain()
        tensors = [t if t.is_contiguous() else t.contiguous() for t in tensors]
        works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for t in tensors]
        self.pending = (tensors, works)

    def result(self):
        self._drain()
        return self.acc


def gather_logZ(logZ, idx, total, group=None):
    """All ranks' log Z in the ORIGINAL batch order: every rank contributes its values at its global indices."""
    out = torch.zeros(total, dtype=logZ.dtype, device=logZ.device)
    out[idx.to(logZ.device)] = logZ.detach()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def data_parallel_step(model, tokens, lengths, group=None):
    """One inside + outside step of `model` (a rbn_b200 SequentialRBN) on this rank's shard; gradients of
    sum_b log Z_b over the WHOLE batch end up on every rank's Parameters.  Returns log Z of the whole batch."""
    tok, ln, idx = shard_batch(tokens, lengths, balance=True)
    params = [p for p in model.parameters()]
    for p in params:
        p.grad = None
    if tok.shape[0] > 0:
        lz = model.log_inside_batch(tokens=tok, lengths=ln)
        lz.sum().backward()
    else:
        lz = torch.zeros(0, dtype=torch.float64, device=params[0].device)
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    allreduce_counts([p.grad for p in params], group)
    return gather_logZ(lz, idx, torch.as_tensor(tokens).shape[0], group)
