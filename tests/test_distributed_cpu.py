"""N > 1 host logic over gloo (world_size 2, CPU): sharding, count all-reduce, log Z gather.  The per-shard chart
pass is played by the fp64 oracle here (the CUDA path needs a GPU); what is under test is rbn_b200.distributed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rbn_b200 import distributed as D


def test_shard_bounds_and_balanced_order():
    assert [D.shard_bounds(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert [D.shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    lengths = torch.tensor([5, 128, 64, 64, 3, 100, 90, 7])
    order, counts = D.balanced_order(lengths, 2)
    assert sorted(order.tolist()) == list(range(8)) and counts == [4, 4]
    work = [(lengths[order[:4]].double() ** 3).sum(), (lengths[order[4:]].double() ** 3).sum()]
    assert abs(work[0] - work[1]) / max(work) < 0.35
    tok = torch.arange(8 * 128).reshape(8, 128)
    t0, l0, i0 = D.shard_batch(tok, lengths, rank=0, world=2)
    t1, l1, i1 = D.shard_batch(tok, lengths, rank=1, world=2)
    assert sorted(i0.tolist() + i1.tolist()) == list(range(8))
    assert t0.shape[1] == int(l0.max()) and t1.shape[1] == int(l1.max())


def _worker(rank, world, port, W, T, pr, tok, lengths, q):
    from oracle import inside_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, ln, idx = D.shard_batch(torch.tensor(tok), torch.tensor(lengths))
    lz, gW, gT, gP = O.flat_inside_with_grads(W, T, pr, t.numpy()[:, :, None], ln.numpy())
    D.allreduce_counts([gW, gT, gP])
    full = D.gather_logZ(lz, idx, tok.shape[0])
    # the same counts through the micro-batched, overlapped reducer (two micro-batches per rank)
    red = D.CountReducer()
    half = t.shape[0] // 2
    for sl in (slice(0, half), slice(half, None)):
        _, *g = O.flat_inside_with_grads(W, T, pr, t.numpy()[sl][:, :, None], ln.numpy()[sl])
        red.add(g)
    for a, b in zip(red.result(), (gW, gT, gP)):
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-12)
    if rank == 0:
        q.put((full.numpy(), gW.numpy(), gT.numpy(), gP.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_over_gloo():
    from oracle import inside_oracle as O
    N, V, L, B = 5, 4, 9, 7
    W, T, pr = O.synthetic_grammar(N, V, seed=2)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=3, ragged=True)
    ref_lz, ref_gW, ref_gT, ref_gP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, W, T, pr, tok, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    lz, gW, gT, gP = q.get(timeout=240)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    np.testing.assert_allclose(lz, ref_lz.numpy(), rtol=1e-12)
    np.testing.assert_allclose(gW, ref_gW.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gT, ref_gT.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gP, ref_gP.numpy(), rtol=1e-10, atol=1e-12)
