"""Golden fixtures at the BENCHMARKED shapes (BASELINE configs[2] and configs[4]): the fp64 oracle
(oracle/inside_oracle.py::flat_inside_with_grads, the pinned restatement of /root/reference/rbnet/pcfg.py:305-314 and
base.py:93-121 with torch.autograd as the outside pass) run ONCE in the build container on seeded inputs.

    python -m oracle.make_golden_headline            # writes tests/golden/headline_*.npz

A full fp64 gradient of W is N^3 entries (134 MB at N = 256, 1 GB at N = 512) -- too large to commit -- so a fixture
keeps, in fp64: log Z, gT and gPrior in full, every STRIDE-th entry of gW (a fixed arithmetic sample of ~130 k
entries), the three marginals of gW (sums over two of its axes, [3, N]), and |gW|_2, sum(gW), max(gW).  The GPU test
(tests/test_gpu_headline.py) compares the same statistics of the CUDA path's gradient; set RBN_FULL_ORACLE=1 there to
re-run the oracle live on the GPU box's host cores and compare all N^3 entries instead.

The inputs are the very batches bench.py builds (rbn_b200.synth, seeds 0 / 1), so sequence 0 of the C3 fixture is
sequence 0 of the benched batch and the bench's own `parity_check` can be cross-checked against it.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "tests", "golden")

# name: (N, V, L, grammar seed, token seed, lengths of the sequences taken from the seeded batch)
CASES = {
    "headline_c3": dict(N=256, V=32, L=128, gseed=0, tseed=1, lengths=[128, 101]),     # one full, one ragged
    "headline_c5": dict(N=512, V=32, L=128, gseed=0, tseed=1, lengths=[128]),
}


def sample_stride(N):
    n3 = N ** 3
    return max(1, n3 // 131072) | 1          # odd stride: walks every residue of the (b, c, a) index


def inputs(case):
    from rbn_b200 import synth
    W, T, pr = synth.synthetic_grammar(case["N"], case["V"], seed=case["gseed"])
    tok, _ = synth.synthetic_tokens(len(case["lengths"]), case["L"], case["V"], seed=case["tseed"])
    lengths = np.asarray(case["lengths"], dtype=np.int64)
    for b, n in enumerate(lengths):
        tok[b, n:] = -1
    return W, T, pr, tok, lengths


def grad_stats(gW, stride):
    """The statistics a fixture keeps of a [N, N, N] gradient (any float dtype; accumulated in fp64)."""
    g = np.asarray(gW)
    g64 = g.astype(np.float64, copy=False) if g.dtype == np.float64 else None
    flat = g.reshape(-1)
    out = dict(sample=flat[::stride].astype(np.float64))
    if g64 is None:
        # accumulate slab by slab so that a 1 GB fp32 tensor never needs a 2 GB fp64 copy
        N = g.shape[0]
        m0, m1, m2 = np.zeros(N), np.zeros(N), np.zeros(N)
        l2 = 0.0
        mx = 0.0
        for b in range(N):
            s = g[b].astype(np.float64)
            m0[b] = s.sum()
            m1 += s.sum(axis=1)
            m2 += s.sum(axis=0)
            l2 += float((s * s).sum())
            mx = max(mx, float(np.abs(s).max()))
        out.update(marg=np.stack([m0, m1, m2]), l2=np.sqrt(l2), total=m0.sum(), amax=mx)
    else:
        out.update(marg=np.stack([g64.sum(axis=(1, 2)), g64.sum(axis=(0, 2)), g64.sum(axis=(0, 1))]),
                   l2=float(np.sqrt((g64 * g64).sum())), total=float(g64.sum()), amax=float(np.abs(g64).max()))
    return out


def main(only=None):
    import torch
    from oracle import inside_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        W, T, pr, tok, lengths = inputs(case)
        t0 = time.perf_counter()
        logZ, gW, gT, gP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
        dt = time.perf_counter() - t0
        stride = sample_stride(case["N"])
        st = grad_stats(gW.numpy(), stride)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), logZ=logZ.numpy(), gT=gT.numpy(), gP=gP.numpy(),
                            gW_sample=st["sample"], gW_marg=st["marg"], gW_l2=st["l2"], gW_total=st["total"],
                            gW_amax=st["amax"], stride=stride, lengths=lengths,
                            meta=np.array([case["N"], case["V"], case["L"], case["gseed"], case["tseed"]]))
        print(f"{name}: logZ={logZ.numpy()} oracle {dt:.1f}s, |gW|={st['l2']:.6e}, {st['sample'].size} sampled entries", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
