"""EM step on top of the fused chart pass: the step that follows the hot path in a training loop.

E-step  expected rule counts = raw gradient of sum_b log Z_b w.r.t. the log-parameters, from the CUDA outside pass
        (micro-batched so that chart + outside chart fit HBM, counts accumulate in `.grad`), then ONE flat all-reduce
        over the ranks (`rbn_b200.distributed.allreduce_counts`).
M-step  every constrained parameter becomes its normalised expected counts over the dims the reference normalises
        (`Prob.enforce_constraints` / `LogProb.enforce_constraints`, /root/reference/rbnet/util.py:158-167, 193-198);
        `LogProb` tensors on the GPU (float32 or the API's float64) are normalised by the fused kernel behind `rbn_lognormalize` / `rbn_lognormalize_f64`.

The reference has no EM loop of its own; what it prescribes is the constraint set (which dims sum to one) and the
gradient these counts are.  Groups that received no count keep their old distribution.
"""
import ctypes

import torch

from . import _lib
from .util import LogProb, Prob, raw_gradients


def constrained_parameters(model):
    """[(module, parameter)] for every `Prob` / `LogProb` in the model, in `named_modules()` order."""
    out = []
    for _, m in model.named_modules():
        if isinstance(m, LogProb):
            out.append((m, m.log_p))
        elif isinstance(m, Prob):
            out.append((m, m.p))
    return out


def expected_counts(model, tokens=None, lengths=None, sequences=None, micro_batch=512, reduce=True):
    """E-step.  Returns (sum of log Z over this rank's finite sequences, number of them, [counts per constrained
    parameter]).  Sequences with Z = 0 (log Z = -inf) contribute nothing, as in the reference where their gradient
    is undefined.  With an initialised process group and `reduce=True` the counts (and the two scalars) are summed
    over ranks with one flat all-reduce."""
    if tokens is None:
        engine = model._get_engine()
        tokens, lengths = engine.pack_sequences([model._prepare_sequence(s) for s in sequences])
    tokens = torch.as_tensor(tokens)
    B = tokens.shape[0]
    if lengths is None:
        lengths = torch.full((B,), tokens.shape[1], dtype=torch.int32)
    lengths = torch.as_tensor(lengths)
    cps = constrained_parameters(model)
    for _, p in cps:
        p.grad = None
    total = torch.zeros((), dtype=torch.float64, device=cps[0][1].device)
    n_ok = torch.zeros((), dtype=torch.float64, device=cps[0][1].device)
    with raw_gradients():
        for b0 in range(0, B, micro_batch):
            lz = model.log_inside_batch(tokens=tokens[b0:b0 + micro_batch], lengths=lengths[b0:b0 + micro_batch])
            ok = torch.isfinite(lz)
            lz_ok = torch.where(ok, lz, torch.zeros_like(lz))
            lz_ok.sum().backward()
            total += lz_ok.detach().sum()
            n_ok += ok.sum()
    counts = []
    for m, p in cps:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        counts.append((g * p.detach()) if isinstance(m, Prob) else g.clone())   # d/dp * p = d/dlog p = count
        p.grad = None
    if reduce and torch.distributed.is_available() and torch.distributed.is_initialized():
        from .distributed import allreduce_counts
        scal = torch.stack([total, n_ok])
        allreduce_counts(counts + [scal])
        total, n_ok = scal[0], scal[1]
    return total, n_ok, counts


def _lognormalize_(log_p, dims):
    """log_p -= logsumexp over `dims`, in place.  CUDA tensors (fp32 / fp64) whose normalised dims are the LEADING ones go
    through the fused kernel (rbn_lognormalize: element (g, k) at log_p[k * groups + g])."""
    nd = log_p.dim()
    lead = tuple(range(len(dims)))
    if log_p.is_cuda and log_p.dtype in (torch.float32, torch.float64) and log_p.is_contiguous() and tuple(sorted(dims)) == lead and nd >= len(dims):
        group = 1
        for d in dims:
            group *= log_p.shape[d]
        groups = log_p.numel() // max(group, 1)
        lib = _lib.load()
        with torch.cuda.device(log_p.device):
            fn = lib.rbn_lognormalize if log_p.dtype == torch.float32 else lib.rbn_lognormalize_f64
            rc = fn(ctypes.c_void_p(log_p.data_ptr()), group, groups, None,
                    ctypes.c_void_p(torch.cuda.current_stream(log_p.device).cuda_stream))
        _lib.check(rc, "rbn_lognormalize")
        return log_p
    return log_p.sub_(torch.logsumexp(log_p, dim=dims, keepdim=True))


def m_step(model, counts, pseudo_count=0.0):
    """M-step: parameters <- normalised (counts + pseudo_count).  Returns the number of normalisation groups that had
    no mass and were left untouched."""
    untouched = 0
    with torch.no_grad():
        for (m, p), c in zip(constrained_parameters(model), counts):
            dims = m._dims()
            c = c.to(p.dtype).clamp_min(0) + pseudo_count
            mass = c.sum(dim=dims, keepdim=True)
            empty = mass <= 0
            untouched += int(empty.sum())
            if isinstance(m, Prob):
                new = torch.where(empty, p, c / torch.where(empty, torch.ones_like(mass), mass))
                p.copy_(new)
            else:
                logc = torch.where(c > 0, torch.log(c.clamp_min(torch.finfo(p.dtype).tiny)), torch.full_like(c, -float("inf")))
                logc = torch.where(empty, p, logc)          # groups without mass keep their (normalised) old values
                p.copy_(_lognormalize_(logc.contiguous(), dims))
    return untouched


def em_step(model, tokens=None, lengths=None, sequences=None, micro_batch=512, pseudo_count=0.0):
    """One EM iteration; returns the mean log Z (over all ranks' finite sequences) BEFORE the update."""
    total, n_ok, counts = expected_counts(model, tokens, lengths, sequences, micro_batch)
    m_step(model, counts, pseudo_count)        # (the engine re-flattens the grammar on every pass: nothing to invalidate)
    return float(total / torch.clamp(n_ok, min=1.0))
