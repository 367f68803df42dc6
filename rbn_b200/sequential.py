"""Sequential (CYK) instantiation of the RBN template (API of /root/reference/rbnet/sequential.py).

`SequentialRBN.inside(sequence)` keeps the reference's call signature, return value (0-dim float64 linear-space
marginal likelihood carrying a grad_fn) and post-call state (`n`, `terminal_chart`, `inside_chart`, `root_location`),
but when every brick is one of the discrete bricks of `rbn_b200.pcfg` the width-major schedule of
sequential.py:23-26, the split enumeration of sequential.py:51-71 and the chart writes of sequential.py:35-36 all
happen inside the span-diagonal CUDA kernels.  Additive, batched entry points: `log_inside_batch`, `inside_batch`.
"""
import warnings

import torch

from .base import RBN, Transition
from .util import ConstrainedModuleList


class HostTemplateWarning(UserWarning):
    """Raised (as a warning) when a model falls outside the fused CUDA path and the Python template loop runs instead."""



class SequentialRBN(RBN, torch.nn.Module):

    def __init__(self, cells, prior, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._cells = ConstrainedModuleList(cells)
        self._prior = prior
        self.n = None
        self._terminal_chart = None
        self._inside_chart = None
        self._outside_chart = None
        self._engine = None

    # ---- reference hooks (sequential.py:18-48) ----------------------------------------------------------------
    def init_inside(self, sequence):
        self.n = len(sequence)
        self._terminal_chart = sequence
        self._inside_chart = [cell.variable.get_chart(self.n) for cell in self._cells]

    def inside_schedule(self, *args, **kwargs):
        for width in range(1, self.n + 1):
            for start in range(0, self.n + 1 - width):
                yield start, start + width

    @RBN.root_location.getter
    def root_location(self):
        return 0, self.n

    def cells(self):
        return self._cells

    def update_inside_chart(self, var_idx, locations, values):
        self._inside_chart[var_idx][locations] = values

    @RBN.inside_chart.getter
    def inside_chart(self):
        chart = self._inside_chart
        return chart.materialise() if hasattr(chart, "materialise") else chart

    @RBN.terminal_chart.getter
    def terminal_chart(self):
        return self._terminal_chart

    @RBN.prior.getter
    def prior(self):
        return self._prior

    # ---- the B200 path ---------------------------------------------------------------------------------------------
    def _get_engine(self):
        from .engine import InsideEngine
        if self._engine is None:
            self._engine = InsideEngine(self)
        return self._engine

    def _fused(self):
        from .engine import InsideEngine
        return InsideEngine.supports(self)

    def inside(self, *args, **kwargs):
        """Marginal likelihood of one sequence (base.py:93-121), linear probability space, float64, differentiable."""
        if not self._fused():
            # Bricks that are not the discrete ones of rbn_b200.pcfg are arbitrary Python callables: only the template
            # loop of base.py:93-121 can evaluate them, on the host, through their own hooks.  That is NOT the CUDA path;
            # say so loudly instead of degrading silently (the discrete bricks themselves never evaluate on the host).
            warnings.warn("SequentialRBN.inside: this model has bricks the fused CUDA path does not cover; running the "
                          "host-side template loop over the bricks' own hooks (slow, not the B200 path)",
                          HostTemplateWarning, stacklevel=2)
            return super().inside(*args, **kwargs)
        sequence = self._prepare_sequence(*args, **kwargs)
        self.n = len(sequence)
        self._terminal_chart = sequence
        engine = self._get_engine()
        Z, lazy_chart = engine.inside_single(sequence)
        self._inside_chart = lazy_chart
        return Z

    def _prepare_sequence(self, sequence):
        return sequence

    def log_inside_batch(self, sequences=None, tokens=None, lengths=None):
        """log Z for a batch (additive API).  Either `sequences` (iterable of reference-style sequences) or a padded
        integer `tokens[B, L]` / `tokens[B, L, n_tvars]` (-1 = None / padding) plus optional `lengths[B]`; the values are
        PER TERMINAL VARIABLE (what a reference sequence holds at `sequence[i][t]`), range-checked like the reference's
        indexing would (IndexError).  Host tensors (e.g. a loader's pinned batch) are copied to the device
        asynchronously and the pass never synchronises; device tensors cost one small device -> host copy.
        Returns a float64 tensor [B] on the parameters' device, differentiable w.r.t. every brick parameter."""
        engine = self._get_engine()
        if tokens is None:
            tokens, lengths = engine.pack_sequences([self._prepare_sequence(s) for s in sequences])
            return engine.log_inside(tokens, lengths, prepacked=True)
        return engine.log_inside(tokens, lengths, prepacked=False)

    def inside_batch(self, sequences=None, tokens=None, lengths=None):
        """Z for a batch in linear space (may underflow for long sequences exactly as exp(log Z) does)."""
        return torch.exp(self.log_inside_batch(sequences, tokens, lengths))

    def log_inside(self, sequence):
        return self.log_inside_batch(sequences=[sequence])[0]

    def viterbi(self, sequences):
        """Most probable derivation of each sequence (additive API, not in the reference): (log probability [B], trees).
        A tree is ((variable, value), start, end, left, right) for a binary node and ((variable, value), start, end) for a
        leaf; None when the sequence has no derivation."""
        return self._get_engine().viterbi([self._prepare_sequence(s) for s in sequences])


class SequentialBinaryTransition(Transition):

    def iterate_inside_splits(self, location):
        """(start, split, end) for every split point strictly inside the span (sequential.py:53-60)."""
        start, end = location
        for split in range(start + 1, end):
            yield start, split, end


class SequentialTerminalTransition(Transition):

    def iterate_inside_splits(self, location):
        """Only width-1 spans have a terminal transition (sequential.py:65-71)."""
        start, end = location
        if end - start == 1:
            yield start
