"""Constrained probability parameters and small host-side helpers (API of /root/reference/rbnet/util.py).

Only what the inside / marginal-likelihood path needs is mirrored: `Prob`, `LogProb` (the tensors gradients must
land on, /root/reference/rbnet/util.py:121-210), the cooperative `enforce_constraints` mixin (util.py:66-118),
`normalize_non_zero` (used by the grammar front-end, pcfg.py:99-100), `as_detached_tensor`,
`ensure_is_floating_point`, `log_normalize`.  Plotting helpers, `TupleTMap` (autoencoder container) and the
single-item `SequenceDataModule` are out of scope.
"""
import math

import numpy as np
import torch


def ensure_is_floating_point(t, msg=None):
    """Return `t` as a new tensor; TypeError unless floating point (util.py:329-335)."""
    t = t.detach().clone() if torch.is_tensor(t) else torch.tensor(t)
    if not torch.is_floating_point(t):
        raise TypeError(msg or "Values must be of floating point type")
    return t


def as_detached_tensor(t):
    """Detached copy of `t` as a tensor (util.py:338-348)."""
    if torch.is_tensor(t):
        return t.detach().clone()
    return torch.tensor(np.array(t))


def log_normalize(t, *args, **kwargs):
    """t - logsumexp(t, ...) (util.py:351-360)."""
    return t - torch.logsumexp(t, *args, **kwargs)


_AXIS_UNSET = object()


def normalize_non_zero(a, axis=_AXIS_UNSET, make_zeros_uniform=False, skip_type_check=False, check_positivity=True):
    """Normalise slices of `a` along `axis` (int, tuple or None = everything) wherever their sum is non-zero.

    Same contract as /root/reference/rbnet/util.py:244-326: numpy or torch input, in-place division of the
    non-zero slices, `make_zeros_uniform=True` first replaces all-zero slices by ones (so they become uniform),
    TypeError for non-float arrays unless `skip_type_check`, ValueError for negative entries.
    """
    is_np = isinstance(a, np.ndarray)
    if not is_np and not torch.is_tensor(a):
        raise TypeError(f"Not implemented for arrays of type {type(a)}")
    if not skip_type_check:
        is_float = np.issubdtype(a.dtype, np.floating) if is_np else torch.is_floating_point(a)
        if not is_float:
            raise TypeError(f"Cannot guarantee that normalisation works as expected on array of type '{a.dtype}'. "
                            f"Use 'skip_type_check=True' to skip this check.")
    if check_positivity and bool((a < 0).any()):
        raise ValueError("Some elements are negative")
    if axis is _AXIS_UNSET:
        import warnings
        warnings.warn("Not passing an explicit value to 'axis' is deprecated; the last axis is assumed.",
                      DeprecationWarning)
        axis = a.ndim - 1
    if axis is None:
        axis = tuple(range(a.ndim))
    if not isinstance(axis, tuple):
        axis = (axis,)
    total = a.sum(axis=axis, keepdims=True) if is_np else a.sum(dim=axis, keepdim=True)
    zero = total == 0
    if bool(zero.any()) and make_zeros_uniform:
        fill = np.broadcast_to(zero, a.shape) if is_np else zero.expand_as(a)
        a = a.copy() if is_np else a.clone()
        a[fill] = 1
        total = a.sum(axis=axis, keepdims=True) if is_np else a.sum(dim=axis, keepdim=True)
        zero = total == 0
    if is_np:
        safe = np.where(zero, 1, total)
    else:
        safe = torch.where(zero, torch.ones_like(total), total)
    a /= safe
    return a


class ConstrainedModuleMixin:
    """Cooperative constraint handling: `enforce_constraints()` walks the module tree (util.py:66-113)."""

    def enforce_constraints(self, recurse=True):
        if not recurse:
            return
        for child in self.children():
            fn = getattr(child, "enforce_constraints", None)
            if fn is not None:
                fn()

    def remap(self, param, _top_level=True, prefix=None):
        for child in self.children():
            fn = getattr(child, "remap", None)
            if fn is None:
                continue
            try:
                found = fn(param, _top_level=False)
            except KeyError:
                continue
            return f"{prefix}{found}" if prefix is not None else found
        if not _top_level:
            raise KeyError
        return str(param) if prefix is not None else param


class ConstrainedModuleList(torch.nn.ModuleList, ConstrainedModuleMixin):
    """ModuleList that takes part in `enforce_constraints` recursion (util.py:116-118)."""


_PROJECT_GRADIENTS = True


class raw_gradients:
    """Context manager: inside it the `project_grad` hooks of `Prob` / `LogProb` pass gradients through unchanged.
    With `LogProb`, the raw gradient of sum_b log Z_b w.r.t. `log_p` IS the expected rule count,
    which the EM step (rbn_b200/em.py) needs; the reference's hook would subtract its mean (util.py:200-202)."""

    def __enter__(self):
        global _PROJECT_GRADIENTS
        self._saved = _PROJECT_GRADIENTS
        _PROJECT_GRADIENTS = False
        return self

    def __exit__(self, *exc):
        global _PROJECT_GRADIENTS
        _PROJECT_GRADIENTS = self._saved
        return False


def _project_to_tangent(grad, dims):
    """g - mean of g over the normalised dims: the reference's gradient projection (util.py:147-156, 200-202)."""
    if not _PROJECT_GRADIENTS:
        return grad
    count = math.prod(grad.shape[d] for d in dims)
    return grad - grad.sum(dim=dims, keepdim=True) / count


class Prob(torch.nn.Module, ConstrainedModuleMixin):
    """Probabilities stored directly; clipped at zero and renormalised over `dim` (util.py:121-167)."""

    def __init__(self, p, dim=None, raise_zero_norms=True, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.p = torch.nn.Parameter(ensure_is_floating_point(p, "Probabilities 'p' must be of floating point type"))
        self.p.register_hook(self.project_grad)
        self.dim = tuple(range(self.p.dim())) if dim is None else dim
        self.raise_zero_norms = raise_zero_norms
        self.enforce_constraints()

    def _dims(self):
        return self.dim if isinstance(self.dim, tuple) else (self.dim,)

    def project_grad(self, grad):
        return _project_to_tangent(grad, self._dims())

    def enforce_constraints(self, recurse=True):
        with torch.no_grad():
            self.p.clamp_(min=0)
            norm = self.p.sum(dim=self._dims(), keepdim=True)
            if self.raise_zero_norms and bool((norm == 0).any()):
                raise RuntimeError("Some normalisation constants are zero")
            self.p.div_(norm)


class LogProb(torch.nn.Module, ConstrainedModuleMixin):
    """Probabilities stored as unconstrained logs, `p = exp(log_p)` (util.py:170-210)."""

    def __init__(self, p=None, log_p=None, dim=None, raise_zero_norms=True, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if (p is None) == (log_p is None):
            raise ValueError("Have to provide exactly one of 'p' and 'log_p'")
        if log_p is None:
            log_p = torch.log(torch.as_tensor(p))
        self.log_p = torch.nn.Parameter(log_p)
        self.log_p.register_hook(self.project_grad)
        self.dim = tuple(range(self.log_p.dim())) if dim is None else dim
        self.raise_zero_norms = raise_zero_norms
        self.enforce_constraints()

    def _dims(self):
        return self.dim if isinstance(self.dim, tuple) else (self.dim,)

    def remap(self, param, _top_level=True, prefix=None):
        if param is self.log_p:
            return self.p
        raise KeyError

    def project_grad(self, grad):
        return _project_to_tangent(grad, self._dims())

    def enforce_constraints(self, recurse=True):
        with torch.no_grad():
            lse = torch.logsumexp(self.log_p, dim=self._dims(), keepdim=True)
            if self.raise_zero_norms and bool(torch.isinf(lse).any()):
                raise RuntimeError("Some log-normalisation constants are inf")
            self.log_p.sub_(lse)

    @property
    def p(self):
        return torch.exp(self.log_p)

    @p.setter
    def p(self, value):
        self.log_p = torch.log(value)
