"""`MultivariateNormal`: the wrapper surface of /root/reference/rbnet/multivariate_normal.py:14-191, rewritten.

Same constructor (loc; exactly one of covariance_matrix / precision_matrix / scale_tril; at most one of norm / log_norm;
`dim` for scalar locations), the same three parametrisations of the matrix argument -- full `[..., D, D]`, diagonal
`[..., D]` and scalar `[...]`, told apart by the rank of the matrix relative to the location as the reference does
(:150) --, the lazily built `torch.distributions.MultivariateNormal` behind `.torch`, and `__mul__`: the product of two
Gaussian densities as one normalised Gaussian times a constant (the reference routes it through PairwiseProduct,
:183-191; the closed form is the one of :332-396, which is also what `csrc/gauss_kernels.cu` evaluates per split
inside the chart pass and what oracle/gauss_oracle.py::pairwise_product restates):

    N(x; m1, S1) N(x; m2, S2) = N(m1; m2, S1 + S2) * N(x; m, C),   m = m1 + S1 (S1+S2)^-1 (m2 - m1),   C = S1 - S1 (S1+S2)^-1 S1

This class is host-level API (plain torch ops on whatever device its tensors live on); the chart pass itself never
goes through it -- `GaussianRBN` hands covariances to the CUDA kernels as dense matrices.
"""
import math

import torch
from torch.distributions import MultivariateNormal as _TorchMVN


class MultivariateNormal:
    FULL, DIAG, SCAL = "__FULL__", "__DIAG__", "__SCAL__"

    @staticmethod
    def _as_tensor(t):
        return t if isinstance(t, torch.Tensor) else torch.as_tensor(t, dtype=torch.get_default_dtype())

    @staticmethod
    def expand_matrix(mat, kind, dim):
        """Scalar `[...]` / diagonal `[..., D]` / full `[..., D, D]` specification -> dense `[..., D, D]`."""
        if kind == MultivariateNormal.SCAL:
            return mat[..., None, None] * torch.eye(dim, dtype=mat.dtype, device=mat.device)
        if kind == MultivariateNormal.DIAG:
            if mat.shape[-1] != dim:
                raise ValueError(f"Last dimension of tensor has size {mat.shape[-1]}, cannot expand to diagonal of size {dim}")
            return torch.diag_embed(mat)
        return mat

    # names the reference's own tests call (tests/test_multivariate_normal.py:35-66)
    @classmethod
    def _expand_scalar(cls, diag, dim):
        return cls.expand_matrix(cls._as_tensor(diag), cls.SCAL, dim)

    @classmethod
    def _expand_diagonal(cls, diag, dim=None):
        diag = cls._as_tensor(diag)
        return cls.expand_matrix(diag, cls.DIAG, diag.shape[-1] if dim is None else dim)

    _to_tensor = _as_tensor

    def __init__(self, loc, covariance_matrix=None, precision_matrix=None, scale_tril=None, norm=None, log_norm=None,
                 validate_args=None, dim=False):
        if (norm is not None) + (log_norm is not None) > 1:
            raise ValueError("At most one of 'norm' and 'log_norm' may be specified.")
        if log_norm is not None:
            self._log_norm = self._as_tensor(log_norm)
        elif norm is not None:
            self._log_norm = torch.log(self._as_tensor(norm))
        else:
            self._log_norm = self._as_tensor(0.)
        self._loc = self._as_tensor(loc)
        if dim:                                   # scalar location(s) broadcast over an explicit event dimension
            self._event_dim = int(dim)
            self._expanded_loc = self._loc[..., None].expand(self._loc.shape + (self._event_dim,))
        else:
            if self._loc.dim() < 1:
                raise ValueError("a scalar location needs the event dimensionality: pass dim=")
            self._event_dim = int(self._loc.shape[-1])
            self._expanded_loc = self._loc
        given = [(k, m) for k, m in (("scale_tril", scale_tril), ("covariance_matrix", covariance_matrix),
                                     ("precision_matrix", precision_matrix)) if m is not None]
        if len(given) != 1:
            raise ValueError("Exactly one of covariance_matrix or precision_matrix or scale_tril must be specified.")
        self._mat_name, mat = given[0][0], self._as_tensor(given[0][1])
        rank = mat.dim() - self._expanded_loc.dim() + 1          # 0 scalar, 1 diagonal, 2 full (reference :150)
        if rank not in (0, 1, 2):
            raise ValueError("matrix argument must be a scalar, a diagonal or a full matrix per location")
        self._mat_type = (self.SCAL, self.DIAG, self.FULL)[rank]
        self._mat = mat
        self._batch_shape = torch.broadcast_shapes(mat.shape[:self._expanded_loc.dim() - 1], self._expanded_loc.shape[:-1])
        self._validate_args = validate_args
        self._torch = None

    # ---- views ----------------------------------------------------------------------------------------------------------
    @property
    def loc(self):
        return self._expanded_loc

    mean = loc

    @property
    def log_norm(self):
        return self._log_norm

    @property
    def event_dim(self):
        return self._event_dim

    @property
    def mat_type(self):
        return self._mat_type

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def torch(self):
        """The wrapped `torch.distributions.MultivariateNormal` (built on first use, reference :165-181)."""
        if self._torch is None:
            dense = self.expand_matrix(self._mat, self._mat_type, self._event_dim)
            self._torch = _TorchMVN(loc=self._expanded_loc, validate_args=self._validate_args, **{self._mat_name: dense})
        return self._torch

    @property
    def covariance_matrix(self):
        return self.torch.covariance_matrix

    def log_prob(self, x):
        """log of the (unnormalised) density: log_norm + log N(x; loc, cov)."""
        return self._log_norm + self.torch.log_prob(x)

    # ---- algebra ----------------------------------------------------------------------------------------------------------
    def __mul__(self, other):
        if not isinstance(other, MultivariateNormal):
            return NotImplemented
        mean, cov, log_c = pairwise_product(self.loc, self.covariance_matrix, other.loc, other.covariance_matrix)
        return MultivariateNormal(loc=mean, scale_tril=torch.linalg.cholesky(cov),
                                  log_norm=log_c + self._log_norm + other._log_norm)


def pairwise_product(mean1, cov1, mean2, cov2):
    """(mean, cov, log constant) of N(x; mean1, cov1) N(x; mean2, cov2), batched over leading dims (closed form above)."""
    S = cov1 + cov2
    chol = torch.linalg.cholesky(S)
    d = (mean2 - mean1)[..., None]
    sol = torch.cholesky_solve(d, chol)                                  # S^-1 (m2 - m1)
    mean = mean1 + (cov1 @ sol)[..., 0]
    cov = cov1 - cov1 @ torch.cholesky_solve(cov1, chol)
    cov = 0.5 * (cov + cov.transpose(-1, -2))
    D = mean1.shape[-1]
    log_det = 2.0 * torch.log(torch.diagonal(chol, dim1=-2, dim2=-1)).sum(-1)
    log_c = -0.5 * (D * math.log(2.0 * math.pi) + log_det + (d * sol).sum((-2, -1)))
    return mean, cov, log_c
