"""Continuous RBN with multivariate-normal transitions (BASELINE config 4) on the CUDA path.

The reference ships the Gaussian algebra (`rbnet.multivariate_normal`: MultivariateNormal, PairwiseProduct, Product,
ApproximateMixture) but no RBN bricks that use it.  `GaussianRBN` is the composition
SURVEY.md section 8a describes: beta_{i:k}(x) = c N(x; mu, Sigma); terminal p(y|x) = N(y; x, Sigma_T); binary
p(x_l, x_r | x) = N(x_l; x, Sigma_L) N(x_r; x, Sigma_R); splits collapsed by moment matching; prior N(x; mu_P, Sigma_P).
Covariances are parametrised by lower-triangular factors (Sigma = L L^T), the parametrisation
`MultivariateNormal(scale_tril=...)` accepts (/root/reference/rbnet/multivariate_normal.py:147-149).

All chart arithmetic runs in librbn_b200.so (rbn_gauss_inside_forward / rbn_gauss_inside_backward); there is no CPU
fallback.

Chart precision: float64 by default (what the reference's tests use, tests/test_multivariate_normal.py:89-91, and what the
parity tests pin).  `chart_dtype=torch.float32` stores observations, means and covariances in float32 and runs the
per-split algebra in fp32 -- the reference's DEFAULT dtype (multivariate_normal.py:150 builds its tensors with
torch.get_default_dtype()) -- through rbn_gauss_inside_*_f32; log weights, log Z, parameters and parameter gradients
stay float64.  Instantiated for D = 8; other D keep the float64 kernels.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class _GaussFn(torch.autograd.Function):
    """(y, Sigma_T, Sigma_L, Sigma_R, mu_P, Sigma_P, log_struct) -> log Z [B]."""

    @staticmethod
    def forward(ctx, y, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct, lengths, f32):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.RbnError("rbn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = y.device if y.is_cuda else torch.device("cuda", torch.cuda.current_device())
        f64 = lambda t: t.detach().to(device=dev, dtype=torch.float64).contiguous()
        args = [f64(t) for t in (cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct)]
        args.insert(0, y.detach().to(device=dev, dtype=torch.float32 if f32 else torch.float64).contiguous())
        B, L, D = args[0].shape
        lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
        desc = _lib.GaussDesc(D=D, B=B, L=L, flags=_lib.GAUSS_F32 if f32 else 0)
        fwd = lib.rbn_gauss_inside_forward_f32 if f32 else lib.rbn_gauss_inside_forward
        nbytes = lib.rbn_gauss_workspace_bytes(ctypes.byref(desc))
        if nbytes == 0:
            _lib.check(-1, "rbn_gauss_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        logZ = torch.empty(B, dtype=torch.float64, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            rc = fwd(ctypes.byref(desc), *[_ptr(a) for a in args[:1]], _ptr(lengths),
                     *[_ptr(a) for a in args[1:]], _ptr(ws), nbytes, _ptr(logZ), stream)
        _lib.check(rc, "rbn_gauss_inside_forward")
        ctx.state = (desc, args, lengths, ws, nbytes, f32)
        ctx.meta = [(t.dtype, t.device) for t in (y, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct)]
        return logZ.to(y.device)

    @staticmethod
    def backward(ctx, grad_logZ):
        lib = _lib.load()
        desc, args, lengths, ws, nbytes, f32 = ctx.state
        dev = args[0].device
        coef = grad_logZ.to(device=dev, dtype=torch.float64).contiguous()
        grads = [torch.zeros_like(a) for a in args]
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        bwd = lib.rbn_gauss_inside_backward_f32 if f32 else lib.rbn_gauss_inside_backward
        with torch.cuda.device(dev):
            rc = bwd(ctypes.byref(desc), _ptr(args[0]), _ptr(lengths),
                     *[_ptr(a) for a in args[1:]], _ptr(ws), nbytes, _ptr(coef),
                     *[_ptr(g) for g in grads], stream)
        _lib.check(rc, "rbn_gauss_inside_backward")
        outs = [g.to(device=dv, dtype=dt) for g, (dt, dv) in zip(grads, ctx.meta)]
        return (*outs, None, None)


def _use_f32(chart_dtype, dim):
    if chart_dtype in (None, torch.float64):
        return False
    if chart_dtype != torch.float32:
        raise ValueError(f"chart_dtype must be torch.float64 or torch.float32 (got {chart_dtype})")
    return dim == 8          # the fp32 row kernels exist for D = 8; other D run the float64 kernels


def gauss_logZ(y, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct, lengths=None, chart_dtype=None):
    """Functional entry point: log Z [B] (float64) for observations y [B, L, D].  `chart_dtype=torch.float32`: fp32
    charts and per-split algebra (module docstring)."""
    y = torch.as_tensor(y)
    if lengths is None:
        lengths = torch.full((y.shape[0],), y.shape[1], dtype=torch.int32)
    return _GaussFn.apply(y, cov_term, cov_left, cov_right, prior_mean, prior_cov, log_struct, torch.as_tensor(lengths),
                          _use_f32(chart_dtype, y.shape[-1]))


class GaussianRBN(torch.nn.Module):
    """Sequential RBN with one continuous D-dimensional non-terminal variable and Gaussian transitions."""

    def __init__(self, dim, cov_term=None, cov_left=None, cov_right=None, prior_mean=None, prior_cov=None,
                 struct_weights=(0.5, 0.5), chart_dtype=torch.float64):
        super().__init__()
        self.dim = dim
        _use_f32(chart_dtype, dim)              # validates
        self.chart_dtype = chart_dtype
        eye = np.eye(dim)

        def tril(c):
            # a covariance may be given the three ways rbnet.multivariate_normal.MultivariateNormal takes its matrix
            # argument (multivariate_normal.py:150): full [D, D], diagonal [D] (variances) or scalar (one variance), or
            # as a MultivariateNormal whose covariance is taken.  The Parameter keeps the given parametrisation:
            # lower-triangular factor, vector of standard deviations, or one standard deviation.
            from .multivariate_normal import MultivariateNormal
            if isinstance(c, MultivariateNormal):
                c = c.covariance_matrix.detach().cpu().numpy()
            c = eye if c is None else np.asarray(c, dtype=np.float64)
            if c.ndim == 2:
                if c.shape != (dim, dim):
                    raise ValueError(f"covariance of shape {c.shape} does not match dim={dim}")
                return torch.nn.Parameter(torch.linalg.cholesky(torch.as_tensor(c, dtype=torch.float64)))
            if c.ndim == 1 and c.shape[0] != dim:
                raise ValueError(f"Last dimension of tensor has size {c.shape[0]}, cannot expand to diagonal of size {dim}")
            if c.ndim > 1 or (c < 0).any():
                raise ValueError("a covariance is a non-negative scalar, a vector of variances or a [D, D] matrix")
            return torch.nn.Parameter(torch.sqrt(torch.as_tensor(c, dtype=torch.float64)))

        self.term_scale_tril = tril(cov_term)
        self.left_scale_tril = tril(cov_left)
        self.right_scale_tril = tril(cov_right)
        self.prior_scale_tril = tril(prior_cov)
        self.prior_mean = torch.nn.Parameter(torch.zeros(dim, dtype=torch.float64) if prior_mean is None
                                             else torch.as_tensor(np.asarray(prior_mean), dtype=torch.float64))
        sw = torch.as_tensor(np.asarray(struct_weights, dtype=np.float64))
        self.struct_log_p = torch.nn.Parameter(torch.log(sw / sw.sum()))

    def _cov(self, scale):
        if scale.dim() == 2:                       # full: Sigma = L L^T
            low = torch.tril(scale)
            return low @ low.T
        eye = torch.eye(self.dim, dtype=scale.dtype, device=scale.device)
        return eye * (scale * scale)               # diagonal [D] or scalar []: Sigma = diag(s^2)

    def log_inside_batch(self, y, lengths=None):
        """log marginal likelihood [B] of observation sequences y [B, L, D] (float), differentiable."""
        log_struct = self.struct_log_p - torch.logsumexp(self.struct_log_p, dim=0)
        return gauss_logZ(y, self._cov(self.term_scale_tril), self._cov(self.left_scale_tril),
                          self._cov(self.right_scale_tril), self.prior_mean, self._cov(self.prior_scale_tril),
                          log_struct, lengths, chart_dtype=self.chart_dtype)

    def inside(self, sequence):
        """Marginal likelihood (linear space) of one sequence [L, D], the RBN.inside contract (base.py:93-121)."""
        y = torch.as_tensor(np.asarray(sequence), dtype=torch.float64)[None]
        return torch.exp(self.log_inside_batch(y)[0])
