"""Host side of the fused chart pass: grammar flattening, token packing, the autograd boundary.

`InsideEngine` turns a `SequentialRBN` made of discrete bricks into the flat grammar the C ABI takes
(include/rbn_b200.h), with plain differentiable torch ops, so gradients flow back to the very Parameters the
reference differentiates (and their `project_grad` hooks still fire, /root/reference/rbnet/util.py:139,179):

    W[o_l+b, o_r+c, o_v+a] += S_v[tau, a] * W_tau[b, c, a]     binary transition tau of cell v
                                                               (pcfg.py:312 weighted by pcfg.py:395-398)
    T[q_t+y, o_v+a]        += S_v[tau, a] * T_tau[y, a]        terminal transition tau with term_idx t (pcfg.py:351)
    prior[o_v+a]            = omega_v * pi_v[a]                (pcfg.py:270-273)

The inside pass and the outside pass (expected rule counts) run in librbn_b200.so; `_InsideFn` is the
torch.autograd.Function that joins them.  No arithmetic of the path happens on the host.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .tmap import TMap


def _cuda_device():
    if not torch.cuda.is_available():
        raise _lib.RbnError("rbn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _cache_budget(dev, base_bytes):
    """Bytes of device memory the split-sum cache of one pass may take (RbnDesc.cache_bytes): what is free on the device
    plus what torch's allocator holds unused, minus the rest of the workspace and a safety margin.  RBN_Y_CACHE_GB
    overrides it (0 switches the cache off).  The library clamps the figure to what the batch can use."""
    env = os.environ.get("RBN_Y_CACHE_GB")
    if env is not None:
        return max(0, int(float(env) * 1e9))
    free, total = torch.cuda.mem_get_info(dev)
    reusable = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
    return max(0, int(free + reusable - base_bytes - 4e9 - 0.03 * total))


# One retired workspace per device is kept for the next pass (the split-sum cache makes a workspace most of the device's
# memory: handing that back to torch's allocator and asking for it again every micro-batch invites fragmentation -- a
# 67 MB tensor carved out of the freed block forces a fresh 150 GB cudaMalloc).  Same-stream reuse, like the allocator's.
_WS_POOL = {}


def _pool_take(dev, at_least):
    t = _WS_POOL.get(dev.index)
    if t is not None and t.numel() >= at_least:
        del _WS_POOL[dev.index]
        return t
    return None


def _pool_give(t):
    if t is None or not t.is_cuda or os.environ.get("RBN_WS_POOL") == "0":      # RBN_WS_POOL=0: always go through the allocator
        return
    cur = _WS_POOL.get(t.device.index)
    if cur is None or cur.numel() < t.numel():
        _WS_POOL[t.device.index] = t


def release_workspaces():
    """Give the retained workspaces back to torch's allocator (and, with torch.cuda.empty_cache(), to the device)."""
    _WS_POOL.clear()


class _Pass:
    """One forward pass on the device: inputs, workspace (holds the chart) and outputs.  `keep_split_sums`: let the
    inside pass keep as many per-span split-sums as device memory allows, for the outside pass to read back."""

    def __init__(self, W, T, prior, tokens, lengths, path=_lib.PATH_AUTO, chunk_items=0, keep_split_sums=False,
                 len_host=None):
        lib = _lib.load()
        dev = W.device
        B, L, nt = tokens.shape
        self.desc = _lib.RbnDesc(N=W.shape[0], V=T.shape[0], B=B, L=L, n_tvars=nt, path=path,
                                 chunk_items=chunk_items, flags=0, cache_bytes=0)
        nbytes = lib.rbn_workspace_bytes(ctypes.byref(self.desc))
        if nbytes == 0:
            _lib.check(-1, "rbn_workspace_bytes")
        self._owner = _pool_take(dev, nbytes + 1024)
        if keep_split_sums and lib.rbn_resolve_path(ctypes.byref(self.desc)) == _lib.PATH_TCGEN05:
            # a retained workspace is used as it is (whatever cache fits in it); otherwise size one from the free memory
            base = nbytes
            if self._owner is None:
                self.desc.cache_bytes = _cache_budget(dev, base)
                nbytes = lib.rbn_workspace_bytes(ctypes.byref(self.desc))
            else:
                room = self._owner.numel() - 1024
                self.desc.cache_bytes = 1 << 60                   # everything the batch can use (the library clamps)...
                nbytes = lib.rbn_workspace_bytes(ctypes.byref(self.desc))
                if nbytes > room:                                 # ...or what fits beside the rest (alignment slack: 4 x 1 KB)
                    self.desc.cache_bytes = max(0, room - base - 8192)
                    nbytes = lib.rbn_workspace_bytes(ctypes.byref(self.desc))
        # ragged batch: sort by length (longest first) so that diagonal w only visits the sequences that reach it
        # (rbn_inside_forward_sorted); `order[k]` = original index of the k-th sorted sequence, `pos[b]` its inverse
        self.order = self.pos = self.active = None
        if len_host is None:              # callers that validated the batch hand their host copy down (no second sync)
            len_host = lengths.detach().cpu().numpy().astype(np.int64)
        if B > 1 and int(len_host.min()) != int(len_host.max()):
            order = np.argsort(-len_host, kind="stable")
            self.order = torch.as_tensor(order, device=dev)
            self.pos = torch.as_tensor(np.argsort(order, kind="stable"), device=dev)
            tokens = tokens.index_select(0, self.order).contiguous()
            lengths = lengths.index_select(0, self.order).contiguous()
            srt = len_host[order]
            self.active = (ctypes.c_int32 * (L + 1))(*[int((srt >= w).sum()) for w in range(L + 1)])
        self.W, self.T, self.prior, self.tokens, self.lengths = W, T, prior, tokens, lengths
        if self._owner is None:
            self._owner = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        self.workspace = self._owner[:nbytes + 1024]
        if os.environ.get("RBN_DEBUG_POISON") == "1":     # tests: whatever the kernels read without having written shows as NaN
            self.workspace.fill_(0xFF)
        self.ws_off = (-self.workspace.data_ptr()) % 1024
        self.ws_bytes = nbytes
        self.logZ = torch.empty(B, dtype=torch.float32, device=dev)
        self.zroot = torch.empty(B, dtype=torch.float32, device=dev)
        self.rootexp = torch.empty(B, dtype=torch.float32, device=dev)

    def release(self):
        """Hand the workspace on to the next pass (the chart and the kept split-sums are gone after this)."""
        owner, self._owner, self.workspace = self._owner, None, None
        _pool_give(owner)

    def __del__(self):
        try:
            self.release()
        except Exception:          # interpreter shutdown
            pass

    def _live(self):
        if self.workspace is None:
            raise _lib.RbnError("this pass has handed its workspace on (it does so after its backward pass unless a `holder` "
                                "keeps it for chart export): run the inside pass again, or pass holder=[] to inside_logZ")

    def _ws(self):
        return ctypes.c_void_p(self.workspace.data_ptr() + self.ws_off)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.W.device).cuda_stream)

    def forward(self):
        lib = _lib.load()
        with torch.cuda.device(self.W.device):
            rc = lib.rbn_inside_forward_sorted(ctypes.byref(self.desc), _ptr(self.W), _ptr(self.T), _ptr(self.prior),
                                               _ptr(self.tokens), _ptr(self.lengths), self._ws(), self.ws_bytes,
                                               _ptr(self.logZ), _ptr(self.zroot), _ptr(self.rootexp),
                                               ctypes.cast(self.active, ctypes.c_void_p) if self.active is not None else None,
                                               self._stream())
        _lib.check(rc, "rbn_inside_forward")
        if self.pos is not None:          # hand the outputs back in the caller's order
            self.logZ, self.zroot, self.rootexp = (t.index_select(0, self.pos) for t in (self.logZ, self.zroot, self.rootexp))

    def backward(self, coef):
        self._live()
        lib = _lib.load()
        dev = self.W.device
        gW = torch.empty_like(self.W)
        gT = torch.empty_like(self.T)
        gP = torch.empty_like(self.prior)
        coef = coef.to(device=dev, dtype=torch.float32).contiguous()
        if self.order is not None:
            coef = coef.index_select(0, self.order).contiguous()
        with torch.cuda.device(dev):
            rc = lib.rbn_inside_backward_sorted(ctypes.byref(self.desc), _ptr(self.W), _ptr(self.T), _ptr(self.prior),
                                                _ptr(self.tokens), _ptr(self.lengths), self._ws(), self.ws_bytes,
                                                _ptr(coef), _ptr(gW), _ptr(gT), _ptr(gP),
                                                ctypes.cast(self.active, ctypes.c_void_p) if self.active is not None else None,
                                                self._stream())
        _lib.check(rc, "rbn_inside_backward")
        return gW, gT, gP

    def export_chart(self, seq):
        """float64 [n(n+1)/2, N] linear-space chart of sequence `seq` in TMap row order (device tensor)."""
        self._live()
        lib = _lib.load()
        if self.pos is not None:
            seq = int(self.pos[seq])          # position of the caller's sequence in the length-sorted batch
        n = int(self.lengths[seq])
        out = torch.empty((n * (n + 1) // 2, self.desc.N), dtype=torch.float64, device=self.W.device)
        with torch.cuda.device(self.W.device):
            rc = lib.rbn_export_chart(ctypes.byref(self.desc), _ptr(self.lengths), self._ws(), self.ws_bytes,
                                      int(seq), _ptr(out), self._stream())
        _lib.check(rc, "rbn_export_chart")
        return out


class _InsideFn(torch.autograd.Function):
    """(W, T, prior) -> log Z [B] or Z [B]; backward = the CUDA outside pass (expected rule counts)."""

    @staticmethod
    def forward(ctx, W, T, prior, tokens, lengths, linear, holder, path, chunk_items, len_host=None):
        dev = _cuda_device() if not W.is_cuda else W.device
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        p = _Pass(f32(W), f32(T), f32(prior),
                  tokens.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous(),
                  lengths.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous(), path, chunk_items,
                  keep_split_sums=any(ctx.needs_input_grad[:3]), len_host=len_host)
        p.forward()
        ctx.release = holder is None       # nobody can ask this pass for its chart: its workspace moves on after backward
        if holder is not None:
            holder.append(p)
        ctx.p = p
        ctx.linear = linear
        ctx.in_meta = [(t.dtype, t.device) for t in (W, T, prior)]
        z64, e64 = p.zroot.double(), p.rootexp.double()
        if linear:
            out = torch.ldexp(z64, e64.to(torch.int32))
        else:
            out = torch.where(z64 > 0, torch.log(z64) + e64 * float(np.log(2.0)), torch.full_like(z64, -np.inf))
        ctx.save_for_backward(z64, e64)
        return out.to(W.device)

    @staticmethod
    def backward(ctx, grad_out):
        z64, e64 = ctx.saved_tensors
        g = torch.nan_to_num(grad_out.to(device=z64.device, dtype=torch.float64), nan=0.0, posinf=0.0, neginf=0.0)
        pos = z64 > 0
        post = None
        if ctx.linear:
            # dZ/dtheta = 2^R * (what the kernel accumulates); R = rootexp + log2(zroot) for Z > 0, rootexp otherwise.
            # 2^R leaves fp32 after ~30 tokens, so the device pass runs in the log-Z normalisation: it gets
            # coef_b = g_b 2^(R_b - S) with S = max_b (log2|g_b| + R_b) (|coef| <= 1), and the fp32 sums it returns are
            # multiplied by 2^S in fp64 here -- exact down to what fp64 itself can hold, like the reference's autograd.
            R = e64 + torch.where(pos, torch.log2(torch.where(pos, z64, torch.ones_like(z64))), torch.zeros_like(z64))
            mag = torch.where(g != 0, torch.log2(g.abs().clamp_min(1e-300)) + R, torch.full_like(R, -float("inf")))
            S = mag.max()
            if bool(torch.isfinite(S)):
                S = torch.round(S)
                coef = torch.sign(g) * torch.exp2(mag - S)
                post = S
            else:
                coef = torch.zeros_like(g)
        else:
            coef = torch.where(pos, g, torch.zeros_like(g))     # log Z = -inf has no gradient
        gW, gT, gP = ctx.p.backward(coef)
        if ctx.release:
            ctx.p.release()
        outs = []
        for t, (dt, dv) in zip((gW, gT, gP), ctx.in_meta):
            if post is not None:
                t = torch.ldexp(t.double(), post.to(torch.int32))
            outs.append(t.to(device=dv, dtype=dt))
        return outs[0], outs[1], outs[2], None, None, None, None, None, None, None


TC_MAX_L = 1025       # tensor-core path: spans wider than 129 tokens run in windows of 128 splits (csrc/api.cu kTcMaxL)


def _padded_cardinality(N, batch_spans, L=2):
    """Nonterminal count the tensor-core kernels take (64, 128, 192, 256, 512) when padding N up to it pays: the pad
    symbols get probability zero everywhere, so every result is unchanged.  Small problems stay on the CUDA-core path
    (they are latency-bound, padding would only add work), and so do sequences longer than the tensor-core kernels take."""
    if N % 64 == 0 or N > 512 or N < 24 or batch_spans < 4096 or L > TC_MAX_L:
        return N
    if N > 256:
        return 512 if N >= 400 else N          # 257..399 would waste > 2.2x of the contraction: stay on the CUDA-core path
    return (N + 63) // 64 * 64


def _pad_grammar(W, T, prior, Np):
    N = W.shape[0]
    if Np == N:
        return W, T, prior
    Wp = torch.nn.functional.pad(W, (0, Np - N, 0, Np - N, 0, Np - N))
    return Wp, torch.nn.functional.pad(T, (0, Np - N)), torch.nn.functional.pad(prior, (0, Np - N))


def _host_view(tokens, lengths):
    """(lengths as int64 numpy, largest token) for validation, length sorting and the ragged schedule, with at most ONE
    device -> host copy: inputs that already live on the host (the usual case: a data loader's pinned batch) cost none,
    so a pass then never synchronises the device."""
    B, L = int(tokens.shape[0]), int(tokens.shape[1])
    dev_parts, where = [], []
    if lengths is not None and lengths.is_cuda:
        dev_parts.append(lengths.reshape(-1).to(torch.int64))
        where.append("len")
    if tokens.is_cuda and tokens.numel():
        dev_parts.append(tokens.max().reshape(1).to(torch.int64))
        where.append("tok")
    host = None
    if dev_parts:
        same = all(p.device == dev_parts[0].device for p in dev_parts)
        host = torch.cat(dev_parts).cpu().numpy() if same else np.concatenate([p.cpu().numpy() for p in dev_parts])
    if lengths is None:
        len_host = np.full((B,), L, dtype=np.int64)
    elif "len" in where:
        len_host = host[:B].astype(np.int64)
    else:
        len_host = lengths.detach().reshape(-1).numpy().astype(np.int64)
    if "tok" in where:
        tok_max = int(host[-1])
    else:
        tok_max = int(tokens.max()) if tokens.numel() else -1
    return len_host, tok_max


def inside_logZ(W, T, prior, tokens, lengths=None, linear=False, path=_lib.PATH_AUTO, chunk_items=0, holder=None):
    """Functional entry: flat grammar tensors (any float dtype / device) + integer tokens -> log Z [B] (or Z).
    With `path=AUTO`, a nonterminal count that is not one the tensor-core kernels take is zero-padded up to the next one
    when the batch is large enough for that to pay (differentiable: the pad is sliced off the gradients by autograd)."""
    tokens = torch.as_tensor(tokens)
    if tokens.dim() == 2:
        tokens = tokens[:, :, None]
    if lengths is not None:
        lengths = torch.as_tensor(lengths)
    len_host, tok_max = _host_view(tokens, lengths)
    if lengths is None:
        lengths = torch.from_numpy(len_host.astype(np.int32))
    if int(len_host.min()) < 1 or int(len_host.max()) > tokens.shape[1]:
        raise ValueError("lengths must satisfy 1 <= lengths[b] <= tokens.shape[1]")
    if tok_max >= int(T.shape[0]):
        # the kernels index T[token] unchecked (pcfg.py:351 would raise the same IndexError from torch indexing)
        raise IndexError(f"token {tok_max} out of range for a terminal alphabet of {int(T.shape[0])} rows")
    if path == _lib.PATH_AUTO:
        L = int(tokens.shape[1])
        Np = _padded_cardinality(int(W.shape[0]), int(tokens.shape[0]) * L * (L - 1) // 2, L)
        W, T, prior = _pad_grammar(W, T, prior, Np)
    return _InsideFn.apply(W, T, prior, tokens, lengths, linear, holder, path, chunk_items, len_host)


def viterbi(W, T, prior, tokens, lengths=None):
    """Most probable derivation per sequence (max-semiring variant of the inside pass on the CUDA-core kernels; an
    addition, the reference has no MAP parse).  Returns (logP [B] float64 tensor, trees): a tree is the nested tuple
    (a, i, k, left, right) / (a, i, i + 1) over FLAT nonterminal indices, None for a sequence without a derivation."""
    lib = _lib.load()
    dev = _cuda_device() if not W.is_cuda else W.device
    tokens = torch.as_tensor(tokens)
    if tokens.dim() == 2:
        tokens = tokens[:, :, None]
    B, L, nt = tokens.shape
    if lengths is None:
        lengths = torch.full((B,), L, dtype=torch.int32)
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    Wd, Td, pd = f32(W), f32(T), f32(prior)
    tok = tokens.to(device=dev, dtype=torch.int32).contiguous()
    ln = torch.as_tensor(lengths).to(device=dev, dtype=torch.int32).contiguous()
    N = int(Wd.shape[0])
    desc = _lib.RbnDesc(N=N, V=int(Td.shape[0]), B=B, L=L, n_tvars=nt, path=_lib.PATH_SIMT, chunk_items=0, flags=0)
    nbytes = lib.rbn_workspace_bytes(ctypes.byref(desc))
    if nbytes == 0:
        _lib.check(-1, "rbn_workspace_bytes")
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    C = L * (L + 1) // 2
    logP = torch.empty(B, dtype=torch.float32, device=dev)
    back = torch.empty((B, C, N), dtype=torch.int32, device=dev)
    root = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.rbn_viterbi(ctypes.byref(desc), _ptr(Wd), _ptr(Td), _ptr(pd), _ptr(tok), _ptr(ln),
                             ctypes.c_void_p(ws.data_ptr() + off), nbytes, _ptr(logP), _ptr(back), _ptr(root),
                             ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "rbn_viterbi")
    back_h, root_h, len_h = back.cpu().numpy(), root.cpu().numpy(), ln.cpu().numpy()

    def cell(i, w):
        return (w - 1) * (L + 1) - ((w - 1) * w) // 2 + i

    def build(b, a, i, k):
        if k - i == 1:
            return (a, i, k)
        p = int(back_h[b, cell(i, k - i), a])
        u, bl, cr = p >> 20, (p >> 10) & 1023, p & 1023
        return (a, i, k, build(b, bl, i, i + u), build(b, cr, i + u, k))

    trees = [build(b, int(root_h[b]), 0, int(len_h[b])) if root_h[b] >= 0 else None for b in range(B)]
    return logP.double(), trees


class LazyChart:
    """Stands in for `SequentialRBN._inside_chart` after a fused pass: the per-variable TMap list is exported from
    the device (rbn_export_chart) the first time somebody looks at it (sequential.py:38-40, pcfg.py:30-41)."""

    def __init__(self, device_pass, seq, offsets):
        self._pass, self._seq, self._offsets = device_pass, seq, offsets
        self._charts = None

    def materialise(self):
        if self._charts is None:
            flat = self._pass.export_chart(self._seq).cpu().to(torch.get_default_dtype())
            self._charts = [TMap(flat[:, lo:hi].contiguous()) for lo, hi in zip(self._offsets[:-1], self._offsets[1:])]
        return self._charts


class InsideEngine:
    """Flattens a discrete SequentialRBN and runs it through librbn_b200."""

    def __init__(self, model, path=_lib.PATH_AUTO, chunk_items=0):
        self.model = model
        self.path = path
        self.chunk_items = chunk_items
        self.last_pass = None

    # ---- which models the fused path covers -----------------------------------------------------------------------
    @staticmethod
    def supports(model):
        from . import pcfg
        try:
            cells = list(model.cells())
            ok = all(isinstance(c, pcfg.DiscreteCell) and
                     all(isinstance(t, (pcfg.DiscreteTerminalTransition, pcfg.DiscreteBinaryNonTerminalTransition))
                         for t in c.transitions()) for c in cells)
            return ok and isinstance(model.prior, pcfg.DiscretePrior) and len(cells) > 0
        except (NotImplementedError, AttributeError):
            return False

    # ---- structure (host ints only) ---------------------------------------------------------------------------------
    def _structure(self):
        from . import pcfg
        cells = list(self.model.cells())
        cards = [int(c.variable.cardinality) for c in cells]
        off = np.concatenate([[0], np.cumsum(cards)]).astype(int)
        tsize = {}
        for c in cells:
            for t in c.transitions():
                if hasattr(t, "term_idx"):          # duck-typed so reference (rbnet) bricks flatten too
                    V = int(t.transition_probabilities.p.shape[0])
                    tsize[int(t.term_idx)] = max(tsize.get(int(t.term_idx), 0), V)
        n_tvars = (max(tsize) + 1) if tsize else 1
        toff = np.concatenate([[0], np.cumsum([tsize.get(t, 0) for t in range(n_tvars)])]).astype(int)
        return cells, off, toff

    def flatten(self):
        """(W[N,N,N], T[Vtot,N], prior[N]) as differentiable tensors on the parameters' device and dtype."""
        from . import pcfg
        cells, off, toff = self._structure()
        N = int(off[-1])
        if len(cells) == 1:
            # single-cell grammar (AbstractedPCFG and most uses): W and T are plain products, no block assembly
            trs = list(cells[0].transitions())
            kinds = [hasattr(t, "term_idx") for t in trs]
            if len(trs) == 2 and sorted(kinds) == [False, True] and len(toff) == 2:
                S = cells[0].transition_probabilities.p
                kt, kb = kinds.index(True), kinds.index(False)
                Pt, Pb = trs[kt].transition_probabilities.p, trs[kb].transition_probabilities.p
                if tuple(Pb.shape) == (N, N, N) and Pt.shape[1] == N and int(trs[kb].left_idx) == 0 and int(trs[kb].right_idx) == 0:
                    omega = self.model.prior.structural_distribution.p
                    return (S[kb][None, None, :] * Pb, S[kt][None, :] * Pt,
                            omega[0] * self.model.prior.prior_distributions[0].p)
        ref = self.model.prior.structural_distribution.p
        W = torch.zeros((N, N, N), dtype=ref.dtype, device=ref.device)
        T = torch.zeros((max(int(toff[-1]), 1), N), dtype=ref.dtype, device=ref.device)
        for v, cell in enumerate(cells):
            S = cell.transition_probabilities.p
            pv = slice(int(off[v]), int(off[v + 1]))
            for k, tr in enumerate(cell.transitions()):
                P = tr.transition_probabilities.p
                if hasattr(tr, "term_idx"):
                    t = int(tr.term_idx)
                    rows = slice(int(toff[t]), int(toff[t]) + P.shape[0])
                    if P.shape[1] != off[v + 1] - off[v]:
                        raise ValueError("terminal transition does not match the cardinality of its cell's variable")
                    T[rows, pv] = T[rows, pv] + S[k][None, :] * P
                else:
                    li, ri = int(tr.left_idx), int(tr.right_idx)
                    want = (off[li + 1] - off[li], off[ri + 1] - off[ri], off[v + 1] - off[v])
                    if tuple(P.shape) != tuple(int(x) for x in want):
                        raise ValueError(f"binary transition of shape {tuple(P.shape)} does not match the cardinalities "
                                         f"{tuple(int(x) for x in want)} of (left, right, parent)")
                    sl = (slice(int(off[li]), int(off[li + 1])), slice(int(off[ri]), int(off[ri + 1])), pv)
                    W[sl] = W[sl] + S[k][None, None, :] * P
        omega = self.model.prior.structural_distribution.p
        prior = torch.cat([omega[v] * d.p for v, d in enumerate(self.model.prior.prior_distributions)])
        return W, T, prior

    # ---- tokens -------------------------------------------------------------------------------------------------------
    def pack_sequences(self, sequences):
        """Reference-style sequences (per position: list over terminal variables of int / None, or a bare int /
        length-1 array) -> padded int32 tokens[B, L, n_tvars] (offset into the flat alphabet, -1 = None) + lengths."""
        _, _, toff = self._structure()
        nt = len(toff) - 1
        B = len(sequences)
        L = max(len(s) for s in sequences)
        if min(len(s) for s in sequences) < 1:
            raise ValueError("cannot parse an empty sequence")
        tokens = -np.ones((B, L, nt), dtype=np.int32)
        lengths = np.zeros(B, dtype=np.int32)
        for b, seq in enumerate(sequences):
            lengths[b] = len(seq)
            for i, pos in enumerate(seq):
                vals = list(pos) if isinstance(pos, (list, tuple, np.ndarray)) else [pos]
                for t in range(min(nt, len(vals))):
                    if vals[t] is not None:
                        y = int(vals[t])
                        if not 0 <= y < toff[t + 1] - toff[t]:
                            raise IndexError(f"terminal value {y} out of range for terminal variable {t}")
                        tokens[b, i, t] = y + toff[t]
        return torch.from_numpy(tokens), torch.from_numpy(lengths)

    def pack_tokens(self, tokens, lengths=None):
        """Integer tokens[B, L(, n_tvars)] with per-variable values (-1 = None) -> flat-alphabet tokens."""
        _, _, toff = self._structure()
        tokens = torch.as_tensor(tokens)
        if tokens.dim() == 2:
            tokens = tokens[:, :, None]
        nt = len(toff) - 1
        if tokens.shape[2] != nt:
            raise ValueError(f"tokens carry {tokens.shape[2]} terminal variables per position, the model has {nt}")
        if nt > 1:
            # several terminal variables: variable t owns rows toff[t] .. toff[t+1] of the flat T; a value beyond its own
            # alphabet would silently read another variable's rows, so check per variable before shifting
            size = torch.as_tensor(np.diff(toff), dtype=tokens.dtype, device=tokens.device)
            if bool((tokens >= size[None, None, :]).any()):
                raise IndexError("terminal value out of range for its terminal variable")
            shift = torch.as_tensor(toff[:-1], dtype=tokens.dtype, device=tokens.device)
            tokens = torch.where(tokens >= 0, tokens + shift[None, None, :], tokens)
        # (one terminal variable: the flat alphabet is its own, inside_logZ checks the range)
        return tokens, (None if lengths is None else torch.as_tensor(lengths))

    # ---- entry points --------------------------------------------------------------------------------------------------
    def log_inside(self, tokens, lengths=None, prepacked=False):
        """`tokens[B, L(, n_tvars)]` hold PER-VARIABLE terminal values (-1 = None / padding), as the reference's sequences
        do; they are shifted into the flat alphabet here.  `prepacked=True`: they already index the rows of the flat T
        (what `pack_sequences` returns).  Values are range-checked either way."""
        if not prepacked:
            tokens, lengths = self.pack_tokens(tokens, lengths)
        W, T, prior = self.flatten()
        self.last_pass = None
        return inside_logZ(W, T, prior, tokens, lengths, linear=False, path=self.path, chunk_items=self.chunk_items)

    def viterbi(self, sequences):
        """Best derivation of reference-style sequences under the model: (logP [B], trees over (variable, value) symbols)."""
        tokens, lengths = self.pack_sequences(sequences)
        with torch.no_grad():
            W, T, prior = self.flatten()
        logP, trees = viterbi(W, T, prior, tokens, lengths)
        _, off, _ = self._structure()

        def name(t):
            if t is None:
                return None
            v = int(np.searchsorted(off, t[0], side="right") - 1)
            sym = (v, int(t[0] - off[v]))
            return (sym, t[1], t[2]) if len(t) == 3 else (sym, t[1], t[2], name(t[3]), name(t[4]))
        return logP, [name(t) for t in trees]

    def inside_single(self, sequence):
        """One reference-style sequence -> (Z as 0-dim float64 tensor with grad_fn, LazyChart)."""
        tokens, lengths = self.pack_sequences([sequence])
        W, T, prior = self.flatten()
        holder = []
        Z = inside_logZ(W, T, prior, tokens, lengths, linear=True, path=self.path, chunk_items=self.chunk_items,
                        holder=holder)[0]
        self.last_pass = holder[0]
        _, off, _ = self._structure()
        return Z.to(torch.float64), LazyChart(holder[0], 0, [int(o) for o in off])
