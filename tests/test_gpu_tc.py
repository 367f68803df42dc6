"""Tensor-core (tcgen05) path: the GEMM kernel template against torch, and the full chart pass against the fp64
oracle and against the CUDA-core path."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGZ_RTOL = 1e-4
GRAD_RTOL = 1e-3


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("M,K,seg", [(128, 64, 0), (304, 192, 0), (1000, 1024, 0), (200, 4096 * 3 + 128, 1), (128, 65536, 1)])
def test_tcgen05_gemm_template_vs_torch(a_mn, b_mn, M, K, seg):
    from rbn_b200 import _lib
    lib = _lib.load()
    N = 256
    g = torch.Generator(device="cuda").manual_seed(M * 7 + K + a_mn * 2 + b_mn)
    A = torch.rand((M, K), device="cuda", generator=g).half()
    Bm = torch.rand((N, K), device="cuda", generator=g).half()
    ref = (A.double() @ Bm.double().t()).float()
    Ain = A.t().contiguous() if a_mn else A.contiguous()
    Bin = Bm.t().contiguous() if b_mn else Bm.contiguous()
    D = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.rbn_debug_gemm(M, N, K, ctypes.c_void_p(Ain.data_ptr()), ctypes.c_void_p(Bin.data_ptr()),
                            ctypes.c_void_p(D.data_ptr()), a_mn, b_mn, seg,
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "rbn_debug_gemm")
    torch.cuda.synchronize()
    err = float((D - ref).abs().max() / ref.abs().max())
    bias = float(((D - ref) / ref).mean())
    assert err < 4e-5 and abs(bias) < 3e-5, (a_mn, b_mn, M, K, err, bias)


@pytest.mark.parametrize("mn", [0, 1])
@pytest.mark.parametrize("M,K,seg", [(256, 64, 2), (300 + 4, 192, 2), (1000, 1024, 2), (520, 4096 * 3 + 128, 3), (256, 65536, 3)])
def test_tcgen05_two_cta_gemm_vs_torch(mn, M, K, seg):
    """cta_group::2 variant of the GEMM template (CTA pairs, 256-row tiles, B split across the pair)."""
    from rbn_b200 import _lib
    lib = _lib.load()
    N = 256
    g = torch.Generator(device="cuda").manual_seed(M + K + mn)
    A = torch.rand((M, K), device="cuda", generator=g).half()
    Bm = torch.rand((N, K), device="cuda", generator=g).half()
    ref = (A.double() @ Bm.double().t()).float()
    Ain = A.t().contiguous() if mn else A.contiguous()
    Bin = Bm.t().contiguous() if mn else Bm.contiguous()
    D = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.rbn_debug_gemm(M, N, K, ctypes.c_void_p(Ain.data_ptr()), ctypes.c_void_p(Bin.data_ptr()),
                            ctypes.c_void_p(D.data_ptr()), mn, mn, seg,
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "rbn_debug_gemm")
    torch.cuda.synchronize()
    err = float((D - ref).abs().max() / ref.abs().max())
    bias = float(((D - ref) / ref).mean())
    assert err < 4e-5 and abs(bias) < 3e-5, (mn, M, K, err, bias)


def _run(N, V, L, B, ragged, seed, path, chunk=0):
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ
    W, T, pr = O.synthetic_grammar(N, V, seed=seed)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=seed + 1, ragged=ragged)
    coef = np.random.default_rng(seed + 2).uniform(0.5, 1.5, B)
    Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=path, chunk_items=chunk)
    (lz * torch.tensor(coef, device="cuda")).sum().backward()
    return (W, T, pr, tok, lengths, coef), lz.detach().cpu().numpy(), Wt.grad.cpu().numpy(), Tt.grad.cpu().numpy(), pt.grad.cpu().numpy()


# N = 64 runs the split-sum kernel in pair mode (two spans per MMA): odd span counts per launch (B = 3, 5; chunk 127),
# spans beyond their sequence paired with live ones (ragged), and L = 80, whose diagonals wider than 65 fall back to
# one span per item inside the same pass
@pytest.mark.parametrize("N,V,L,B,ragged,chunk", [(64, 16, 8, 3, False, 0), (64, 16, 12, 5, True, 128), (64, 16, 13, 5, True, 127),
                                                  (64, 8, 80, 3, True, 0), (128, 8, 9, 3, True, 0),
                                                  (192, 8, 6, 2, False, 0), (256, 32, 7, 2, True, 0), (256, 32, 20, 2, False, 0),
                                                  (512, 8, 6, 2, True, 0), (512, 16, 19, 3, True, 0)])
def test_tc_path_vs_fp64_oracle(N, V, L, B, ragged, chunk):
    from oracle import inside_oracle as O
    from rbn_b200 import _lib
    inp, lz, gW, gT, gP = _run(N, V, L, B, ragged, seed=100 + N + L, path=_lib.PATH_TCGEN05, chunk=chunk)
    W, T, pr, tok, lengths, coef = inp
    lz_ref, gW_ref, gT_ref, gP_ref = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths, grad_logZ=coef)
    assert np.abs(lz - lz_ref.numpy()).max() <= LOGZ_RTOL * np.abs(lz_ref.numpy()).max(), (lz, lz_ref)
    errs = (_rel(gW, gW_ref.numpy()), _rel(gT, gT_ref.numpy()), _rel(gP, gP_ref.numpy()))
    assert max(errs) <= GRAD_RTOL, errs


def test_tc_path_vs_cuda_core_path_long():
    """L = 64 (beyond what linear space can hold): tensor-core path against the fp32 CUDA-core path + count identities."""
    from rbn_b200 import _lib
    N, V, L, B = 64, 32, 64, 4
    _, lz_s, gW_s, gT_s, gP_s = _run(N, V, L, B, True, seed=7, path=_lib.PATH_SIMT)
    inp, lz_t, gW_t, gT_t, gP_t = _run(N, V, L, B, True, seed=7, path=_lib.PATH_TCGEN05, chunk=256)
    assert np.abs(lz_t - lz_s).max() <= LOGZ_RTOL * np.abs(lz_s).max()
    assert max(_rel(gW_t, gW_s), _rel(gT_t, gT_s), _rel(gP_t, gP_s)) <= GRAD_RTOL
    W, T, pr, tok, lengths, coef = inp
    assert abs((gW_t * W).sum() - (coef * (lengths - 1)).sum()) <= 1e-3 * (coef * lengths).sum()
    assert abs((gT_t * T).sum() - (coef * lengths).sum()) <= 1e-3 * (coef * lengths).sum()


def test_full_size_shape_properties_n256_l128():
    """BASELINE configs[2] shape (N = 256, L = 128) through size-independent properties: expected-count identities
    (binary L-1, terminal L, prior 1 per sequence), permutation invariance over the batch, and agreement of the
    shared prefix lengths with independently run shorter batches."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    N, V, L = 256, 32, 128
    W, T, pr = O.synthetic_grammar(N, V, seed=0)
    tok, _ = O.synthetic_tokens(6, L, V, seed=1)
    lengths = np.array([128, 128, 100, 77, 64, 5])
    for b in range(6):
        tok[b, lengths[b]:] = -1
    Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_TCGEN05)
    assert torch.isfinite(lz).all()
    lz.sum().backward()
    gW, gT, gP = Wt.grad.cpu().numpy(), Tt.grad.cpu().numpy(), pt.grad.cpu().numpy()
    tot = float(lengths.sum())
    assert abs((gW * W).sum() - (lengths - 1).sum()) <= 1e-3 * tot
    assert abs((gT * T).sum() - tot) <= 1e-3 * tot
    assert abs((gP * pr).sum() - 6) <= 1e-3 * 6
    # permutation of the batch + a different chunking must not change any sequence's result
    perm = np.array([3, 0, 5, 2, 1, 4])
    lz2 = inside_logZ(torch.tensor(W, device="cuda"), torch.tensor(T, device="cuda"), torch.tensor(pr, device="cuda"),
                      torch.tensor(tok[perm]), torch.tensor(lengths[perm]), path=_lib.PATH_TCGEN05, chunk_items=256)
    # (the chunking changes how the rule GEMM's K range is cut over the CTA pairs, i.e. the length of the tcgen05
    # accumulation chains and the order of the partial sums: results agree to the accumulation error, not bit for bit)
    np.testing.assert_allclose(lz2.cpu().numpy(), lz.detach().cpu().numpy()[perm], rtol=1e-5)


def test_n256_vs_oracle_medium_length():
    from oracle import inside_oracle as O
    from rbn_b200 import _lib
    inp, lz, gW, gT, gP = _run(256, 32, 40, 2, True, seed=21, path=_lib.PATH_TCGEN05)
    W, T, pr, tok, lengths, coef = inp
    lz_ref, gW_ref, gT_ref, gP_ref = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths, grad_logZ=coef)
    assert np.abs(lz - lz_ref.numpy()).max() <= LOGZ_RTOL * np.abs(lz_ref.numpy()).max()
    errs = (_rel(gW, gW_ref.numpy()), _rel(gT, gT_ref.numpy()), _rel(gP, gP_ref.numpy()))
    assert max(errs) <= GRAD_RTOL, errs


def test_n512_count_identities_and_chunk_invariance():
    """N = 512 (BASELINE configs[4] grammar size; sub-blocked split-sum, n-tiled rule / count GEMMs): expected-count
    identities at the full length L = 128 (two split slabs) and independence of the chunking."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    N, V, L = 512, 32, 128                       # the full BASELINE configs[4] shape
    W, T, pr = O.synthetic_grammar(N, V, seed=3)
    tok, _ = O.synthetic_tokens(3, L, V, seed=4)
    lengths = np.array([128, 72, 9])
    for b in range(3):
        tok[b, lengths[b]:] = -1
    Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
    lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_TCGEN05)
    assert torch.isfinite(lz).all()
    lz.sum().backward()
    gW, gT, gP = Wt.grad.cpu().numpy(), Tt.grad.cpu().numpy(), pt.grad.cpu().numpy()
    tot = float(lengths.sum())
    assert abs((gW.astype(np.float64) * W).sum() - (lengths - 1).sum()) <= 1e-3 * tot
    assert abs((gT.astype(np.float64) * T).sum() - tot) <= 1e-3 * tot
    assert abs((gP.astype(np.float64) * pr).sum() - 3) <= 1e-3 * 3
    lz2 = inside_logZ(torch.tensor(W, device="cuda"), torch.tensor(T, device="cuda"), torch.tensor(pr, device="cuda"),
                      torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_TCGEN05, chunk_items=128)
    np.testing.assert_allclose(lz2.cpu().numpy(), lz.detach().cpu().numpy(), rtol=1e-5)


@pytest.mark.parametrize("mode,shape", [("1", (128, 8, 14, 9)), ("2", (128, 8, 14, 9)), ("2", (256, 8, 8, 5)), ("3", (128, 8, 14, 9))])
def test_two_stream_schedule_matches_oracle(mode, shape):
    """RBN_TC_OVERLAP = 1 (two launch streams, SM partitions, sequence groups, slot ring), 2 (co-resident half-SM
    kernels: single-accumulator split-sum, rule GEMM without in-TMEM segmentation) and 3 (tail filling), each in a
    fresh process (the schedule is read once per process)."""
    import os
    import subprocess
    import sys
    code = r"""
import numpy as np, torch
from oracle import inside_oracle as O
from rbn_b200 import inside_logZ, _lib
N, V, L, B = SHAPE
W, T, pr = O.synthetic_grammar(N, V, seed=5)
tok, lengths = O.synthetic_tokens(B, L, V, seed=6, ragged=True)
Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_TCGEN05, chunk_items=128)
lz.sum().backward()
ref, gW, gT, gP = O.flat_inside_with_grads(W, T, pr, tok[:, :, None], lengths)
assert np.abs(lz.detach().cpu().numpy() - ref.numpy()).max() <= 1e-4 * np.abs(ref.numpy()).max()
for a, b in ((Wt.grad, gW), (Tt.grad, gT), (pt.grad, gP)):
    assert float((a.cpu() - b).abs().max() / b.abs().max()) <= 1e-3
print("two-stream OK")
""".replace("SHAPE", repr(shape))
    env = dict(os.environ, RBN_TC_OVERLAP=mode, RBN_TC_GROUPS="2", RBN_TC_SLOTS="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "two-stream OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]



def test_full_size_config2_shape_inside_only():
    """BASELINE configs[1] shape (N = 64, L = 64, inside only) on a slice of the batch: log Z against the fp32 CUDA-core
    path for every sequence, and invariance to where a sequence sits in the batch."""
    from oracle import inside_oracle as O
    from rbn_b200 import inside_logZ, _lib
    N, V, L, B = 64, 32, 64, 48
    W, T, pr = O.synthetic_grammar(N, V, seed=0)
    tok, lengths = O.synthetic_tokens(B, L, V, seed=1)
    dev = [torch.tensor(x, device="cuda") for x in (W, T, pr)]
    lz_t = inside_logZ(*dev, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_TCGEN05).cpu().numpy()
    lz_s = inside_logZ(*dev, torch.tensor(tok), torch.tensor(lengths), path=_lib.PATH_SIMT).cpu().numpy()
    assert np.isfinite(lz_t).all()
    np.testing.assert_allclose(lz_t, lz_s, rtol=LOGZ_RTOL)
    perm = np.random.default_rng(5).permutation(B)
    lz_p = inside_logZ(*dev, torch.tensor(tok[perm]), torch.tensor(lengths[perm]), path=_lib.PATH_TCGEN05, chunk_items=512).cpu().numpy()
    np.testing.assert_allclose(lz_p, lz_t[perm], rtol=1e-5)
