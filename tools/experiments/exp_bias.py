import ctypes, torch, numpy as np
from rbn_b200 import _lib
lib = _lib.load()
N=256
for K in [1024, 8192, 65536]:
    M=128
    g = torch.Generator(device="cuda").manual_seed(K)
    A = torch.rand((M,K), device="cuda", generator=g).half()
    B = torch.rand((N,K), device="cuda", generator=g).half()
    ref = A.double() @ B.double().t()
    D = torch.zeros((M,N), device="cuda")
    rc = lib.rbn_debug_gemm(M,N,K, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(D.data_ptr()),0,0,SEG, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    rel = ((D.double()-ref)/ref)
    t32 = (A.float() @ B.float().t()).double()
    rel32 = ((t32-ref)/ref)
    h = (A @ B.t()).double()  # cublas fp16 out
    print(f"K={K}: tcgen05 mean rel bias {rel.mean().item():.3e} max|rel| {rel.abs().max().item():.3e} ; torch fp32 matmul mean {rel32.mean().item():.3e} max {rel32.abs().max().item():.3e}")
