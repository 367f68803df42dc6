"""Inside + outside pass of sequences longer than 129 tokens: tensor-core path (split windows) vs CUDA-core path."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rbn_b200 import inside_logZ, _lib, synth

for (N, L, B) in [(256, 256, 16), (256, 192, 32), (64, 256, 64)]:
    W, T, pr = synth.synthetic_grammar(N, 32, seed=0)
    tok, lengths = synth.synthetic_tokens(B, L, 32, seed=1)
    for name, path in (("tcgen05", _lib.PATH_TCGEN05), ("cuda-core", _lib.PATH_SIMT)):
        Wt, Tt, pt = (torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True) for x in (W, T, pr))
        tk, ln = torch.tensor(tok, device="cuda"), torch.tensor(lengths, device="cuda")
        reps = 3 if name == "tcgen05" else 1
        for it in range(1 + reps):
            if it == 1:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            for p in (Wt, Tt, pt):
                p.grad = None
            lz = inside_logZ(Wt, Tt, pt, tk, ln, path=path)
            lz.sum().backward()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"N={N} L={L} B={B} {name}: {dt * 1e3:.1f} ms per inside+outside pass = {B / dt:.1f} seq/s, logZ mean {float(lz.mean()):.4f}", flush=True)
