"""Debug: tensor-core path vs CUDA-core path on sequences longer than 129 tokens (log Z and gradients)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rbn_b200 import inside_logZ, _lib, synth

N, V = int(os.environ.get("N", 64)), 8
W, T, pr = synth.synthetic_grammar(N, V, seed=11)
for L in [int(x) for x in os.environ.get("LS", "129,130,131,140,160,200,257,258,300").split(",")]:
    tok, lengths = synth.synthetic_tokens(2, L, V, seed=12)
    out = {}
    for name, path in (("simt", _lib.PATH_SIMT), ("tc", _lib.PATH_TCGEN05)):
        Wt, Tt, pt = (torch.tensor(x, device="cuda", requires_grad=True) for x in (W, T, pr))
        lz = inside_logZ(Wt, Tt, pt, torch.tensor(tok), torch.tensor(lengths), path=path)
        lz.sum().backward()
        out[name] = (lz.detach().cpu().numpy(), Wt.grad.cpu().numpy(), Tt.grad.cpu().numpy())
    a, b = out["simt"], out["tc"]
    print(f"N={N} L={L}: logZ simt {a[0]} tc {b[0]} diff {b[0] - a[0]}  gW max-rel {np.abs(a[1] - b[1]).max() / np.abs(a[1]).max():.2e} "
          f"sum gW*W simt {float((a[1] * W).sum()):.3f} tc {float((b[1] * W).sum()):.3f} (expect {2 * (L - 1)})  gT max-rel {np.abs(a[2] - b[2]).max() / np.abs(a[2]).max():.2e}", flush=True)
