"""rbn_b200 -- B200-native inside / outside chart pass behind the `rbnet` Python API.

Module names mirror the reference package (`rbnet.base`, `rbnet.sequential`, `rbnet.pcfg`, `rbnet.util`) so that
`import rbn_b200.pcfg as pcfg` is a drop-in for `import rbnet.pcfg as pcfg` on the inside / marginal-likelihood
path.  All arithmetic of that path runs in hand-written sm_100a CUDA kernels reached through the C ABI in
include/rbn_b200.h (librbn_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import base, sequential, pcfg, util, tmap, data, em  # noqa: F401
from .engine import InsideEngine, inside_logZ, release_workspaces  # noqa: F401
