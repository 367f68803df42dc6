"""Import the UNMODIFIED reference package `rbnet` from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY -- nothing under rbn_b200/ may import this.  /root/reference does not exist on the
GPU box, so this loader is used only (a) by oracle/make_golden.py to generate the committed fixtures under
tests/golden/ and (b) by `-m "not gpu"` tests that are skipped when /root/reference is absent.

The reference cannot be imported as shipped here: `pytorch_lightning`, `matplotlib` and `triangularmap` are
not installed.  oracle/ref_shims/ holds three stand-ins that perform no arithmetic.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on the GPU box the byte-for-byte copy oracle/vendor_reference.py made at build
# time (oracle/_ref/, git-ignored build output)
REFERENCE_ROOT = "/root/reference" if os.path.isdir("/root/reference/rbnet") else os.path.join(_HERE, "_ref")
_SHIMS = os.path.join(_HERE, "ref_shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rbnet"))


def load_reference():
    """Return the reference's modules as a dict {name: module}; raises if /root/reference is absent."""
    if not reference_available():
        raise RuntimeError("neither /root/reference nor oracle/_ref is present -- use the committed fixtures instead")
    for p in (_SHIMS, REFERENCE_ROOT):          # appended, not prepended: /root/reference has a `tests` package of its
        if p not in sys.path:                   # own that must not shadow this repo's (spawned test workers import ours)
            sys.path.append(p)
    names = ["rbnet.base", "rbnet.util", "rbnet.sequential", "rbnet.pcfg", "rbnet.multivariate_normal"]
    return {n.split(".")[1]: importlib.import_module(n) for n in names}
