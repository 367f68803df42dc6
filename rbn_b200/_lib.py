"""ctypes binding of librbn_b200.so (the C ABI declared in include/rbn_b200.h).

The library is built in-tree by `make -C rbn_b200/csrc` (or __graft_entry__.build()).  There is no fallback: if
the shared object is missing, `load()` raises and every product entry point fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbn_b200.so")

PATH_AUTO, PATH_SIMT, PATH_TCGEN05 = 0, 1, 2
GAUSS_F32 = 1            # RbnGaussDesc.flags: float32 observations / charts (the _f32 entry points)
ABI_VERSION = 3          # RBN_ABI_VERSION of include/rbn_b200.h this binding was written against
KERNEL_CLASSES = ["emission", "split_sum", "rule_gemm", "gy_gemm", "count_gemm", "child_grad", "grad_prep",
                  "simt_inside", "simt_outside", "other", "rule_finalize"]


class RbnDesc(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("V", ctypes.c_int32), ("B", ctypes.c_int32), ("L", ctypes.c_int32),
                ("n_tvars", ctypes.c_int32), ("path", ctypes.c_int32), ("chunk_items", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("cache_bytes", ctypes.c_int64)]


class GaussDesc(ctypes.Structure):
    _fields_ = [("D", ctypes.c_int32), ("B", ctypes.c_int32), ("L", ctypes.c_int32), ("flags", ctypes.c_int32)]


class RbnError(RuntimeError):
    pass


_lib = None

_P = ctypes.c_void_p
_SIGNATURES = {
    "rbn_abi_version": (ctypes.c_int, []),
    "rbn_last_error": (ctypes.c_char_p, []),
    "rbn_resolve_path": (ctypes.c_int, [ctypes.POINTER(RbnDesc)]),
    "rbn_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(RbnDesc)]),
    "rbn_inside_forward": (ctypes.c_int, [ctypes.POINTER(RbnDesc), _P, _P, _P, _P, _P, _P, ctypes.c_size_t,
                                          _P, _P, _P, _P]),
    "rbn_inside_backward": (ctypes.c_int, [ctypes.POINTER(RbnDesc), _P, _P, _P, _P, _P, _P, ctypes.c_size_t,
                                           _P, _P, _P, _P, _P]),
    "rbn_inside_forward_sorted": (ctypes.c_int, [ctypes.POINTER(RbnDesc), _P, _P, _P, _P, _P, _P, ctypes.c_size_t,
                                                 _P, _P, _P, _P, _P]),
    "rbn_inside_backward_sorted": (ctypes.c_int, [ctypes.POINTER(RbnDesc), _P, _P, _P, _P, _P, _P, ctypes.c_size_t,
                                                  _P, _P, _P, _P, _P, _P]),
    "rbn_viterbi": (ctypes.c_int, [ctypes.POINTER(RbnDesc), _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _P, _P, _P, _P]),
    "rbn_export_chart": (ctypes.c_int, [ctypes.POINTER(RbnDesc), _P, _P, ctypes.c_size_t, ctypes.c_int32, _P, _P]),
    "rbn_lognormalize": (ctypes.c_int, [_P, ctypes.c_int64, ctypes.c_int64, _P, _P]),
    "rbn_lognormalize_f64": (ctypes.c_int, [_P, ctypes.c_int64, ctypes.c_int64, _P, _P]),
    "rbn_gauss_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(GaussDesc)]),
    "rbn_gauss_inside_forward": (ctypes.c_int, [ctypes.POINTER(GaussDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                                ctypes.c_size_t, _P, _P]),
    "rbn_gauss_inside_backward": (ctypes.c_int, [ctypes.POINTER(GaussDesc)] + [_P] * 9 + [ctypes.c_size_t] + [_P] * 9),
    "rbn_gauss_inside_forward_f32": (ctypes.c_int, [ctypes.POINTER(GaussDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                                    ctypes.c_size_t, _P, _P]),
    "rbn_gauss_inside_backward_f32": (ctypes.c_int, [ctypes.POINTER(GaussDesc)] + [_P] * 9 + [ctypes.c_size_t] + [_P] * 9),
    "rbn_profile_enable": (ctypes.c_int, [ctypes.c_int32]),
    "rbn_profile_read": (ctypes.c_int, [_P, _P, ctypes.c_int32, ctypes.c_int32]),
    "rbn_debug_gemm": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, ctypes.c_int32,
                                      ctypes.c_int32, ctypes.c_int32, _P]),
}


def exported_symbols():
    """Names the header declares; tests check the built library exports every one of them."""
    return sorted(_SIGNATURES)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RbnError(f"{LIB_PATH} is missing: build it with `make -C rbn_b200/csrc` "
                           f"(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.rbn_abi_version() != ABI_VERSION:
            raise RbnError(f"{LIB_PATH} has ABI version {lib.rbn_abi_version()}, this binding needs {ABI_VERSION}: "
                           f"rebuild it with `make -C rbn_b200/csrc`")
        _lib = lib
    return _lib


def register(name, restype, argtypes):
    """Used by optional components (Gaussian kernels) to add their entry points to the checked symbol list."""
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(code, what):
    if code != 0:
        msg = load().rbn_last_error().decode("utf-8", "replace")
        raise RbnError(f"{what} failed with code {code}: {msg}")


def profile_enable(on):
    check(load().rbn_profile_enable(1 if on else 0), "rbn_profile_enable")


def profile_read(reset=True):
    """{class: (ms, launches)} since the last reset (synchronises the device)."""
    n = len(KERNEL_CLASSES)
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    check(load().rbn_profile_read(ms, cnt, n, 1 if reset else 0), "rbn_profile_read")
    return {k: (ms[i], cnt[i]) for i, k in enumerate(KERNEL_CLASSES)}
