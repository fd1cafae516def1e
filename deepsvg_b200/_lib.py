"""ctypes binding of libdsvg_b200.so (the C ABI declared in include/dsvg_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
Loading the library itself needs no GPU (the CPU test-suite checks the exported symbols).
"""
import ctypes as C
import os
import sys
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libdsvg_b200.so"
_lib = None


class Epilogue(C.Structure):
    """Mirror of `dsvg_epilogue` (include/dsvg_b200.h)."""
    _fields_ = [
        ("acc_scale_dev", C.c_void_p),
        ("bias", C.c_void_p),
        ("scale_cols", C.c_int),
        ("scale", C.c_float),
        ("relu", C.c_int),
        ("drop_p", C.c_float),
        ("drop_site", C.c_uint32),
        ("seed", C.c_uint64),
        ("rowvec", C.c_void_p),
        ("rowvec_ld", C.c_int),
        ("rows_per_group", C.c_int),
        ("mask", C.c_void_p),
        ("mask_lo_off", C.c_size_t),
        ("mask_ld", C.c_int),
        ("mask_scale", C.c_float),
        ("residual", C.c_void_p),
        ("res_ld", C.c_int),
        ("out_f32", C.c_void_p),
        ("out_f32_ld", C.c_int),
        ("out_act", C.c_void_p),
        ("out_lo_off", C.c_size_t),
        ("out_act_ld", C.c_int),
    ]


class OuterProblem(C.Structure):
    """Mirror of `dsvg_outer_problem` (include/dsvg_b200.h)."""
    _fields_ = [
        ("A", C.c_void_p),
        ("lda", C.c_int),
        ("B", C.c_void_p),
        ("ldb", C.c_int),
        ("P", C.c_int),
        ("Q", C.c_int),
        ("alpha", C.c_float),
        ("alpha_dev", C.c_void_p),
        ("C", C.c_void_p),
        ("ldc", C.c_int),
        ("colsum_out", C.c_void_p),
    ]


def lib_path():
    return _LIB_PATH


def load():
    """Returns the loaded library, building it first when nvcc is present and sources changed."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("DSVG_NO_BUILD", "0") != "1":
        try:
            from .csrc.build import build
            build()          # takes a file lock: under torchrun every rank calls this, one of them compiles
        except Exception as e:
            if not _LIB_PATH.exists():
                raise RuntimeError(f"libdsvg_b200.so is missing and could not be built: {e}") from e
            # a library exists but the rebuild failed: it may predate the sources -- the ABI check below is the gate
            sys.stderr.write(f"deepsvg_b200: WARNING: rebuilding libdsvg_b200.so failed ({e}); "
                             "loading the existing (possibly stale) library\n")
    if not _LIB_PATH.exists():
        raise RuntimeError(f"{_LIB_PATH} not found: run `python -m deepsvg_b200.csrc.build` (no CPU fallback exists)")
    lib = C.CDLL(str(_LIB_PATH))
    from ._abi import ABI_VERSION
    try:
        lib.dsvg_abi_version.restype = C.c_int
        have = int(lib.dsvg_abi_version())
    except AttributeError:
        have = -1
    if have != ABI_VERSION:
        raise RuntimeError(f"{_LIB_PATH} exports ABI version {have}, the Python binding expects {ABI_VERSION}: rebuild "
                           "with `python -m deepsvg_b200.csrc.build --force`")
    lib.dsvg_last_error.restype = C.c_char_p
    lib.dsvg_launch_count.restype = C.c_ulonglong
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib):
    from ._abi import SIGNATURES
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => header/library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes


def check(rc, what=""):
    if rc != 0:
        msg = load().dsvg_last_error()
        raise RuntimeError(f"libdsvg_b200 {what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def launch_count():
    return int(load().dsvg_launch_count())
