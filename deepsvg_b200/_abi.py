"""ctypes signatures of every symbol declared in include/dsvg_b200.h (kept in the same order as the header).

tests/test_abi.py parses the header and checks that this table and the built library export exactly those names.
"""
import ctypes as C

P = C.c_void_p
I = C.c_int
F = C.c_float
Z = C.c_size_t
U32 = C.c_uint32
U64 = C.c_uint64

SIGNATURES = {
    "dsvg_abi_version": (I, []),
    # Y = epilogue(X . W^T)
    "dsvg_linear": (I, [P, Z, I, P, Z, I, I, I, I, P, P]),
    # C += alpha * A^T . B
    "dsvg_outer": (I, [P, Z, I, P, Z, I, I, I, I, F, P, I, P]),
}
