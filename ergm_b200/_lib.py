"""ctypes binding of the C ABI in include/ergm_b200.h (libergm_b200.so).

The product path has no CPU / eager fallback: if the shared library is missing
or a call returns non-zero, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libergm_b200.so")

ERGM_MAJOR_K, ERGM_MAJOR_MN = 0, 1
EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_ATOMIC, EPI_DROPOUT, EPI_PREACT, EPI_EXACT, EPI_GELU_GRAD = 1, 2, 4, 8, 16, 32, 64, 128
DT_BF16, DT_F32 = 0, 1


class ErgmError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("d", ctypes.c_void_p),
        ("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("preact", ctypes.c_void_p),
        ("colsum", ctypes.c_void_p),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldd", ctypes.c_int64), ("ldr", ctypes.c_int64),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("a_major", ctypes.c_int32), ("b_major", ctypes.c_int32),
        ("d_dtype", ctypes.c_int32), ("epilogue", ctypes.c_int32), ("split_k", ctypes.c_int32),
        ("block_n", ctypes.c_int32),
        ("dropout_p", ctypes.c_float),
        ("seed", ctypes.c_uint64), ("offset", ctypes.c_uint64),
        ("dyn_count", ctypes.c_void_p), ("dyn_dim", ctypes.c_int32), ("dyn_hint", ctypes.c_int32),
    ]


class PackDesc(ctypes.Structure):
    """ergm_pack of include/ergm_b200.h: device int32 arrays describing a packed variable-length batch."""
    _fields_ = [("cu_rows", ctypes.c_void_p), ("row_b", ctypes.c_void_p), ("row_t", ctypes.c_void_p),
                ("n_rows", ctypes.c_void_p), ("kv_lens", ctypes.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ErgmError(
                "libergm_b200.so is not built (%s). Run `python -m ergm_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ergm_b200.h")

_CTYPES = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64,
           "uint64_t": ctypes.c_uint64, "float": ctypes.c_float}


def header_prototypes(path=HEADER_PATH):
    """Parses `int ergm_*(...)` prototypes out of include/ergm_b200.h ->
    {name: [ctypes argtypes]} so the binding can never drift from the header."""
    import re
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(ergm_\w+)\s*\(([^)]*)\)\s*;", src):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(ctypes.c_void_p)
                else:
                    base = a.replace("const", "").split()[0]
                    types.append(_CTYPES[base])
        protos[name] = types
    return protos


def _declare(L):
    for name, argtypes in header_prototypes().items():
        fn = getattr(L, name)  # AttributeError here = header declares a symbol the .so lacks
        fn.restype = ctypes.c_int
        fn.argtypes = argtypes


def check(rc, what):
    if rc != 0:
        if rc > 0:
            raise ErgmError("%s: CUDA error %d" % (what, rc))
        raise ErgmError("%s: %s" % (what, {-1: "invalid argument", -2: "unsupported", -3: "driver entry point / tensor map failure"}.get(rc, rc)))
