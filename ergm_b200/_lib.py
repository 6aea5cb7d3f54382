"""ctypes binding of the C ABI in include/ergm_b200.h (libergm_b200.so).

The product path has no CPU / eager fallback: if the shared library is missing
or a call returns non-zero, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libergm_b200.so")

ERGM_MAJOR_K, ERGM_MAJOR_MN = 0, 1
EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_ATOMIC, EPI_DROPOUT, EPI_PREACT, EPI_EXACT = 1, 2, 4, 8, 16, 32, 64
DT_BF16, DT_F32 = 0, 1


class ErgmError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("d", ctypes.c_void_p),
        ("bias", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("preact", ctypes.c_void_p),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldd", ctypes.c_int64), ("ldr", ctypes.c_int64),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("a_major", ctypes.c_int32), ("b_major", ctypes.c_int32),
        ("d_dtype", ctypes.c_int32), ("epilogue", ctypes.c_int32), ("split_k", ctypes.c_int32),
        ("block_n", ctypes.c_int32),
        ("dropout_p", ctypes.c_float),
        ("seed", ctypes.c_uint64), ("offset", ctypes.c_uint64),
    ]


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


def _declare(L):
    L.ergm_abi_version.restype = ctypes.c_int
    L.ergm_device_sm_count.restype = ctypes.c_int
    L.ergm_gemm_bf16.restype = ctypes.c_int
    L.ergm_gemm_bf16.argtypes = [ctypes.POINTER(GemmArgs), ctypes.c_void_p]


def check(rc, what):
    if rc != 0:
        if rc > 0:
            raise ErgmError("%s: CUDA error %d" % (what, rc))
        raise ErgmError("%s: %s" % (what, {-1: "invalid argument", -2: "unsupported", -3: "driver entry point / tensor map failure"}.get(rc, rc)))
