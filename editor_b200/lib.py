"""ctypes binding of libeditor_b200.so (C ABI in include/editor_b200.h).

The product path has no fallback: if the shared library is missing or a call fails this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeditor_b200.so")

EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_GELU_BWD, EPI_ATOMIC = 0, 1, 2, 3, 4
PREC_BF16, PREC_FP32 = 0, 1


class EdbError(RuntimeError):
    pass


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("A", ctypes.c_void_p), ("lda", ctypes.c_longlong), ("a_mn_major", ctypes.c_int),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_longlong), ("b_mn_major", ctypes.c_int),
        ("D", ctypes.c_void_p), ("ldd", ctypes.c_longlong), ("out_f32", ctypes.c_int),
        ("epilogue", ctypes.c_int),
        ("bias", ctypes.c_void_p),
        ("aux", ctypes.c_void_p), ("ld_aux", ctypes.c_longlong), ("aux_f32", ctypes.c_int),
        ("out2", ctypes.c_void_p), ("ld_out2", ctypes.c_longlong),
        ("alpha", ctypes.c_float),
        ("split_k", ctypes.c_int),
    ]


_lib = None


def load():
    """Load the shared library once; fail loudly when it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EdbError("libeditor_b200.so not built at %s -- run __graft_entry__.build()" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.edb_version.restype = ctypes.c_int
        lib.edb_last_error.restype = ctypes.c_char_p
        lib.edb_gemm_bf16.argtypes = [ctypes.POINTER(GemmDesc), ctypes.c_void_p]
        lib.edb_gemm_bf16.restype = ctypes.c_int
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise EdbError("editor_b200 call failed (%d): %s" % (rc, load().edb_last_error().decode()))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(A, B, D, M, N, K, a_mn=False, b_mn=False, epilogue=EPI_STORE, bias=None, aux=None, out2=None,
         alpha=1.0, split_k=1):
    """D[M,N] = epi(sum_k A(m,k) B(n,k)); A/B bf16 CUDA tensors, D bf16 or fp32 (2-D, row pitch = stride(0))."""
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = A.data_ptr(), A.stride(0), int(a_mn)
    d.B, d.ldb, d.b_mn_major = B.data_ptr(), B.stride(0), int(b_mn)
    d.D, d.ldd, d.out_f32 = D.data_ptr(), D.stride(0), int(D.dtype == torch.float32)
    d.epilogue = epilogue
    d.bias = bias.data_ptr() if bias is not None else None
    if aux is not None:
        d.aux, d.ld_aux, d.aux_f32 = aux.data_ptr(), aux.stride(0), int(aux.dtype == torch.float32)
    if out2 is not None:
        d.out2, d.ld_out2 = out2.data_ptr(), out2.stride(0)
    d.alpha = alpha
    d.split_k = split_k
    check(load().edb_gemm_bf16(ctypes.byref(d), stream_ptr()))
    return D
