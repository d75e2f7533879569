"""ctypes binding of libsegclip_b200.so (the C ABI declared in include/segclip_b200.h).

There is deliberately no fallback: if the shared library is missing or an entry point fails the
caller gets an exception -- the product never computes on the CPU or through PyTorch ops."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsegclip_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_QUICKGELU, ACT_GELU_ERF = 0, 1, 2
_DT = {torch.float32: F32, torch.bfloat16: BF16}


class SegclipB200Error(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("in_dtype", C.c_int32),
        ("trans_a", C.c_int32), ("trans_b", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64),
        ("alpha", C.c_float),
        ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("ld_rowbias", C.c_int64),
        ("rowbias_idx", C.c_void_p), ("rowbias_mod", C.c_int32), ("act", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("c_dtype", C.c_int32),
        ("C2", C.c_void_p), ("c2_dtype", C.c_int32),
        ("accumulate", C.c_int32), ("split_k", C.c_int32), ("force_simt", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SegclipB200Error(
                "libsegclip_b200.so not found at %s -- build it with `python -m segclip_b200.build` "
                "(there is no CPU / PyTorch fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.sc_last_error.restype = C.c_char_p
        l.sc_launch_count.restype = C.c_longlong
        if l.sc_abi_version() != 1:
            raise SegclipB200Error("libsegclip_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise SegclipB200Error("%s failed (%d): %s" % (what, rc, lib().sc_last_error().decode()))


def dt(t):
    return _DT[t.dtype]


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().sc_launch_count())
