"""ctypes binding of libsegclip_b200.so, generated from include/segclip_b200.h at import time.

The header is the single source of truth for the C ABI: struct layouts and prototypes are parsed from
it, so the Python side cannot drift.  There is deliberately no fallback: if the shared library is
missing or an entry point fails the caller gets an exception -- the product never computes on the CPU
or through PyTorch ops."""
import ctypes as C
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SEGCLIP_B200_LIB selects a variant build of the same library (A/B measurements, trace builds); never a different backend
LIB_PATH = os.environ.get("SEGCLIP_B200_LIB") or os.path.join(_HERE, "lib", "libsegclip_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "segclip_b200.h")

F32, BF16 = 0, 1
ACT_NONE, ACT_QUICKGELU, ACT_GELU_ERF = 0, 1, 2
_DT = {torch.float32: F32, torch.bfloat16: BF16}
TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16}

_SCALARS = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "uint8_t": C.c_uint8, "int64_t": C.c_int64, "float": C.c_float, "int": C.c_int,
            "size_t": C.c_size_t, "long long": C.c_longlong}


class SegclipB200Error(RuntimeError):
    pass


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def parse_header(path=HEADER_PATH):
    """Returns ({struct name: ctypes.Structure}, {function name: (restype, [argtypes])})."""
    src = _strip_comments(open(path).read())
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        body, name = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            decl = decl.replace("const ", "")
            tname, rest = decl.split(" ", 1)
            if tname == "long" and rest.startswith("long "):
                tname, rest = "long long", rest[5:]
            for var in rest.split(","):
                var = var.strip()
                is_ptr = "*" in var or tname.endswith("*")
                var = var.replace("*", "").strip()
                base = tname.replace("*", "")
                if is_ptr:
                    ctype = C.c_void_p
                elif base in _SCALARS:
                    ctype = _SCALARS[base]
                elif base in structs:
                    ctype = structs[base]
                else:
                    raise SegclipB200Error("header parse: unknown type %r in %s" % (tname, name))
                fields.append((var, ctype))
        structs[name] = type(name, (C.Structure,), {"_fields_": fields})
    funcs = {}
    body = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    for m in re.finditer(r"(const\s+char\s*\*|long\s+long|int|void)\s+(sc_\w+)\s*\(([^)]*)\)\s*;", body):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = {"int": C.c_int, "long long": C.c_longlong, "void": None}.get(" ".join(ret.split()), C.c_char_p)
        argtypes = []
        for a in args.split(","):
            a = " ".join(a.replace("const ", "").split())
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(C.c_void_p)
            else:
                t = a.rsplit(" ", 1)[0]
                argtypes.append(_SCALARS[t])
        funcs[name] = (restype, argtypes)
    return structs, funcs


STRUCTS, FUNCS = parse_header()
# integer #defines of the header (SC_NVLS_MAX_BLOCKS, ...)
DEFINES = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(SC_\w+)\s+\(?(-?\d+)\)?\s", open(HEADER_PATH).read())}
SC_NVLS_MAX_BLOCKS = DEFINES["SC_NVLS_MAX_BLOCKS"]
GemmDesc = STRUCTS["sc_gemm_desc"]
LnDesc = STRUCTS["sc_ln_desc"]
LnBwdDesc = STRUCTS["sc_ln_bwd_desc"]
AttnDesc = STRUCTS["sc_attn_desc"]
AttnBwdDesc = STRUCTS["sc_attn_bwd_desc"]
CastItem = STRUCTS["sc_cast_item"]
AssignDesc = STRUCTS["sc_assign_desc"]
AssignBwdDesc = STRUCTS["sc_assign_bwd_desc"]
OptItem = STRUCTS["sc_opt_item"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SegclipB200Error(
                "libsegclip_b200.so not found at %s -- build it with `python -m segclip_b200.build` "
                "(there is no CPU / PyTorch fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in FUNCS.items():
            try:
                fn = getattr(l, name)
            except AttributeError:
                raise SegclipB200Error("libsegclip_b200.so does not export %s (declared in segclip_b200.h)" % name)
            fn.restype = restype
            fn.argtypes = argtypes
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
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().sc_launch_count())


KERNEL_KINDS = {"gemm_tc2": 0, "gemm_tc1": 1, "gemm_simt": 2, "attn_fwd_tc": 3, "attn_bwd_tc": 4, "attn_mma": 5, "attn_generic": 6,
                "gemm_tc2_tail": 7}


def kernel_launches():
    """{kernel family: launches so far in this process} (SC_K_* counters of the library)."""
    l = lib()
    return {k: int(l.sc_kernel_launches(v)) for k, v in KERNEL_KINDS.items()}
