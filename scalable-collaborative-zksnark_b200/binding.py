"""ctypes binding of libscz.so -- the same C ABI (include/scz.h) the Rust shim binds."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCZ_LIB") or os.path.join(_HERE, "libscz.so")   # SCZ_LIB: dev override (kernel variants)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "scz.h")

_LIB = None

ERRORS = {
    -1: "SCZ_ERR_BAD_ARG", -2: "SCZ_ERR_LEN_MISMATCH", -3: "SCZ_ERR_NOT_POW2", -4: "SCZ_ERR_LEVEL_OOB",
    -5: "SCZ_ERR_CUDA", -6: "SCZ_ERR_NET", -7: "SCZ_ERR_NOMEM",
}


class SczError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"{ERRORS.get(code, code)}: {text}")
        self.code = code


class NetVTable(C.Structure):
    _COLL = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p)
    _SYNC = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p)
    _ROOTED = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p)
    _fields_ = [("user", C.c_void_p), ("gather", _COLL), ("scatter", _COLL), ("all_gather", _COLL), ("sync", _SYNC),
                ("gather_to", _ROOTED), ("scatter_from", _ROOTED)]


def declared_symbols():
    """every function include/scz.h declares"""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scz_[a-z0-9_]+)\s*\(", text)) - {"scz_net_vtable"})


def lib():
    """Load libscz.so; raises loudly when it has not been built (no fallback exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or make -C scalable-collaborative-zksnark_b200); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.scz_last_error.restype = C.c_char_p
        L.scz_ctx_launch_count.restype = C.c_uint64
        L.scz_ctx_destroy.restype = None
        L.scz_pp_free.restype = None
        if hasattr(L, "scz_srs_free"):
            L.scz_srs_free.restype = None
        _LIB = L
    return _LIB
