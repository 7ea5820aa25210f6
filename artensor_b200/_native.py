"""ctypes binding of libtnc_b200.so (C ABI: include/tnc_b200.h).

There is NO fallback: if the shared library is missing or the ABI version differs, importing
anything that computes raises.  The struct mirrors below must match the header byte for byte
(`tests/test_abi.py` checks sizes and every exported symbol).
"""
import ctypes as C
import os

TNC_MAX_BITS = 40
TNC_MAX_SLICED = 8
TNC_ABI_VERSION = 7
TNC_WORKSPACE_TAIL_BYTES = 4352
TNC_PROFILE_SLOTS = 4

TNC_C64 = 0
TNC_PHASE_ONCE, TNC_PHASE_SLICE = 0, 1
TNC_ALGO_SIMT, TNC_ALGO_TC, TNC_ALGO_STEM, TNC_ALGO_SKINNY = 0, 1, 2, 3
TNC_ROWS_NONE, TNC_ROWS_IDENTITY = -1, -2
TNC_EINSUM_OUTER_ROWS = 1
TNC_EINSUM_OUTER_PAIRS = 2
TNC_EINSUM_RUN_WITH_READER = 4
TNC_TC_3XTF32, TNC_TC_3XF16, TNC_TC_F16 = 0, 1, 2
TNC_OPT_TC_PRECISION, TNC_OPT_CUDA_GRAPH, TNC_OPT_FUSE_AMAX, TNC_OPT_SLICE_REUSE = 0, 1, 2, 3
TC_PRECISIONS = {"3xtf32": TNC_TC_3XTF32, "3xf16": TNC_TC_3XF16, "f16": TNC_TC_F16}

STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "NOMEM", 4: "UNSUPPORTED", 5: "STATE"}

LIB_NAME = "libtnc_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

Bits = C.c_int8 * TNC_MAX_BITS
Sliced = C.c_int8 * TNC_MAX_SLICED


class TncTensor(C.Structure):
    _fields_ = [("offset", C.c_int64), ("rank", C.c_int32), ("rows", C.c_int32)]


class TncEinsum(C.Structure):
    _fields_ = [
        ("a", TncTensor), ("b", TncTensor), ("c", TncTensor),
        ("nb", C.c_int32), ("rows_a", C.c_int32), ("rows_b", C.c_int32),
        ("n_m", C.c_int32), ("n_n", C.c_int32), ("n_k", C.c_int32), ("n_h", C.c_int32),
        ("m_a", Bits), ("m_c", Bits), ("n_b", Bits), ("n_c", Bits), ("k_a", Bits), ("k_b", Bits),
        ("h_a", Bits), ("h_b", Bits), ("h_c", Bits),
        ("algo", C.c_int32), ("flags", C.c_int32),
        ("scratch_offset", C.c_int64), ("scratch_bytes", C.c_int64),
    ]


class TncPermute(C.Structure):
    _fields_ = [("src", TncTensor), ("dst", TncTensor), ("perm", Bits)]


class TncLeaf(C.Structure):
    _fields_ = [
        ("src_offset", C.c_int64), ("dst", TncTensor), ("src_rank", C.c_int32), ("n_sliced", C.c_int32),
        ("sliced_pos", Sliced), ("sliced_bond", Sliced), ("keep_pos", Bits),
    ]


class TncAccum(C.Structure):
    _fields_ = [("src", TncTensor), ("out_pos", Bits)]


# name -> (restype, argtypes): every symbol include/tnc_b200.h declares
SYMBOLS = {
    "tnc_abi_version": (C.c_int, []),
    "tnc_einsum_tc_scratch_bytes": (C.c_int64, [C.c_int32, C.POINTER(TncEinsum)]),
    "tnc_plan_create": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "tnc_plan_set_option": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    "tnc_plan_add_table": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int32)]),
    "tnc_plan_add_leaves": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(TncLeaf), C.c_int32]),
    "tnc_plan_add_einsum": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(TncEinsum)]),
    "tnc_plan_add_permute": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(TncPermute)]),
    "tnc_plan_add_accum": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(TncAccum)]),
    "tnc_plan_finalize": (C.c_int, [C.c_void_p, C.c_int64]),
    "tnc_plan_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "tnc_plan_num_ops": (C.c_int64, [C.c_void_p, C.c_int32]),
    "tnc_plan_last_launches": (C.c_int64, [C.c_void_p]),
    "tnc_plan_num_fused_amax": (C.c_int64, [C.c_void_p, C.c_int32]),
    "tnc_plan_destroy": (None, [C.c_void_p]),
    "tnc_plan_execute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p]),
    "tnc_plan_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tnc_permute_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int8), C.c_int32,
                                   C.c_void_p]),
    "tnc_last_error": (C.c_char_p, []),
}


class NativeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"tnc_b200: {STATUS.get(status, status)}: {message}")
        self.status = status


_lib = None


def load():
    """Load the shared library (once).  Raises if it is missing: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C artensor_b200/csrc`. artensor_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.tnc_abi_version() != TNC_ABI_VERSION:
        raise ImportError(f"{LIB_NAME}: ABI version {lib.tnc_abi_version()} != expected {TNC_ABI_VERSION}")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().tnc_last_error()
        raise NativeError(status, msg.decode() if msg else "")


def bits(values):
    arr = Bits()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr
