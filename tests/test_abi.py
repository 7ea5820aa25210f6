"""The C-ABI library loads, exports every symbol include/tnc_b200.h declares, and the ctypes
struct mirrors match the header's layout (checked by compiling a probe with gcc)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from artensor_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tnc_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tnc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in tnc_b200.h but not exported"
        assert s in N.SYMBOLS, f"{s} has no ctypes prototype in _native.SYMBOLS"
    assert lib.tnc_abi_version() == N.TNC_ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    probe = tmp_path / "probe.c"
    probe.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "tnc_b200.h"\n'
        "int main(){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %d %d\\n\", sizeof(tnc_tensor), sizeof(tnc_einsum),"
        "sizeof(tnc_permute), sizeof(tnc_leaf), sizeof(tnc_accum), offsetof(tnc_einsum, algo),"
        "offsetof(tnc_leaf, keep_pos), offsetof(tnc_einsum, m_a), TNC_MAX_BITS, TNC_MAX_SLICED);return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    vals = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert vals == [C.sizeof(N.TncTensor), C.sizeof(N.TncEinsum), C.sizeof(N.TncPermute), C.sizeof(N.TncLeaf),
                    C.sizeof(N.TncAccum), N.TncEinsum.algo.offset, N.TncLeaf.keep_pos.offset, N.TncEinsum.m_a.offset,
                    N.TNC_MAX_BITS, N.TNC_MAX_SLICED]


def test_plan_builder_validates_without_gpu():
    """Host-side validation paths of the builder API (no CUDA call is made before finalize)."""
    lib = N.load()
    h = C.c_void_p()
    assert lib.tnc_plan_create(N.TNC_C64, 0, C.byref(h)) == 0
    e = N.TncEinsum()
    e.a = N.TncTensor(0, 2, 1)
    e.b = N.TncTensor(1024, 2, 1)
    e.c = N.TncTensor(2048, 2, 1)
    e.nb, e.rows_a, e.rows_b = 1, N.TNC_ROWS_NONE, N.TNC_ROWS_NONE
    e.n_m, e.n_n, e.n_k, e.n_h = 1, 1, 1, 0
    e.m_a, e.m_c, e.n_b, e.n_c, e.k_a, e.k_b = N.bits([1]), N.bits([1]), N.bits([0]), N.bits([0]), N.bits([0]), N.bits([1])
    assert lib.tnc_plan_add_einsum(h, N.TNC_PHASE_SLICE, C.byref(e)) == 0
    e.k_a = N.bits([1])      # collides with m_a
    assert lib.tnc_plan_add_einsum(h, N.TNC_PHASE_SLICE, C.byref(e)) == 1
    assert b"position" in lib.tnc_last_error()
    e.k_a = N.bits([0])
    e.n_k = 2                # counts no longer match the ranks
    assert lib.tnc_plan_add_einsum(h, N.TNC_PHASE_SLICE, C.byref(e)) == 1
    assert lib.tnc_plan_num_ops(h, N.TNC_PHASE_SLICE) == 1
    # executing an unfinalized plan is a state error, not a crash
    assert lib.tnc_plan_execute(h, None, 0, 1, None, None, 0, None) == 5
    lib.tnc_plan_destroy(h)
    assert lib.tnc_plan_create(7, 0, C.byref(h)) == 1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", "/nonexistent/libtnc_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        N.load()


def test_cpu_tensors_are_rejected():
    import torch
    from artensor_b200 import tensor_contraction
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        tensor_contraction([torch.zeros(2, 2, dtype=torch.complex64)] * 2, [((0, 1), "ab,bc->ac")])
