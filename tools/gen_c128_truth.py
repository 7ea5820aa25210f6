#!/usr/bin/env python
"""Adds complex128 truth to fixtures that only hold the reference's complex64 outputs.

Runs in the build container only (imports the reference from /root/reference).  For every named
case the REFERENCE executor (`artensor.contraction.tensor_contraction[_sparse]`) is run in
complex128 with shift-corrected leaf slicing (as tools/gen_cases.py does for the n12 cases):

  per_slice_c128   the fixture's own slice ids, so that the GPU tests can apply
                   relerr(cuda, c128) <= relerr(reference c64, c128) + 1e-5 per amplitude;
  sum_slice_ids /  (with --sum N) the complex128 SUM over the first N slice ids and the
  sum_c128 /       reference's complex64 sum over the same slices: the GPU sum over many slices
  sum_c64          must not drift away from it (coherent accumulator bias).

Usage:  python tools/gen_c128_truth.py n30_sparse64_sc26 n53_m12_sparse1024 --sum 64
"""
import argparse
import os
import sys
import time

import numpy as np

REF = os.environ.get("ARTENSOR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402
from artensor.contraction import tensor_contraction, tensor_contraction_sparse  # noqa: E402  (the reference)

from artensor_b200.cases import load_case, slice_leaves  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def run(case, slice_ids, dtype):
    func = tensor_contraction if case.pattern == "normal" else tensor_contraction_sparse
    sidx = case.slicing_indices()
    leaves = {k: v.to(dtype) for k, v in case.leaves.items()}
    for s in slice_ids:
        t0 = time.time()
        res = func(slice_leaves(leaves, case.slicing_bonds, sidx, int(s)), case.scheme)
        print(f"   slice {s} [{dtype}] {time.time() - t0:.1f}s", flush=True)
        yield res.contiguous().reshape(-1).numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="+")
    ap.add_argument("--sum", type=int, default=0, help="also store the c128 / c64 sums over the first N slice ids")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    for name in a.names:
        case = load_case(os.path.join(GOLD, f"{name}.case.gz"))
        path = os.path.join(GOLD, f"{name}.expected.npz")
        exp = dict(np.load(path))
        if "per_slice_c128" not in exp:
            exp["per_slice_c128"] = np.stack(list(run(case, exp["slice_ids"], torch.complex128)))
            err = np.abs(exp["per_slice_c128"] - exp["per_slice_c64"]).max() / np.sqrt(np.mean(np.abs(exp["per_slice_c128"]) ** 2))
            print(f"[{name}] reference c64 vs c128: max |err| / rms = {err:.3e}")
        if a.sum and case.n_slices >= a.sum and "sum_c128" not in exp:
            ids = np.arange(a.sum, dtype=np.int64)
            exp["sum_slice_ids"] = ids
            exp["sum_c128"] = sum(run(case, ids, torch.complex128))
            acc = None                       # complex64 accumulation, slice after slice (simulation.py:114)
            for r in run(case, ids, torch.complex64):
                acc = r.copy() if acc is None else acc + r
            exp["sum_c64"] = acc
        np.savez_compressed(path, **exp)
        print(f"[{name}] wrote {path}: {sorted(exp)}")


if __name__ == "__main__":
    main()
