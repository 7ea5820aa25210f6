#!/usr/bin/env python
"""Stand-alone bit-permutation (tnc_permute_bits, the pack kernel in copy mode) on a 2^rank
complex64 tensor: CUDA-event time and achieved HBM GB/s (2 * 8 * 2^rank algorithmic bytes) for a
few permutations of the kind the contraction tree produces (runs on the GPU box)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from artensor_b200 import _native as N


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    lib = N.load()
    dev = torch.device("cuda:0")
    r = a.rank
    src = torch.randn(1 << r, dtype=torch.complex64, device=dev)
    dst = torch.empty_like(src)
    rng = np.random.RandomState(0)
    perms = {
        "identity": list(range(r)),
        "swap_halves": list(range(r // 2, r)) + list(range(r // 2)),          # a matrix transpose
        "rotate_5": [(i + 5) % r for i in range(r)],
        "reverse": list(range(r))[::-1],
        "random_a": [int(x) for x in rng.permutation(r)],
        "random_b": [int(x) for x in rng.permutation(r)],
    }
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, perm in perms.items():
        arr = (C.c_int8 * r)(*perm)
        for _ in range(2):
            N.check(lib.tnc_permute_bits(src.data_ptr(), dst.data_ptr(), r, 1, arr, 8, st))
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(lib.tnc_permute_bits(src.data_ptr(), dst.data_ptr(), r, 1, arr, 8, st))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        gbs = 2 * 8 * (1 << r) / (best * 1e-3) / 1e9
        out[name] = {"ms": best, "GBps": gbs}
        print(f"rank {r} {name:12s} {best:8.3f} ms  {gbs:7.0f} GB/s", flush=True)
    # reference point: a plain device copy of the same tensor
    for _ in range(2):
        dst.copy_(src)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["torch_copy"] = {"ms": best, "GBps": 2 * 8 * (1 << r) / (best * 1e-3) / 1e9}
    print(f"rank {r} torch copy_   {best:8.3f} ms  {out['torch_copy']['GBps']:7.0f} GB/s")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "permute_bench.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
