#!/usr/bin/env python
"""One scheme step of a given shape on random device operands, repeated: the target of ncu
captures and of kernel micro-timings (runs on the GPU box).

    python tools/one_step.py M N K [--algo tc|stem|simt] [--precision 3xf16|3xtf32|f16] [--reps 3]

M, N, K = number of left-only / right-only / contracted bonds (extent 2 each).  Prints the CUDA-event
time of every launch of the step (tnc_plan_profile) for each repetition."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from artensor_b200 import ContractionPlan, PlanOptions
from artensor_b200 import _native as N

LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXY"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("m", type=int)
    ap.add_argument("n", type=int)
    ap.add_argument("k", type=int)
    ap.add_argument("--algo", default="tc", choices=["tc", "stem", "simt", "skinny"])
    ap.add_argument("--precision", default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--shuffle", action="store_true", help="shuffled mode order (default: A = [m][k], B = [k][n])")
    ap.add_argument("--ka", default=None, help="comma-separated address-bit positions of the contracted bonds in A")
    a = ap.parse_args()
    rng = np.random.RandomState(a.seed)
    lm, ln, lk = LETTERS[:a.m], LETTERS[a.m:a.m + a.n], LETTERS[a.m + a.n:a.m + a.n + a.k]
    la, lb, lo = list(lm + lk), list(lk + ln), list(lm + ln)
    if a.shuffle:
        rng.shuffle(la), rng.shuffle(lb), rng.shuffle(lo)
    if a.ka is not None:
        pos = [int(x) for x in a.ka.split(",")]
        assert len(pos) == a.k and len(set(pos)) == a.k and all(0 <= q < a.m + a.k for q in pos)
        r = a.m + a.k
        slots = [None] * r                          # slots[i] = mode at dim i; address bit = r - 1 - i
        for ch, q in zip(lk, pos):
            slots[r - 1 - q] = ch
        rest = iter(lm)
        la = [ch if ch is not None else next(rest) for ch in slots]
    eq = "".join(la) + "," + "".join(lb) + "->" + "".join(lo)
    extra = {} if a.precision is None else {"tc_precision": a.precision}
    opts = {"tc": PlanOptions(tc_min_flops=0, tc_min_intensity=0, skinny_min_elems=1 << 62, **extra),
            "skinny": PlanOptions(skinny_min_elems=0, skinny_min_n=1, tc_min_flops=float("inf"), **extra),
            "stem": PlanOptions(tc_min_flops=float("inf"), stem_min_elems=0, skinny_min_elems=1 << 62),
            "simt": PlanOptions(tc_min_flops=float("inf"), stem_min_elems=1 << 62, skinny_min_elems=1 << 62)}[a.algo]
    shapes = {0: (2,) * (a.m + a.k), 1: (2,) * (a.k + a.n)}
    plan = ContractionPlan([((0, 1), eq)], shapes, False, options=opts)
    dev = torch.device("cuda:0")
    torch.manual_seed(a.seed)
    blob = torch.randn(plan.leaf_blob_elems, dtype=torch.complex64, device=dev)
    out = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
    ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    flops = 8.0 * 2.0 ** (a.m + a.n + a.k)
    nbytes = 8.0 * (2.0 ** (a.m + a.k) + 2.0 ** (a.k + a.n) + 2.0 ** (a.m + a.n))
    print(f"{eq}  algo={a.algo} options={opts}  workspace {plan.workspace_bytes / 2**30:.2f} GiB", flush=True)
    SL = N.TNC_PROFILE_SLOTS
    for r in range(a.reps):
        out.zero_()
        ms, _ = plan.profile(blob, out, 0, ws, st)          # nothing is sliced: the step is slice-invariant
        i = [j for j, (kind, _) in enumerate(plan.ops[N.TNC_PHASE_ONCE]) if kind == "einsum"][0]
        t = ms[i * SL:(i + 1) * SL]
        print(f"rep {r}: step {t[0]:.3f} ms  launches {[round(x, 3) for x in t[1:]]}  "
              f"{flops / (t[0] * 1e-3) / 1e12:.1f} TF/s  {nbytes / (t[0] * 1e-3) / 1e9:.0f} GB/s (algorithmic)", flush=True)


if __name__ == "__main__":
    main()
