#!/usr/bin/env python
"""Rounding-bias calibration of the tcgen05 path (runs on the GPU box): single-step GEMMs of
growing K against an fp64 einsum; prints the best-fit scale of the result (scale - 1 = coherent
bias of the tensor core's round-toward-zero accumulation) and the residual after removing it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_parity import single_step_case, run_single_step

dev = torch.device("cuda:0")
for (m, n, k) in [(9, 6, 2), (9, 6, 3), (9, 6, 4), (9, 6, 5), (9, 6, 6), (9, 6, 7), (8, 7, 8), (8, 7, 10), (7, 7, 12), (7, 6, 14)]:
    scheme, leaves, want = single_step_case(m, n, k, seed=11 + k)
    out = {}
    for algo in ("tc", "stem"):
        got = run_single_step(dev, scheme, leaves, algo).astype(np.complex128)
        scale = np.vdot(want, got) / np.vdot(want, want)
        rms = np.sqrt(np.mean(np.abs(want) ** 2))
        out[algo] = (scale.real - 1, np.abs(got - want).max() / rms, np.abs(got - scale * want).max() / rms)
    print(f"KC={os.environ.get('TNC_TC_KC','4')} k={k:2d} K_real={2 << k:6d}  tc: scale-1 {out['tc'][0]:+.3e} max {out['tc'][1]:.2e} resid {out['tc'][2]:.2e}"
          f"   stem: scale-1 {out['stem'][0]:+.3e} max {out['stem'][1]:.2e}", flush=True)
