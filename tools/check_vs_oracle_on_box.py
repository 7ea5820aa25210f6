#!/usr/bin/env python
"""Parity of ONE slice of a large case against the CPU torch oracle, run on the GPU box (its host
has the memory a 2^31..2^32-amplitude slice needs; the build container does not).  The oracle
(oracle/tn_oracle_torch.py) restates the reference executor and is pinned to the reference's
recorded outputs by tests/test_oracle.py; the reference itself does not exist on the box.

    python tools/check_vs_oracle_on_box.py n53_m20_sparse1024_sc32 [slice id]

Prints one JSON line and writes gpurun_out/oracle_check_<case>.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from artensor_b200 import TensorNetworkSimulation, load_case
from oracle import tn_oracle_torch as OT


def main():
    name = sys.argv[1]
    s = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 0
    case = load_case(os.path.join(ROOT, "tests", "golden", f"{name}.case.gz"))
    dev = torch.device("cuda:0")
    from artensor_b200 import PlanOptions
    from artensor_b200 import contraction as C
    extra = {}
    for prec in ("3xtf32",):                     # the other fp32-accurate tensor-core path on the same slice
        sim = TensorNetworkSimulation.from_case(case)
        sim.plan_options = PlanOptions(tc_precision=prec)
        extra[prec] = sim.contraction(device=dev, slice_range=(s, s + 1)).cpu().numpy().reshape(-1).astype(np.complex128)
        del sim
        C.release_workspaces()
        torch.cuda.empty_cache()
    sim = TensorNetworkSimulation.from_case(case)
    plan = sim.plan()
    got = sim.contraction(device=dev, slice_range=(s, s + 1))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    got = sim.contraction(device=dev, slice_range=(s, s + 1))
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1)
    got = got.cpu().numpy().reshape(-1).astype(np.complex128)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    want = OT.contract_slices(case, [s]).reshape(-1).numpy().astype(np.complex128)
    cpu_s = time.perf_counter() - t0
    rms = np.sqrt(np.mean(np.abs(want) ** 2))
    err = np.abs(got - want)
    line = {"case": name, "slice": s, "sliced_bonds": plan.n_sliced, "workspace_gib": plan.workspace_bytes / 2 ** 30,
            "gpu_ms_per_slice": gpu_ms, "cpu_oracle_seconds": cpu_s, "cpu_cores": os.cpu_count(),
            "max_err_over_rms": float(err.max() / rms),
            "within_1e-5_of_max_amp_rms": bool((err <= 1e-5 * np.maximum(np.abs(want), rms)).all()),
            "extrapolated_full_task_seconds": (2.0 ** plan.n_sliced) * gpu_ms * 1e-3}
    if "--c128" in sys.argv:                     # complex128 truth (twice the host memory): who is closer to it?
        t0 = time.perf_counter()
        truth = OT.contract_slices(case, [s], dtype=torch.complex128).reshape(-1).numpy()
        line["cpu_oracle_c128_seconds"] = time.perf_counter() - t0
        trms = np.sqrt(np.mean(np.abs(truth) ** 2))
        for tag, v in [("cuda_default", got), ("cpu_oracle_c64", want)] + [(f"cuda_{k}", x) for k, x in extra.items()]:
            line[f"vs_c128_max_err_over_rms_{tag}"] = float(np.abs(v - truth).max() / trms)
            line[f"vs_c128_rms_err_over_rms_{tag}"] = float(np.sqrt(np.mean(np.abs(v - truth) ** 2)) / trms)
    for prec, other in extra.items():
        line[f"max_err_over_rms_{prec}"] = float(np.abs(other - want).max() / rms)
        line[f"max_diff_over_rms_default_vs_{prec}"] = float(np.abs(other - got).max() / rms)
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    # the oracle's slice, in the fixtures' format (kept as tests/golden/<case>.expected.npz when no reference run exists)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"{name}.oracle_expected.npz"), slice_ids=np.array([s], dtype=np.int64),
                        per_slice_c64=want.astype(np.complex64)[None, :], shape=np.array([len(want)], dtype=np.int64),
                        source=np.array("oracle/tn_oracle_torch.py on the GPU box's host"))
    with open(os.path.join(ROOT, "gpurun_out", f"oracle_check_{name}.json"), "w") as f:
        json.dump(line, f, indent=1)


if __name__ == "__main__":
    main()
