#!/usr/bin/env python
"""Per-operation device-time profile of one slice of a case (runs on the GPU box).

    python tools/gpu_probe.py <case name> [--tc-min-flops X] [--top N]

Prints, for the slice phase, the operations sorted by device time with their algorithmic
flops / bytes and the achieved TFLOP/s and GB/s, and writes gpurun_out/probe_<case>.json."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from artensor_b200 import TensorNetworkSimulation, PlanOptions
from artensor_b200 import _native as N
from artensor_b200 import contraction as C
from artensor_b200.cases import load_case


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--tc-min-flops", type=float, default=None)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--slice", type=int, default=0)
    ap.add_argument("--check", action="store_true", help="compare with the golden per-slice amplitudes")
    ap.add_argument("--precision", default=None, choices=["3xtf32", "3xf16", "f16"])
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-fuse-amax", action="store_true")
    a = ap.parse_args()
    case = load_case(os.path.join(ROOT, "tests", "golden", f"{a.case}.case.gz"))
    sim = TensorNetworkSimulation.from_case(case)
    kw = {}
    if a.tc_min_flops is not None:
        kw["tc_min_flops"] = a.tc_min_flops
    if a.precision is not None:
        kw["tc_precision"] = a.precision
    if a.no_fuse_amax:
        kw["fuse_amax"] = False
    sim.plan_options = PlanOptions(**kw)
    print(f"options: {sim.plan_options}", flush=True)
    plan = sim.plan()
    dev = torch.device("cuda:0")
    blob = plan.pack_leaves(case.leaves, device=dev)
    out = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
    ws = C.get_workspace(dev, plan.workspace_bytes)
    st = torch.cuda.current_stream().cuda_stream
    print(f"{a.case}: workspace {plan.workspace_bytes / 2**30:.2f} GiB, ops once/slice "
          f"{len(plan.ops[0])}/{len(plan.ops[1])}", flush=True)
    plan.profile(blob, out, a.slice, ws, st)          # warm-up (tensor maps, lazy init)
    out.zero_()
    ms_once, ms_slice = plan.profile(blob, out, a.slice, ws, st)
    if a.check:
        exp = np.load(os.path.join(ROOT, "tests", "golden", f"{a.case}.expected.npz"))
        k = int(np.where(exp["slice_ids"] == a.slice)[0][0])
        got = out.cpu().numpy().reshape(-1)
        want = exp["per_slice_c64"][k]
        if "sample_idx" in exp:
            got = got[exp["sample_idx"]]
        rms = np.sqrt(np.mean(np.abs(want) ** 2))
        scale = (np.vdot(want.astype(np.complex128), got.astype(np.complex128)) / np.vdot(want.astype(np.complex128), want.astype(np.complex128)))
        resid = got - scale * want
        print(f"check vs reference slice {a.slice}: max err / rms = {np.abs(got - want).max() / rms:.3e}; best-fit scale - 1 = "
              f"{scale - 1:.3e}; residual after scale max / rms = {np.abs(resid).max() / rms:.3e}", flush=True)
    SL = N.TNC_PROFILE_SLOTS
    rows = []
    for ph, arr in ((0, ms_once), (1, ms_slice)):
        for i, ((kind, rec), s) in enumerate(zip(plan.ops[ph], plan.op_steps[ph])):
            r = {"phase": ph, "op": i, "kind": kind, "ms": arr[i * SL], "launch_ms": arr[i * SL + 1:i * SL + SL]}
            if s is not None:
                r.update(step=s.index, algo=int(rec.algo), m=len(s.m_modes), n=len(s.n_modes), k=len(s.k_modes),
                         rows=s.nb, step_kind=s.kind, flops=s.flops, bytes=s.bytes_c64)
            rows.append(r)
    tot = sum(r["ms"] for r in rows if r["phase"] == 1)
    print(f"slice total {tot:.3f} ms over {len(plan.ops[1])} ops; once phase {sum(r['ms'] for r in rows if r['phase']==0):.3f} ms")
    by_algo = {}
    for r in rows:
        if r["phase"] == 1:
            key = {None: r["kind"], 0: "simt", 1: "tc", 2: "stem", 3: "skinny"}[r.get("algo")]
            by_algo[key] = by_algo.get(key, 0.0) + r["ms"]
    print("by class:", {k: round(v, 3) for k, v in by_algo.items()})
    for r in sorted((r for r in rows if r["phase"] == 1), key=lambda r: -r["ms"])[:a.top]:
        if "flops" in r:
            tf = r["flops"] / (r["ms"] * 1e-3) / 1e12
            gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9
            lm = " ".join(f"{x:.3f}" for x in r["launch_ms"])
            extra = ""
            if r["algo"] == 1 and r["launch_ms"][2] > 0:
                extra = f" gemm-only {r['flops'] / (r['launch_ms'][2] * 1e-3) / 1e12:.1f} TF"
            print(f"  op {r['op']:4d} step {r['step']:4d} {r['step_kind']:7s} algo={r['algo']} m={r['m']:2d} n={r['n']:2d} k={r['k']:2d} "
                  f"rows={r['rows']:5d} {r['ms']:9.3f} ms [{lm}]  {tf:8.2f} TF/s {gb:8.1f} GB/s{extra}")
        else:
            print(f"  op {r['op']:4d} {r['kind']:8s} {r['ms']:9.3f} ms")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"probe_{a.case}{a.tag}.json"), "w") as f:
        json.dump({"case": a.case, "slice_ms": tot, "by_class": by_algo, "ops": rows}, f)


if __name__ == "__main__":
    main()
