#!/usr/bin/env python
"""B200-aware order search: the reference's simulated annealing, unchanged, driven
with a byte-aware memory weight and a sweep of sc_target, every resulting tree scored with this
executor's own cost model.

Runs in the build container only (imports the reference from /root/reference; minutes per run).
The reference scores a tree with log10(alpha * 10^mc + 10^tc) (order_finder.py:11-16): tc = complex
multiply-adds, mc = amplitudes moved.  On a B200 a complex MAC costs 8 flops / ~600 useful TFLOP/s
and an amplitude moved 8 bytes / ~6.5 TB/s, so one amplitude of traffic is worth ~90 MACs: alpha ~ 96
is the machine's own balance (the fixtures of round 1 were searched with the default 32 and the
cheapest annealing schedule).  sc_target bounds the largest intermediate at 2^sc_target amplitudes;
180 GB of HBM hold sc_target 32 (32 GiB tensors + packed panels), which divides the slice count by 4.

For every (sc_target, alpha, seed) the tree is compiled (artensor_b200.scheme), lowered by the
planner WITHOUT the native library, and priced step by step:

    t(step) = max(flops * products / P_tensor, bytes / BW_class) (+ pack traffic of GEMM steps)

with the rates measured in round 2 (profiles/r02_*): the full task is 2^S slices of that.

    python tools/order_search_sweep.py --circuit n53_m20 --sc 30 31 32 --alpha 32 96 --trials 8 --iters 6
"""
import argparse
import json
import os
import sys
import time

REF = os.environ.get("ARTENSOR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402

from artensor_b200 import TensorNetworkSimulation  # noqa: E402
from artensor_b200 import _native as N  # noqa: E402
from artensor_b200.backend import ContractionPlan, tc_uses_3m  # noqa: E402
from artensor_b200.cases import save_case  # noqa: E402

# measured on B200 inside power-capped slices (round 2): issued fp16 tensor flops, HBM-bound classes
RATES = {"tensor_issued": 1.25e15, "pack": 4.5e12, "skinny": 4.0e12, "stem": 3.3e12, "generic_step_s": 2.6e-6}


def price(plan):
    """Predicted seconds of one slice (executed steps only) and of the ONCE phase."""
    t = {N.TNC_PHASE_ONCE: 0.0, N.TNC_PHASE_SLICE: 0.0}
    split = {"tc": 0.0, "pack": 0.0, "skinny": 0.0, "stem": 0.0, "generic": 0.0}
    for ph in t:
        for (kind, rec), st in zip(plan.ops[ph], plan.op_steps[ph]):
            if kind != "einsum":
                continue
            if rec.algo == N.TNC_ALGO_TC:
                prod = 2.25 if tc_uses_3m(st, "3xf16") else 3.0
                gemm = max(st.flops * prod / RATES["tensor_issued"], st.bytes_c64 / 6.0e12)
                pack = (st.a.numel + st.b.numel) * (8 + 8 + 12) / RATES["pack"]       # amax read + pack read + panel write
                dt, key = gemm + pack, "tc"
                if ph == N.TNC_PHASE_SLICE:
                    split["pack"] += pack
                    dt_gemm = gemm
            elif rec.algo == N.TNC_ALGO_SKINNY:
                dt, key = st.bytes_c64 / RATES["skinny"], "skinny"
            elif rec.algo == N.TNC_ALGO_STEM:
                dt, key = st.bytes_c64 / RATES["stem"], "stem"
            else:
                dt, key = RATES["generic_step_s"], "generic"
            t[ph] += dt
            if ph == N.TNC_PHASE_SLICE:
                split[key] += (dt_gemm if key == "tc" else dt)
    return t[N.TNC_PHASE_SLICE], t[N.TNC_PHASE_ONCE], split


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--circuit", default="n53_m20", choices=["n53_m20", "n53_m12", "n30"])
    ap.add_argument("--sc", type=int, nargs="+", default=[30, 31, 32])
    ap.add_argument("--alpha", type=float, nargs="+", default=[32.0, 96.0])
    ap.add_argument("--trials", type=int, default=8)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--seeds", type=int, nargs="+", default=[0])
    ap.add_argument("--out", default="/tmp/tnc_sweep")
    a = ap.parse_args()
    import gen_cases as G
    os.makedirs(a.out, exist_ok=True)
    if a.circuit == "n30":
        qsim, bitstrings = G.n30_qsim(), G.google_amplitudes(10000)[0]
    else:
        qsim, bitstrings = G.n53_qsim(20 if a.circuit == "n53_m20" else 12), G.correlated_bitstrings(53, 10, 0)
    results = []
    for sc in a.sc:
        for alpha in a.alpha:
            for seed in a.seeds:
                t0 = time.time()
                sim = TensorNetworkSimulation.from_circuit_file(qsim, bitstrings)
                sim.prepare_contraction(sc_target=sc, trials=a.trials, iters=a.iters, slicing_repeat=1, start_seed=seed, alpha=alpha)
                G.validate_scheme(sim.scheme, sim.pattern)
                plan = ContractionPlan(sim.scheme, {i: tuple(t.shape) for i, t in sim.tensors.items()}, sim.pattern == "sparse",
                                       slicing_bonds=sim.slicing_bonds, slicing_indices=sim.slicing_indices, build_native=False)
                t_slice, t_once, split = price(plan)
                # the same tree under cross-slice reuse: sliced bonds re-ordered, KEEP region
                from artensor_b200 import PlanOptions
                order = plan.reuse_bond_order()
                reuse_model = plan.reuse_summary(order)
                bonds = [sim.slicing_bonds[i] for i in order]
                from artensor_b200.simulation import slicing_dims
                rplan = ContractionPlan(sim.scheme, {i: tuple(t.shape) for i, t in sim.tensors.items()}, sim.pattern == "sparse",
                                        slicing_bonds=bonds, slicing_indices=slicing_dims(sim.tensors, sim.tensor_bonds, bonds),
                                        options=PlanOptions(slice_reuse=True), build_native=False)
                work = plan.work_summary()
                S = len(sim.slicing_bonds)
                name = f"{a.circuit}_sc{sc}_a{int(alpha)}_s{seed}"
                r = {"name": name, "sc_target": sc, "alpha": alpha, "seed": seed, "trials": a.trials, "iters": a.iters,
                     "sliced_bonds": S, "steps": work["steps"], "flops_per_slice": work["exec_flops_per_slice"],
                     "bytes_per_slice": work["exec_bytes_per_slice"], "workspace_gib": plan.workspace_bytes / 2 ** 30,
                     "predicted_ms_per_slice": 1e3 * t_slice, "predicted_split_ms": {k: 1e3 * v for k, v in split.items()},
                     "predicted_full_task_seconds": (2.0 ** S) * t_slice, "log2_full_task_seconds": S + float(np.log2(t_slice)),
                     "reuse_modelled_ms_per_slice": 1e3 * reuse_model["amortised_s"], "reuse_workspace_gib": rplan.workspace_bytes / 2 ** 30,
                     "reuse_full_task_seconds": (2.0 ** S) * reuse_model["amortised_s"],
                     "search_seconds": time.time() - t0}
                print(json.dumps(r), flush=True)
                results.append(r)
                save_case(os.path.join(a.out, name + ".case.gz"), name=name, pattern=sim.pattern, leaves=sim.tensors,
                          leaf_bonds=sim.tensor_bonds, scheme=sim.scheme, slicing_bonds=sim.slicing_bonds,
                          output_bonds=sim.output_bonds, permute_dims=sim.permute_dims if len(sim.output_bonds) else None,
                          bitstrings_sorted=getattr(sim, "bitstrings_sorted", None), n_qubits=len(sim.final_qubits),
                          extra={"prepare": {"sc_target": sc, "trials": a.trials, "iters": a.iters, "alpha": alpha, "start_seed": seed},
                                 "bitstrings_in": bitstrings, "n_shard_bonds": 0})
                with open(os.path.join(a.out, f"sweep_{a.circuit}.json"), "w") as f:
                    json.dump(results, f, indent=1)
    best = min(results, key=lambda r: r["predicted_full_task_seconds"])
    print("best:", json.dumps(best))
    fits = [r for r in results if r["reuse_workspace_gib"] <= 170]
    if fits:
        print("best with cross-slice reuse (workspace <= 170 GiB):", json.dumps(min(fits, key=lambda r: r["reuse_full_task_seconds"])))


if __name__ == "__main__":
    main()
