"""Cross-slice reuse (PlanOptions.slice_reuse, TNC_OPT_SLICE_REUSE) measured on one GPU: seconds per slice over ranges
of R consecutive slice ids in ONE execute call, reference bit order and the reuse-optimised order
(TensorNetworkSimulation.optimize_slice_order), against contracting every step for every slice.
    python tools/reuse_bench.py n53_m20_sparse1024 [--ranges 2,8,64,512] [--check 4]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from artensor_b200 import TensorNetworkSimulation, PlanOptions, load_case   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--ranges", default="2,8,64,512")
    ap.add_argument("--check", type=int, default=4, help="slices compared bit for bit with full recomputation")
    ap.add_argument("--begin", type=int, default=0)
    ap.add_argument("--orders", default="reference,optimised")
    ap.add_argument("--keep-budget-gib", type=float, default=None, help="PlanOptions.keep_budget_bytes of the reuse plan")
    ap.add_argument("--no-plain", action="store_true", help="skip the plan without reuse (large arenas)")
    ap.add_argument("--graph", action="store_true", help="TNC_OPT_CUDA_GRAPH as well: one graph per class of changed bits")
    a = ap.parse_args()
    case = load_case(os.path.join(ROOT, "tests", "golden", f"{a.case}.case.gz"))
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    rows = []

    def timed(plan, blob, ws, lo, hi):
        out = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        plan.execute(blob, out, lo, hi, ws, st)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out, plan.last_launches

    for order in a.orders.split(","):
        sim = TensorNetworkSimulation.from_case(case)
        model = None
        if order == "optimised":
            sim.plan_options = PlanOptions(slice_reuse=True)
            model = sim.optimize_slice_order()
        plans = {}
        for reuse in (False, True):
            budget = None if a.keep_budget_gib is None else int(a.keep_budget_gib * 2 ** 30)
            sim.plan_options = PlanOptions(slice_reuse=reuse, cuda_graph=bool(a.graph and reuse), keep_budget_bytes=budget if reuse else None)
            sim._plan_cache.clear()
            plans[reuse] = sim.plan()
        n = plans[True].n_slices
        blob = plans[True].pack_leaves(case.leaves, device=dev)
        ws = torch.empty(max(p.workspace_bytes for p in plans.values()), dtype=torch.uint8, device=dev)
        print(f"{a.case} [{order} bit order]: {n.bit_length() - 1} sliced bonds, workspace {plans[False].workspace_bytes / 2**30:.2f} GiB "
              f"-> {plans[True].workspace_bytes / 2**30:.2f} GiB with reuse (KEEP {plans[True].keep_bytes / 2**30:.2f} GiB, "
              f"{sum(plans[True].step_tied)} steps tied to their reader; modelled {plans[True].reuse_summary()['amortised_s'] * 1e3:.2f} ms per slice)"
              + (f"; model {model}" if model else ""), flush=True)
        timed(plans[False], blob, ws, a.begin, a.begin + 1)                      # warm-up
        ms_full, _, l_full = timed(plans[False], blob, ws, a.begin, a.begin + 2)
        print(f"   every step for every slice: {ms_full / 2:9.3f} ms per slice, {l_full // 2} launches per slice", flush=True)
        if a.check:
            # (1) the reuse plan, one call per slice (every operation runs: the first slice of a call) against one call
            # over the range (operations skipped): the same kernels on the same operands, so bit for bit the same;
            # (2) against the plan without reuse, whose runs of tiny steps are chained differently (a step may be
            # summed by another kernel there): equal to rounding
            hi = min(n, a.begin + a.check)
            want = torch.zeros(plans[True].out_shape, dtype=torch.complex64, device=dev)
            for sid in range(a.begin, hi):
                plans[True].execute(blob, want, sid, sid + 1, ws, st)
            _, got, _ = timed(plans[True], blob, ws, a.begin, hi)
            same = bool(torch.equal(want, got))
            _, plain, _ = timed(plans[False], blob, ws, a.begin, hi)
            rms = plain.abs().pow(2).mean().sqrt().item()
            diff = (plain - got).abs().max().item() / rms
            print(f"   slices [{a.begin}, {hi}): one call with reuse vs one call per slice: bit-identical = {same}; "
                  f"vs the plan without reuse: max |diff| / rms = {diff:.2e}", flush=True)
            assert same and diff < 3e-6
        for R in [int(x) for x in a.ranges.split(",")]:
            hi = min(n, a.begin + R)
            timed(plans[True], blob, ws, a.begin, min(hi, a.begin + 2))
            ms, _, launches = timed(plans[True], blob, ws, a.begin, hi)
            per = ms / (hi - a.begin)
            print(f"   reuse, {hi - a.begin:5d} consecutive slices in one call: {ms:10.2f} ms = {per:9.3f} ms per slice "
                  f"({1e3 / per:8.1f} slices/s), {launches / (hi - a.begin):7.1f} launches per slice", flush=True)
            rows.append({"order": order, "slices": hi - a.begin, "ms": ms, "ms_per_slice": per, "launches": launches,
                         "ms_per_slice_full": ms_full / 2})
        del ws, blob
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"case": a.case, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", f"reuse_{a.case}.json"), "w"))


if __name__ == "__main__":
    main()
