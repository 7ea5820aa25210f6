#!/usr/bin/env python
"""Launch-bound share of light slices (runs on the GPU box): slices/s of a case with the slice phase
issued as ~100 stream launches per slice and replayed as one CUDA graph per slice
(PlanOptions.cuda_graph / TNC_OPT_CUDA_GRAPH), and the sum of the kernels' own device times.

    python tools/graph_bench.py n53_m12_sparse1024 [n_slices]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from artensor_b200 import ContractionPlan, PlanOptions, load_case


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "n53_m12_sparse1024"
    case = load_case(os.path.join(ROOT, "tests", "golden", f"{name}.case.gz"))
    dev = torch.device("cuda:0")
    shapes = {k: tuple(v.shape) for k, v in case.leaves.items()}
    res = {"case": name}
    outs = {}
    for tag, g in (("stream_launches", False), ("cuda_graph", True)):
        plan = ContractionPlan(case.scheme, shapes, case.pattern == "sparse", slicing_bonds=case.slicing_bonds,
                               slicing_indices=case.slicing_indices(), options=PlanOptions(cuda_graph=g))
        n = min(plan.n_slices, int(sys.argv[2]) if len(sys.argv) > 2 else 256)
        blob = plan.pack_leaves({k: v.to(dev) for k, v in case.leaves.items()})
        ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
        out = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        plan.execute(blob, out, 0, min(n, 8), ws, st)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            out.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.execute(blob, out, 0, n, ws, st)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        outs[tag] = out.clone()
        res[tag] = {"slices": n, "ms_per_slice": best / n, "slices_per_s": n / (best * 1e-3),
                    "launches_per_slice": plan.last_launches / n}
        del plan, ws
    res["identical_results"] = bool(torch.equal(outs["stream_launches"], outs["cuda_graph"]))
    res["speedup"] = res["stream_launches"]["ms_per_slice"] / res["cuda_graph"]["ms_per_slice"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
