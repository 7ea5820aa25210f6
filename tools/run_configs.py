#!/usr/bin/env python
"""The five BASELINE.json configurations end to end through the public API (runs on the GPU box;
under torchrun the slice range is split over the ranks and the partial amplitudes are summed with
one NCCL all-reduce).  For each configuration: parity against the frozen reference outputs
(tests/golden), device time per slice and slices/s, in complex64 and in the complex-half mode.

    python tools/run_configs.py                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np
import torch
import torch.distributed as dist

from artensor_b200 import TensorNetworkSimulation, load_case
from artensor_b200 import contraction as C

CONFIGS = [
    # (BASELINE config, case, slices to run (None = all), which golden entry to compare with)
    ("1: n12 m14 full amplitude", "n12_full", None),
    ("2: n30 m14 full amplitude", "n30_full", None),
    ("3: n30 m14 sparse 10000 amplitudes (unsliced)", "n30_sparse10000", None),
    ("3': n30 m14 sparse 64 amplitudes, 16 slices, chunked", "n30_sparse64_sc26", None),
    ("3'': n30 m14 sparse 10000 amplitudes, 512 slices", "n30_sparse10000_sc27", None),
    ("4: n53 m12 sparse 1024 amplitudes, sliced", "n53_m12_sparse1024", 64),
    ("5: n53 m20 sparse 1024 amplitudes, sliced", "n53_m20_sparse1024", 8),
]


def fidelity(a, b):
    a, b = a.reshape(-1).to(torch.complex128), b.reshape(-1).to(torch.complex128)
    return (torch.vdot(a, b).abs() ** 2 / (torch.vdot(a, a).real * torch.vdot(b, b).real)).item()


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out_lines = []
    only = [a for a in sys.argv[1:] if not a.startswith("-")]           # optional: case names to run
    for title, name, n_run in CONFIGS:
        if only and name not in only:
            continue
        case = load_case(os.path.join(ROOT, "tests", "golden", f"{name}.case.gz"))
        exp = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.expected.npz"))
        sim = TensorNetworkSimulation.from_case(case)
        sim.permute_dims = None if name == "n30_full" else sim.permute_dims
        total = case.n_slices
        n = total if n_run is None else min(total, n_run * world)
        rng = (0, n)
        res = {}
        for mode, dtype in (("complex64", torch.complex64), ("complex-half", torch.complex32)):
            kw = dict(device=dev, slice_range=rng, dtype=dtype)
            if world > 1:
                kw["group"] = True
            sim.contraction(**kw)                                   # warm-up (plan, tensor maps)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            amps = sim.contraction(**kw)
            e1.record()
            torch.cuda.synchronize(dev)
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[mode] = (amps, float(t.item()))
        a64, ms64 = res["complex64"]
        ah, msh = res["complex-half"]
        line = {"config": title, "case": name, "gpus": world, "slices_run": n, "slices_total": total,
                "c64_ms": ms64, "c64_slices_per_s": n / (ms64 * 1e-3),
                "chalf_ms": msh, "chalf_slices_per_s": n / (msh * 1e-3),
                "chalf_fidelity_vs_c64": fidelity(a64, ah)}
        # parity of the complex64 mode against the frozen reference outputs, where the fixture
        # holds the slices that were run
        ids = list(exp["slice_ids"])
        if all(s in ids for s in range(n)):
            want = sum(exp["per_slice_c64"][ids.index(s)].astype(np.complex128) for s in range(n))
            if sim.permute_dims is not None and "shape" in exp:     # fixture: executor order; API: qubit order
                want = np.transpose(want.reshape(exp["shape"]), sim.permute_dims)
            got = a64.reshape(-1)
            if "sample_idx" in exp:
                got = got[torch.from_numpy(exp["sample_idx"]).to(dev)]
            got = got.cpu().numpy()
            rms = np.sqrt(np.mean(np.abs(want) ** 2))
            line["c64_max_err_over_rms_vs_reference"] = float(np.abs(got - want.reshape(-1)).max() / rms)
        if n == total and "google_amplitudes" in case.extra and case.bitstrings_sorted:
            google = dict(zip(case.extra["bitstrings_in"], case.extra["google_amplitudes"]))
            want = np.array([google[b] for b in case.bitstrings_sorted])
            rel = np.abs(a64.reshape(-1).cpu().numpy() - want) / np.abs(want)
            line["c64_median_rel_err_vs_google_amplitude_file"] = float(np.median(rel))
        if rank == 0:
            print(json.dumps(line), flush=True)
            out_lines.append(line)
        del res, a64, ah
        C.release_workspaces()
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"configs_n{world}.json"), "w") as f:
            json.dump(out_lines, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
