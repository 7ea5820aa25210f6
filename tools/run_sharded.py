#!/usr/bin/env python
"""BASELINE config 2 (n30 m14 full amplitude) on 1/2/4/8 GPUs, two ways (runs on the GPU box):

  sharded   fixture n30_full_shard3: the first 3 output qubits are fixed per shard
            (`prepare_open_qubit_shards`), the 8 shards are block-partitioned over the ranks, every
            rank contracts all 4 regular slices of its shards; results concatenate, nothing is reduced;
  sliced    fixture n30_full: the 4 slices are partitioned over the ranks and the partial 2^28-amplitude
            tensors are summed with one all-reduce (the only option before; at most 4 ranks have work).

    python tools/run_sharded.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_sharded.py

Per rank the sharded run is checked against the reference executor's recorded outputs of the
(shard, slice) ids the fixture holds that fall into the rank's shards."""
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


def timed(fn, dev, world, reps=3):
    fn()
    best = None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
        del out
    return best


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    group = True if world > 1 else None
    line = {"gpus": world}

    case = load_case(os.path.join(ROOT, "tests", "golden", "n30_full_shard3.case.gz"))
    exp = np.load(os.path.join(ROOT, "tests", "golden", "n30_full_shard3.expected.npz"))
    sim = TensorNetworkSimulation.from_case(case)
    plan = sim.plan()
    n_shards = 1 << len(sim.shard_bonds)
    per_shard = plan.n_slices // n_shards
    line["sharded_ms"] = timed(lambda: sim.contraction(device=dev, group=group), dev, world)
    line["sharded_amplitudes_per_rank"] = (n_shards // world) << len(plan.out_shape)
    # parity on this rank's shards
    first, last = n_shards * rank // world, n_shards * (rank + 1) // world
    blob = plan.pack_leaves({k: v.to(dev) for k, v in case.leaves.items()})
    ws = C.get_workspace(dev, plan.workspace_bytes)
    idx = torch.from_numpy(exp["sample_idx"]).to(dev)
    worst, checked = 0.0, 0
    for k, s in enumerate(int(x) for x in exp["slice_ids"]):
        if not first <= s // per_shard < last:
            continue
        out = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
        plan.execute(blob, out, s, s + 1, ws, torch.cuda.current_stream().cuda_stream)
        got = out.reshape(-1)[idx].cpu().numpy()
        want = exp["per_slice_c64"][k]
        worst = max(worst, float(np.abs(got - want).max() / np.sqrt(np.mean(np.abs(want) ** 2))))
        checked += 1
    t = torch.tensor([worst, float(checked)], device=dev, dtype=torch.float64)
    if world > 1:
        w = t.clone()
        dist.all_reduce(w[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        t[0] = w[0]
    line["sharded_max_err_over_rms_vs_reference"] = float(t[0].item())
    line["sharded_fixture_slices_checked"] = int(t[1].item())
    del sim, plan, blob, ws
    C.release_workspaces()
    torch.cuda.empty_cache()

    case = load_case(os.path.join(ROOT, "tests", "golden", "n30_full.case.gz"))
    sim = TensorNetworkSimulation.from_case(case)
    sim.permute_dims = None
    line["sliced_allreduce_ms"] = timed(lambda: sim.contraction(device=dev, group=group), dev, world)
    line["sliced_ranks_with_work"] = min(world, sim.plan().n_slices)
    if rank == 0:
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"sharded_n{world}.json"), "w") as f:
            json.dump(line, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
