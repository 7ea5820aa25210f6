import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from artensor_b200 import TensorNetworkSimulation, load_case, PlanOptions
case = load_case("tests/golden/n30_sparse64_sc26.case.gz")
dev = torch.device("cuda:0")
for prec in ("3xf16", "3xtf32"):
    sim = TensorNetworkSimulation.from_case(case)
    sim.plan_options = PlanOptions(tc_precision=prec)
    plan = sim.plan()
    blob = plan.pack_leaves({k: v.to(dev) for k, v in case.leaves.items()})
    n = 4
    cur = torch.cuda.current_stream()
    wss = [torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    wss[1].fill_(0x7f)
    def run(ws, s, stream):
        o = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
        torch.cuda.synchronize()
        plan.execute(blob, o, s, s + 1, ws, stream.cuda_stream)
        return o
    serial = [run(wss[0], s, cur) for s in range(n)]; torch.cuda.synchronize()
    again = [run(wss[0], s, cur) for s in range(n)]; torch.cuda.synchronize()
    alt = [run(wss[s & 1], s, cur) for s in range(n)]; torch.cuda.synchronize()
    print(prec, "same stream, same ws again:", [(a - b).abs().max().item() for a, b in zip(serial, again)])
    print(prec, "same stream, alternating ws:", [(a - b).abs().max().item() for a, b in zip(serial, alt)])
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    for trial in range(3):
        outs = [torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev) for _ in range(n)]
        torch.cuda.synchronize()
        for s in range(n):
            plan.execute(blob, outs[s], s, s + 1, wss[s & 1], streams[s & 1].cuda_stream)
        torch.cuda.synchronize()
        print(prec, "two streams trial", trial, [(outs[s] - serial[s]).abs().max().item() for s in range(n)])
    del plan, sim
