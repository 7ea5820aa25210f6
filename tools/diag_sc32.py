import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from artensor_b200 import TensorNetworkSimulation, load_case, PlanOptions
from artensor_b200 import contraction as C
name = sys.argv[1] if len(sys.argv) > 1 else "n53_m20_sparse1024_sc32"
case = load_case(os.path.join(ROOT, "tests", "golden", f"{name}.case.gz"))
want = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.expected.npz"))["per_slice_c64"][0].astype(np.complex128)
dev = torch.device("cuda:0")
rms = np.sqrt(np.mean(np.abs(want) ** 2))
def report(tag, got):
    sc = np.vdot(want, got) / np.vdot(want, want)
    print(f"{tag}: max|err|/rms {np.abs(got - want).max() / rms:.3e}  rms err/rms {np.sqrt(np.mean(np.abs(got - want) ** 2)) / rms:.3e}  "
          f"best-fit scale-1 {sc - 1:.3e}  resid max/rms {np.abs(got - sc * want).max() / rms:.3e}", flush=True)
off = 1 << 62
variants = {"default": PlanOptions(), "3xtf32": PlanOptions(tc_precision="3xtf32"),
            "no-skinny(stem fp32 instead)": PlanOptions(skinny_min_elems=off),
            "no-tc(stem/simt only where possible)": PlanOptions(tc_min_flops=float("inf"), skinny_min_elems=off)}
for tag, opt in variants.items():
    try:
        sim = TensorNetworkSimulation.from_case(case); sim.plan_options = opt
        got = sim.contraction(device=dev, slice_range=(0, 1)).cpu().numpy().reshape(-1).astype(np.complex128)
        report(tag, got)
    except Exception as e:
        print(tag, "failed:", str(e)[:200])
    del sim; C.release_workspaces(); torch.cuda.empty_cache()
