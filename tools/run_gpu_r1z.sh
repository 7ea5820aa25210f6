mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests -q -m gpu -x -k "n12 or small or forced or smoke or reference" ) > gpurun_out/t_gpu_z.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_z.log; tail -n 3 gpurun_out/t_gpu_z.log
for c in n53_m12_sparse1024; do
  timeout -s KILL 400 python tools/gpu_probe.py $c --check --top 20 --tag _chain > gpurun_out/p_${c}_chain.log 2>&1
  sed -n 3,5p gpurun_out/p_${c}_chain.log
  TNC_NO_CHAIN=1 timeout -s KILL 400 python tools/gpu_probe.py $c --top 20 --tag _nochain > gpurun_out/p_${c}_nochain.log 2>&1
  sed -n 3,4p gpurun_out/p_${c}_nochain.log
done
python - <<'PY'
import torch, time, os, sys
sys.path.insert(0, os.getcwd())
from artensor_b200 import TensorNetworkSimulation, load_case
for name, n in (("n12_full", 1), ("n53_m12_sparse1024", 64)):
    case = load_case(f"tests/golden/{name}.case.gz")
    sim = TensorNetworkSimulation.from_case(case)
    for rep in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sim.contraction(device="cuda:0", slice_range=(0, n)); e1.record(); torch.cuda.synchronize()
    print(name, "chain" if "TNC_NO_CHAIN" not in os.environ else "nochain", f"{e0.elapsed_time(e1):.3f} ms for {n} slices")
PY
TNC_NO_CHAIN=1 python - <<'PY'
import torch, time, os, sys
sys.path.insert(0, os.getcwd())
from artensor_b200 import TensorNetworkSimulation, load_case
for name, n in (("n12_full", 1), ("n53_m12_sparse1024", 64)):
    case = load_case(f"tests/golden/{name}.case.gz")
    sim = TensorNetworkSimulation.from_case(case)
    for rep in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sim.contraction(device="cuda:0", slice_range=(0, n)); e1.record(); torch.cuda.synchronize()
    print(name, "nochain", f"{e0.elapsed_time(e1):.3f} ms for {n} slices")
PY
