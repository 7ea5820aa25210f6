#!/usr/bin/env python
"""bench.py -- slice-contractions per second of the B200-native contraction executor.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N ...             (the reference's CPU torch path)

Metric (BASELINE.json): slice-contractions/sec on a Sycamore n53 sliced contraction tree.
Workload at N=1: BASELINE config 5, `n53_m20_sparse1024` -- the n53 m20 tree found by the
reference's own order finder (sc_target=30, 50 sliced bonds), 1024 amplitudes per slice, a fixed
subset of slice ids (the full task is 2^50 slices: time extrapolated, not run).  A *step* is one
pass of the hot path over one batch of `--slices-per-step` slices per GPU followed by the ONE sum
of the partial amplitudes (NCCL all-reduce when N > 1).  Weak scaling: every rank contracts its
own `--slices-per-step` slices per step; `value` = slices all ranks contracted / max-over-ranks
device time.

Timed regions
  value   leaf blob, plan, workspace already resident in HBM; CUDA events on the launching stream,
          barrier + synchronize on both sides, max over ranks.  Working set per slice (8 GiB
          intermediates) is far larger than the 126 MB L2, so no explicit L2 flush is needed.
  e2e     the same K steps through the public API `TensorNetworkSimulation.contraction` with HOST
          leaf tensors: per step one pinned host->device copy of the leaf blob and one
          device->host read of the amplitudes.
  roofline  the dominant kernel of a slice (the fat tcgen05 GEMM) timed with CUDA events per launch
          via tnc_plan_profile; achieved = 8*M*N*K algorithmic flops / that time.
  cpu_baseline / --impl reference
          the reference's own executor restated on CPU torch (oracle/tn_oracle_torch.py) on the host
          cores.  `--impl reference` runs ONE TRUE SLICE -- the real leaves through every scheme step --
          spread over its K timed steps (step i = the i-th 1/K of the slice's scheme steps), checks
          the result against the reference's recorded output and caches the seconds under /tmp;
          the GPU arm's `cpu_baseline` quotes that measurement when it finds it on the same host,
          else a bounded timing model (every step's einsum on synthetic operands, large steps on
          sub-blocks), which is also printed beside the true slice as a cross-check.
  slice_reuse  (extra object, NOT the headline) the same tree with PlanOptions.slice_reuse and the sliced bonds
          re-ordered for it: `--reuse-slices` consecutive slice ids per GPU in ONE call through the public API, a
          step contracted again only when a sliced bond behind it changed (bit-identical amplitudes).  `value` and
          `e2e` contract every step for every slice, as the reference's loop does.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line.  NCCL prints its banner ("NCCL version ...") with a plain
# printf to file descriptor 1 whatever NCCL_DEBUG_FILE says, so descriptor 1 is pointed at stderr
# for the whole run and the JSON line is written to a saved copy of the real stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "slice_contractions_per_sec"
UNIT = "slices/s"
DEFAULT_WORKLOAD = "n53_m20_sparse1024"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("TNC_BENCH_WORKLOAD", DEFAULT_WORKLOAD))
    ap.add_argument("--slices-per-step", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tc-min-flops", type=float, default=None)
    ap.add_argument("--cpu-max-elems", type=int, default=1 << 24)
    ap.add_argument("--tc-precision", default=None, choices=["3xtf32", "3xf16", "f16"],
                    help="operand precision of the tensor-core steps (default: the library default, 3xf16)")
    ap.add_argument("--no-half", action="store_true", help="skip the complex-half mode measurement")
    ap.add_argument("--no-fuse-amax", action="store_true", help="A/B aid: keep the separate amax pass of every tensor-core step")
    ap.add_argument("--no-reuse", action="store_true", help="skip the cross-slice reuse measurement (`slice_reuse` key)")
    ap.add_argument("--reuse-slices", type=int, default=256, help="consecutive slice ids per execute call of that measurement")
    return ap.parse_args()


def load_workload(name):
    from artensor_b200.cases import load_case
    return load_case(os.path.join(ROOT, "tests", "golden", f"{name}.case.gz"))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measure_cublas_peak(dev, kind, seconds=1.0):
    """cuBLAS GEMM 8192^3 in the operand type the kernel issues ("tf32": fp32 inputs with allow_tf32,
    "f16": fp16 inputs), measured the way MEASURED_PEAKS.json measures bf16: best of 10 (burst) and
    back-to-back for `seconds` (sustained).  Only a denominator, never timed work."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        dt = torch.float32 if kind == "tf32" else torch.float16
        a = torch.randn(n, n, device=dev, dtype=dt)
        b = torch.randn(n, n, device=dev, dtype=dt)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(seconds * 1e3 / best))
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record(); torch.cuda.synchronize(dev)
        flops = 2.0 * n ** 3
        return {"burst": flops / (best * 1e-3) / 1e12, "sustained": flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def step_shapes_of(plan):
    """(eq, shape_a, shape_b) of every scheme step with the shapes the reference's einsum sees
    (gathered row counts for batched steps)."""
    out = []
    for st, raw in zip(plan.steps, plan.scheme_steps):
        eq = raw[1]
        sa, sb = [2] * st.a.rank, [2] * st.b.rank
        if st.kind == "batched":
            sa, sb = [st.nb] + sa, [st.nb] + sb
        else:
            if st.a.rows is not None:
                sa = [st.a.rows] + sa
            if st.b.rows is not None:
                sb = [st.b.rows] + sb
        out.append((eq, sa, sb))
    return out


def cpu_sample(plan, max_elems, budget_s=240.0):
    """Timing MODEL of one slice on the host cores (a cross-check, not the baseline): every step's
    torch.einsum on synthetic operands of its true shape, steps above `max_elems` elements on a
    sub-block and scaled linearly."""
    import torch
    from oracle import tn_oracle_torch as OT
    torch.set_num_threads(os.cpu_count() or 1)
    secs, n, scaled, wall = OT.estimate_slice_seconds(step_shapes_of(plan), max_elems=max_elems, budget_s=budget_s)
    return secs, n, scaled, wall


def host_mem_gib():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 2 ** 20
    except OSError:
        pass
    return None


def cpu_cache_path(workload):
    import socket
    return os.path.join("/tmp", f"tnc_cpu_true_slice_{workload}_{socket.gethostname()}_{os.cpu_count()}.json")


def check_against_golden(workload, slice_id, result):
    """max |err| / rms of a CPU slice result against the reference's recorded output, when the
    fixture holds this slice id (None otherwise)."""
    import numpy as np
    try:
        exp = np.load(os.path.join(ROOT, "tests", "golden", f"{workload}.expected.npz"))
        ids = [int(x) for x in exp["slice_ids"]]
        if slice_id not in ids:
            return None
        want = exp["per_slice_c64"][ids.index(slice_id)].reshape(-1)
        got = result.reshape(-1).numpy()
        return float(np.abs(got - want).max() / np.sqrt(np.mean(np.abs(want) ** 2)))
    except (OSError, KeyError):
        return None


def reference_arm(args):
    """The reference's CPU path on the host cores; rank 0 only.

    ONE TRUE SLICE of the workload -- its real leaves through every scheme step exactly as
    `tensor_contraction_sparse` executes them (oracle/tn_oracle_torch.py: run_sparse on the sliced
    leaves) -- is spread over the K timed steps: step i executes the i-th contiguous 1/K of the
    slice's scheme steps, so the timed region is one real slice, every step a bounded sample, and
    `value` = 1 slice / (sum of the K step times).  The W warm-up steps execute the leading W/K of
    another slice and are discarded.  The result is compared with the reference's recorded output
    of that slice, and the synthetic-operand timing model is printed beside it as a cross-check."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import tn_oracle_torch as OT
    case = load_workload(args.workload)
    plan = plan_of(case, None, build_native=False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, W = max(1, args.steps), max(0, args.warmup)
    n_steps = len(case.scheme)
    bounds = [(n_steps * i) // K for i in range(K + 1)]
    try:
        import numpy as np
        exp = np.load(os.path.join(ROOT, "tests", "golden", f"{args.workload}.expected.npz"))
        slice_id = int(exp["slice_ids"][0])
    except (OSError, KeyError):
        slice_id = 0
    # host memory a true slice needs: live intermediates + einsum's permuted operand copies + its output
    live, cur, peak = {}, 0, 0
    for st in plan.steps:
        a, b, c = st.a.numel * 8, st.b.numel * 8, st.c.numel * 8
        peak = max(peak, cur + a + b + c)
        cur += c - live.pop(st.i, 0) - live.pop(st.j, 0)
        live[st.i] = c
    need_gib = 1.3 * peak / 2 ** 30 + 2.0
    avail = host_mem_gib()
    mode = "true_slice"
    if avail is not None and avail < need_gib and not os.environ.get("TNC_BENCH_FORCE_TRUE_SLICE"):
        mode = "model"          # the host cannot hold the slice's intermediates: say so loudly, fall back to the model
    step_ms, result_err, true_secs = [], None, None
    if mode == "true_slice":
        # warm-up: leading steps of another slice, discarded
        if W > 0:
            gen = OT.slice_stepper(case, (slice_id + 1) % max(1, 1 << len(case.slicing_bonds)))
            stop = bounds[min(W, K)]
            for k, _ in gen:
                if k is None or k + 1 >= stop:
                    break
            del gen
        gen = OT.slice_stepper(case, slice_id)
        per_step = []
        result = None
        for k, v in gen:
            if k is None:
                result = v
            else:
                per_step.append(v)
        for i in range(K):
            step_ms.append(1e3 * sum(per_step[bounds[i]:bounds[i + 1]]))
        true_secs = sum(per_step)
        result_err = check_against_golden(args.workload, slice_id, result)
        with open(cpu_cache_path(args.workload), "w") as f:
            json.dump({"workload": args.workload, "slice_id": slice_id, "seconds": true_secs, "cores": cores,
                       "torch": torch.__version__, "max_err_over_rms_vs_reference_output": result_err}, f)
    model = None
    try:
        secs, n, scaled, wall = cpu_sample(plan, args.cpu_max_elems, budget_s=120.0)
        model = {"seconds_per_slice": secs, "steps_scaled_from_sub_blocks": scaled, "cpu_seconds_spent": wall,
                 "error_vs_true_slice": (secs / true_secs - 1.0) if true_secs else None}
    except Exception as exc:      # the cross-check must not take the line down
        model = {"failed": str(exc)}
    if mode == "model":
        true_secs = model["seconds_per_slice"]
        step_ms = [1e3 * true_secs / K] * K
    value = 1.0 / true_secs
    if mode == "true_slice":
        sample = (f"ONE true slice (id {slice_id}) of {args.workload}: the real leaves through all {n_steps} scheme steps "
                  f"with torch.einsum / gather / cat exactly as tensor_contraction_sparse runs them "
                  f"(oracle/tn_oracle_torch.py), {true_secs:.1f} s on {cores} threads; the {K} timed steps are the "
                  f"consecutive 1/{K} parts of that slice, the {W} warm-up steps the leading part of another slice; "
                  f"max |err| / rms against the reference's recorded output of the slice: {result_err}")
    else:
        sample = (f"TIMING MODEL ONLY (host has {avail:.0f} GiB available, a true slice needs ~{need_gib:.0f} GiB): every "
                  f"step's torch.einsum on synthetic operands, large steps on sub-blocks scaled linearly")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sum(step_ms) / len(step_ms), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": bench_config(args.workload, plan, args.slices_per_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "mode": mode, "torch": torch.__version__, "timing_model_cross_check": model},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def bench_config(workload, plan, slices_per_step):
    """The `config` object both arms print (identical keys and values for the same workload)."""
    work = plan.work_summary()
    return {"workload": workload, "slices_per_step_per_gpu": slices_per_step, "sliced_bonds": plan.n_sliced,
            "total_slices_of_task": f"2^{plan.n_sliced}",
            "amplitudes_per_slice": int(math.prod(plan.out_shape)),
            "scheme_steps": work["steps"], "l2": "working set >> L2 (multi-GiB intermediates), no flush",
            "every_step_for_every_slice": True}      # the reference's loop; cross-slice reuse only in `slice_reuse`


def plan_of(case, tc_min_flops, build_native=True):
    from artensor_b200.backend import ContractionPlan, PlanOptions
    opts = PlanOptions() if tc_min_flops is None else PlanOptions(tc_min_flops=tc_min_flops)
    plan = ContractionPlan(case.scheme, {k: tuple(v.shape) for k, v in case.leaves.items()}, case.pattern == "sparse",
                           slicing_bonds=case.slicing_bonds, slicing_indices=case.slicing_indices(), options=opts,
                           build_native=build_native)
    plan.scheme_steps = case.scheme
    return plan


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from artensor_b200 import TensorNetworkSimulation, PlanOptions
    from artensor_b200 import _native as N
    from artensor_b200 import contraction as C
    from artensor_b200.backend import tc_uses_3m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (artensor_b200 has no CPU fallback)")
    N.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    case = load_workload(args.workload)
    sim = TensorNetworkSimulation.from_case(case)
    opt_kw = {}
    if args.tc_min_flops is not None:
        opt_kw["tc_min_flops"] = args.tc_min_flops
    if args.tc_precision is not None:
        opt_kw["tc_precision"] = args.tc_precision
    if args.no_fuse_amax:
        opt_kw["fuse_amax"] = False
    sim.plan_options = PlanOptions(**opt_kw)
    precision = sim.plan_options.tc_precision
    plan = sim.plan()
    plan.scheme_steps = case.scheme
    S = args.slices_per_step
    n_slices = plan.n_slices
    total_steps = args.warmup + args.steps

    def slice_range(step):
        lo = ((step * world + rank) * S) % max(1, n_slices - S + 1) if n_slices > S else 0
        return lo, min(lo + S, n_slices)

    stream = torch.cuda.current_stream(dev)
    blob = plan.pack_leaves(case.leaves, device=dev)
    out = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
    ws = C.get_workspace(dev, plan.workspace_bytes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step(step):
        lo, hi = slice_range(step)
        out.zero_()
        plan.execute(blob, out, lo, hi, ws, stream.cuda_stream)
        if world > 1:
            dist.all_reduce(torch.view_as_real(out), op=dist.ReduceOp.SUM)
        return hi - lo

    # ---- device-resident throughput ("value")
    for it in range(args.warmup):
        one_step(it)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    done = 0
    launches = 0
    for it in range(args.warmup, total_steps):
        done += one_step(it)
        launches += plan.last_launches + 1          # + the accumulator clear
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(done), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max, total_slices, launches = float(t.item()), float(cnt[0].item()), int(cnt[1].item())
    value = total_slices / (ms_max * 1e-3)

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        host = {k: v.pin_memory() for k, v in case.leaves.items()}

        def e2e_step(it):
            # the API block-partitions the range it is given over the ranks of the group
            if world > 1:
                glo = (it * world * S) % max(1, n_slices - world * S + 1)
                r = sim.contraction(tensors=host, device=dev, slice_range=(glo, glo + world * S), group=True)
            else:
                r = sim.contraction(tensors=host, device=dev, slice_range=slice_range(it))
            return r.cpu()          # device -> host read of the amplitudes

        res = None
        for it in range(min(2, args.warmup)):
            res = e2e_step(it)
        barrier()
        t0 = time.perf_counter()
        done2 = 0
        for it in range(args.warmup, total_steps):
            res = e2e_step(it)
            done2 += S
        barrier()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], dtype=torch.float64, device=dev)
        c2 = torch.tensor([float(done2)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            dist.all_reduce(c2, op=dist.ReduceOp.SUM)
        e2e = {"value": float(c2.item()) / float(t2.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(plan.leaf_blob_elems * 8), "d2h_bytes_per_step": int(res.numel() * 8)}

    # ---- roofline of the dominant kernel, measured live (rank 0)
    roofline, work, breakdown, slice_roofline = None, plan.work_summary(), None, None
    if rank == 0:
        pk = peaks()
        reps = 2
        acc = None
        for i in range(reps):
            _, ms_slice = plan.profile(blob, out, slice_range(args.warmup + i)[0], ws, stream.cuda_stream)
            acc = ms_slice if acc is None else [a + b for a, b in zip(acc, ms_slice)]
        ms_slice = [a / reps for a in acc]
        SL = N.TNC_PROFILE_SLOTS
        ops = plan.ops[N.TNC_PHASE_SLICE]
        steps = plan.op_steps[N.TNC_PHASE_SLICE]
        slice_ms = sum(ms_slice[i * SL] for i in range(len(ops)))
        best, best_ms = None, -1.0
        gemm_ms = pack_ms = simt_ms = stem_ms = skinny_ms = 0.0
        stem_bytes = skinny_bytes = 0
        for i, ((kind, rec), st) in enumerate(zip(ops, steps)):
            if kind != "einsum":
                continue
            if rec.algo == N.TNC_ALGO_TC:
                k_ms = ms_slice[i * SL + 3]
                gemm_ms += k_ms
                pack_ms += ms_slice[i * SL + 1] + ms_slice[i * SL + 2]
            elif rec.algo == N.TNC_ALGO_STEM:
                k_ms = ms_slice[i * SL]
                stem_ms += k_ms
                stem_bytes += st.bytes_c64
            elif rec.algo == N.TNC_ALGO_SKINNY:
                k_ms = ms_slice[i * SL]
                skinny_ms += k_ms
                skinny_bytes += st.bytes_c64
            else:
                k_ms = ms_slice[i * SL]
                simt_ms += k_ms
            if k_ms > best_ms:
                best, best_ms = (i, rec, st), k_ms
        i, rec, st = best
        ai = st.flops / st.bytes_c64
        tensor_bound = rec.algo == N.TNC_ALGO_TC and ai > 100
        if tensor_bound:
            ach = st.flops / (best_ms * 1e-3) / 1e12
            peak = pk["bf16_tflops_sustained"]
            kind = "tf32" if precision == "3xtf32" else "f16"
            three_m = tc_uses_3m(st, precision)
            # real tensor-core products issued per useful complex one: hi/lo split (x3), 3M complex product (x0.75)
            products = (1 if precision == "f16" else 3) * (0.75 if three_m else 1.0)
            lib = measure_cublas_peak(dev, kind)
            roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "traffic": None, "peak_kind": f"bf16 dense sustained ({pk['source']})",
                        "issued": {"operand_type": kind, "products_per_useful_flop": products,
                                   "issued_tflops": products * ach, f"cublas_{kind}_burst": lib["burst"],
                                   f"cublas_{kind}_sustained": lib["sustained"],
                                   "frac_of_cublas_sustained": products * ach / lib["sustained"],
                                   "frac_of_peak": products * ach / peak},
                        "note": "achieved = 8*M*N*K useful complex64 flops.  The fp32-accurate precisions split every "
                                "operand in hi + lo (3 real products per useful one) and the 3M complex product needs "
                                "6*M*N*K real flops instead of 8 (x0.75), so the useful-flop ceiling is peak / "
                                "products_per_useful_flop; `issued` compares the issued tensor flops with cuBLAS "
                                "8192^3 in the same operand type measured in this run"}
        else:
            ach = st.bytes_c64 / (best_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                        "traffic": None, "peak_kind": f"copy bandwidth ({pk['source']})"}
        kname = {N.TNC_ALGO_TC: f"{'gemm3m_2cta_kernel' if tc_uses_3m(st, precision) else 'gemm_2cta_kernel'}<{precision}>", N.TNC_ALGO_STEM: "stem_kernel", N.TNC_ALGO_SKINNY: "skinny_kernel",
                 N.TNC_ALGO_SIMT: "simt_einsum_kernel"}[rec.algo]
        roofline["kernel"] = kname
        # DRAM bytes per launch of this kernel on a step of this shape: the committed ncu captures
        # (profiles/ncu_traffic.json lists kernel, shape and source file of every capture)
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                caps = json.load(f)["captures"]
            shape = [len(st.m_modes), len(st.n_modes), len(st.k_modes)]
            for cap in caps:
                if cap["kernel"] == kname and cap["shape_bits_mnk"] == shape:
                    roofline["traffic"] = cap["dram_bytes"]
                    roofline["traffic_note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel "
                                                f"on this step shape (ncu --set full, {cap['source']}); algorithmic bytes "
                                                f"of the step: {st.bytes_c64}")
        except (OSError, ValueError, KeyError):
            pass
        roofline["step"] = {"index": st.index, "m_bits": len(st.m_modes), "n_bits": len(st.n_modes),
                            "k_bits": len(st.k_modes), "rows": st.nb, "flops": st.flops, "bytes": st.bytes_c64,
                            "ms": best_ms, "share_of_slice": best_ms / slice_ms}
        breakdown = {"slice_ms_profiled": slice_ms, "gemm_ms": gemm_ms, "pack_ms": pack_ms, "stem_ms": stem_ms,
                     "skinny_ms": skinny_ms, "generic_ms": simt_ms,
                     "other_ms": slice_ms - gemm_ms - pack_ms - simt_ms - stem_ms - skinny_ms,
                     # the HBM-bound step classes inside the power-capped slice: algorithmic bytes / time
                     "stem_gbs": stem_bytes / (stem_ms * 1e-3) / 1e9 if stem_ms > 0 else None,
                     "skinny_gbs": skinny_bytes / (skinny_ms * 1e-3) / 1e9 if skinny_ms > 0 else None,
                     "hbm_peak_gbs": pk["hbm_gbs"]}
        # whole-slice roofline: sum over the executed steps of max(flops / P, bytes / BW), against
        # the steady-state time of a slice in the timed region.  Two ceilings: P = measured dense peak (one
        # tensor-core product per useful flop: the complex-half mode's ceiling) and P / products for the
        # products each step really has to issue in this precision (hi/lo split, 3M where it applies).
        ideal_single = ideal_issued = 0.0
        P, BW = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
        for (kind, rec), st in zip(ops, steps):
            if kind != "einsum":
                continue
            prod = 1.0
            if rec.algo == N.TNC_ALGO_TC:
                prod = (1 if precision == "f16" else 3) * (0.75 if tc_uses_3m(st, precision) else 1.0)
            ideal_single += max(st.flops / P, st.bytes_c64 / BW)
            ideal_issued += max(st.flops * prod / P, st.bytes_c64 / BW)
        ms_per_slice = ms_max / (total_slices / world)
        slice_roofline = {"ideal_ms_one_product": ideal_single * 1e3, "ideal_ms_issued_products": ideal_issued * 1e3,
                          "measured_ms_per_slice": ms_per_slice, "frac_one_product": ideal_single * 1e3 / ms_per_slice,
                          "frac_issued_products": ideal_issued * 1e3 / ms_per_slice,
                          "note": "sum over executed steps of max(8BMNK / P, 8(|A|+|B|+|C|) / BW); P = measured bf16 "
                                  "dense sustained, BW = measured copy bandwidth; measured = steady-state ms per slice "
                                  "of the timed region (`value`)"}

    # ---- the reduced-precision complex-half mode on the same slices (rank 0, N = 1): throughput and
    # fidelity against the complex64 result
    half = None
    if rank == 0 and world == 1 and not args.no_half and precision != "f16":
        from dataclasses import replace
        hplan = C.ContractionPlan(case.scheme, {k: tuple(v.shape) for k, v in case.leaves.items()},
                                  case.pattern == "sparse", slicing_bonds=case.slicing_bonds,
                                  slicing_indices=case.slicing_indices(), options=replace(sim.plan_options, tc_precision="f16"))
        lo, hi = slice_range(args.warmup)
        ws = C.get_workspace(dev, max(plan.workspace_bytes, hplan.workspace_bytes))
        out.zero_()
        plan.execute(blob, out, lo, hi, ws, stream.cuda_stream)
        ref64 = out.clone()
        hout = torch.zeros_like(out)
        for _ in range(2):
            hout.zero_()
            hplan.execute(blob, hout, lo, hi, ws, stream.cuda_stream)
        torch.cuda.synchronize(dev)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        for _ in range(2):
            hout.zero_()
            hplan.execute(blob, hout, lo, hi, ws, stream.cuda_stream)
        h1.record(stream)
        torch.cuda.synchronize(dev)
        a, b = ref64.reshape(-1).to(torch.complex128), hout.reshape(-1).to(torch.complex128)
        fid = (torch.vdot(a, b).abs() ** 2 / (torch.vdot(a, a).real * torch.vdot(b, b).real)).item()
        half = {"value": 2 * (hi - lo) / (h0.elapsed_time(h1) * 1e-3), "unit": UNIT, "tc_precision": "f16",
                "fidelity_vs_complex64": fid,
                "max_err_over_rms": ((a - b).abs().max() / a.abs().pow(2).mean().sqrt()).item()}
        # the same dominant GEMM in this mode issues ONE tensor-core product per useful one: its useful
        # TFLOP/s are the issued ones (CUDA events around the launch, one profiled slice)
        try:
            _, hms = hplan.profile(blob, hout, lo, ws, stream.cuda_stream)
            SLh = N.TNC_PROFILE_SLOTS
            hbest = None
            for i, ((kind, rec), st) in enumerate(zip(hplan.ops[N.TNC_PHASE_SLICE], hplan.op_steps[N.TNC_PHASE_SLICE])):
                if kind == "einsum" and rec.algo == N.TNC_ALGO_TC and (hbest is None or hms[i * SLh + 3] > hbest[0]):
                    hbest = (hms[i * SLh + 3], st)
            if hbest is not None and hbest[0] > 0:
                hach = hbest[1].flops / (hbest[0] * 1e-3) / 1e12
                half["roofline"] = {"bound": "tensor",
                                    "kernel": ("gemm3m_2cta_kernel" if tc_uses_3m(hbest[1], "f16") else "gemm_2cta_kernel") + "<f16>",
                                    "products_per_useful_flop": 0.75 if tc_uses_3m(hbest[1], "f16") else 1.0,
                                    "achieved": hach,
                                    "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                    "frac": hach / pk["bf16_tflops_sustained"],
                                    "step": {"index": hbest[1].index, "ms": hbest[0], "flops": hbest[1].flops}}
        except Exception as exc:      # measurement aid only
            half["roofline_error"] = str(exc)
        del hplan

    # ---- cross-slice reuse (PlanOptions.slice_reuse) on the same tree: every rank contracts ONE range of R consecutive
    # slice ids in one execute call, the sliced bonds re-ordered for reuse.  Not the headline: `value` / `e2e` above
    # contract every step for every slice, as the reference's loop does; here a step is contracted again only when a
    # sliced bond behind it changed, with bit-identical amplitudes (tests/test_gpu_parity.py).
    reuse = None
    if not args.no_reuse and plan.n_sliced >= 2:
        try:
            # set-up (no collectives): every rank then votes, and the collective part below runs on all ranks or on none
            problem = None
            try:
                rsim = TensorNetworkSimulation.from_case(case)
                rsim.plan_options = PlanOptions(**dict(opt_kw, slice_reuse=True, cuda_graph=False))
                model = rsim.optimize_slice_order()
                rplan = rsim.plan()
                ws = None
                C.release_workspaces()
                torch.cuda.empty_cache()
                free_b, _ = torch.cuda.mem_get_info(dev)
                # bound the KEEP region if need be: steps are tied to their readers until the workspace fits
                rplan = rsim.fit_reuse_to_memory(free_b)
                if rplan.workspace_bytes > free_b - (2 << 30):
                    problem = f"workspace with reuse {rplan.workspace_bytes >> 30} GiB > free HBM {free_b >> 30} GiB"
            except Exception as exc:
                problem = str(exc)
            vote = torch.tensor([0.0 if problem is None else 1.0], device=dev)
            if world > 1:
                dist.all_reduce(vote, op=dist.ReduceOp.MAX)
            if vote.item() > 0:
                raise RuntimeError(problem or "another rank could not set up the reuse plan")
            R = min(args.reuse_slices, rplan.n_slices // world)
            rhost = {k: v.pin_memory() for k, v in case.leaves.items()}
            grp = True if world > 1 else None
            # through the public API with host leaves, like `e2e`: the API block-partitions the range over the
            # ranks (R consecutive slice ids each), all-reduces the partial amplitudes, the result is read back
            rsim.contraction(tensors=rhost, device=dev, slice_range=(0, world * min(R, 2)), group=grp).cpu()   # warm-up
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            rsim.contraction(tensors=rhost, device=dev, slice_range=(0, world * R), group=grp).cpu()
            r1.record(stream)
            barrier()
            rt = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            rms = float(rt.item())
            reuse = {"value": world * R / (rms * 1e-3), "unit": UNIT, "slices_per_call_per_gpu": R,
                     "ms_per_slice_per_gpu": rms / R, "launches_per_slice": rplan.last_launches / R,
                     "bit_order": "sliced bonds re-ordered by TensorNetworkSimulation.optimize_slice_order",
                     "workspace_gib": rplan.workspace_bytes / 2 ** 30, "keep_gib": rplan.keep_bytes / 2 ** 30,
                     "steps_tied_to_their_reader": int(sum(rplan.step_tied)),
                     "modelled_ms_per_slice": {"every_step": model["full_s"] * 1e3,
                                               "reuse_reference_bit_order": model["amortised_before_s"] * 1e3,
                                               "reuse": model["amortised_s"] * 1e3},
                     "extrapolated_full_task_seconds": (2.0 ** rplan.n_sliced) * (rms * 1e-3 / R) / world,
                     "measured": "end to end through TensorNetworkSimulation.contraction with pinned host leaves, result read back",
                     "note": "same tree, same slices, bit-identical amplitudes: inside one call a step is contracted again "
                             "only when a sliced bond behind it changed from the previous slice id; value / e2e above do "
                             "NOT use it (every step contracted for every slice, like the reference's slice loop)"}
            del rplan
        except Exception as exc:      # an extra measurement must never take the line down
            reuse = {"value": None, "error": str(exc)}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cached = None
        try:
            with open(cpu_cache_path(args.workload)) as f:
                cached = json.load(f)
        except (OSError, ValueError):
            pass
        try:
            secs, n, scaled, wall = cpu_sample(plan, args.cpu_max_elems)
            model = (f"timing model: all {n} steps run with torch.einsum on synthetic operands of their true shapes, "
                     f"{scaled} steps above 2^{args.cpu_max_elems.bit_length() - 1} elements on a sub-block and scaled "
                     f"linearly ({wall:.1f} s of CPU work) = {secs:.1f} s per slice")
            if cached and cached.get("cores") == cores:
                cpu = {"value": 1.0 / cached["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": (f"ONE true slice (id {cached['slice_id']}) of {args.workload} run on this host by "
                                  f"`bench.py --impl reference` ({cached['seconds']:.1f} s, real leaves through every "
                                  f"scheme step, oracle/tn_oracle_torch.py; max |err| / rms vs the reference's recorded "
                                  f"output {cached.get('max_err_over_rms_vs_reference_output')}); cross-check in this run, "
                                  f"{model} ({secs / cached['seconds'] - 1.0:+.1%} vs the true slice)")}
            else:
                cpu = {"value": 1.0 / secs, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": f"one slice of {args.workload}, {model} (no true-slice measurement cached on this host: "
                                 f"run `bench.py --impl reference` first)"}
        except Exception as exc:  # the baseline must never take the GPU line down
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": f"complex64 ({precision} split-precision products on tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": bench_config(args.workload, plan, S),
            "useful_tflops": work["ref_flops_per_slice"] * value / 1e12,
            "flops_per_slice": work["ref_flops_per_slice"], "bytes_per_slice": work["ref_bytes_per_slice"],
            "extrapolated_full_task_seconds": (2.0 ** plan.n_sliced) / value,
            "e2e": e2e, "gpu_launches": launches, "fused_amax_operands": plan.fused_amax_operands(), "clocks": clocks, "roofline": roofline,
            "slice_roofline": slice_roofline, "breakdown": breakdown,
            "cpu_baseline": cpu, "half_mode": half, "slice_reuse": reuse,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
