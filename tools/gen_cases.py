#!/usr/bin/env python
"""Generate case files + golden outputs by RUNNING the reference (import only).

Runs in the build container only (needs /root/reference on PYTHONPATH); never on the GPU
box.  For every named configuration it
  1. builds the circuit tensor network with the reference's own circuit builder,
  2. runs the reference order search + scheme compiler (`prepare_contraction`),
  3. validates the scheme (chunk coverage),
  4. freezes leaves + scheme + slicing info into tests/golden/<name>.case.gz,
  5. runs the REFERENCE executor (`artensor.contraction.tensor_contraction[_sparse]`) on CPU
     over a fixed list of slice ids in complex64 (the parity oracle) and complex128
     (truth), with shift-corrected leaf slicing, and stores the results in
     tests/golden/<name>.expected.npz.

Usage:  PYTHONHASHSEED=0 python tools/gen_cases.py <config> [<config> ...]
"""
import os
import re
import sys
import time

import numpy as np

REF = os.environ.get("ARTENSOR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402
import artensor  # noqa: E402  (the reference)
from artensor import TensorNetworkSimulation  # noqa: E402
from artensor.contraction import tensor_contraction, tensor_contraction_sparse  # noqa: E402

from artensor_b200.cases import save_case, load_case, slice_leaves  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GOLD = os.path.join(ROOT, "tests", "golden")
TMP = os.environ.get("TNC_TMP", "/tmp/tnc_cases")
os.makedirs(TMP, exist_ok=True)
os.makedirs(GOLD, exist_ok=True)


# --------------------------------------------------------------------------- circuits
def n12_qsim():
    return os.path.join(REF, "tests", "circuit_n12_m14_s0_e0_pEFGH.qsim")


def n30_qsim():
    """examples/circuit_n30_m14_s0_e0_pEFGH.py is a cirq script; cirq is not installed and
    the package's parser only reads .qsim (`circuit.py:48-73`).  Text-convert it; the gate
    mapping is 1:1."""
    out = os.path.join(TMP, "circuit_n30_m14_s0_e0_pEFGH.qsim")
    if os.path.exists(out):
        return out
    src = open(os.path.join(REF, "examples", "circuit_n30_m14_s0_e0_pEFGH.py")).read()
    head, body = src.split("CIRCUIT = cirq.Circuit(")
    order = re.findall(r"cirq\.GridQubit\((\d+), (\d+)\)", head)
    qid = {q: i for i, q in enumerate(order)}
    moments = body.split("cirq.Moment(operations=[")[1:]
    lines = [str(len(order))]
    gq = r"cirq\.GridQubit\((\d+),\s*(\d+)\)"
    tok = re.compile(
        r"cirq\.PhasedXPowGate\(phase_exponent=0\.25,\s*exponent=0\.5\)\.on\(" + gq + r"\)"
        r"|\(cirq\.X\*\*0\.5\)\.on\(" + gq + r"\)"
        r"|\(cirq\.Y\*\*0\.5\)\.on\(" + gq + r"\)"
        r"|cirq\.Rz\(np\.pi \* ([-0-9.e]+)\)\.on\(" + gq + r"\)"
        r"|cirq\.FSimGate\(theta=([-0-9.e]+),\s*phi=([-0-9.e]+)\)\.on\(\s*" + gq + r",\s*" + gq + r"\)"
    )
    for layer, m in enumerate(moments):
        for g in tok.finditer(m):
            v = g.groups()
            if v[0] is not None:
                lines.append(f"{layer} hz_1_2 {qid[(v[0], v[1])]}")
            elif v[2] is not None:
                lines.append(f"{layer} x_1_2 {qid[(v[2], v[3])]}")
            elif v[4] is not None:
                lines.append(f"{layer} y_1_2 {qid[(v[4], v[5])]}")
            elif v[6] is not None:
                lines.append(f"{layer} rz {qid[(v[7], v[8])]} {repr(np.pi * float(v[6]))}")
            else:
                lines.append(f"{layer} fs {qid[(v[11], v[12])]} {qid[(v[13], v[14])]} {v[9]} {v[10]}")
    assert len(moments) == 57 and len(lines) == 1271, (len(moments), len(lines))
    open(out, "w").write("\n".join(lines) + "\n")
    return out


def n53_qsim(m):
    src = os.path.join(REF, "examples", "circuits", "circuit_n53_m20_s0_e0_pABCDCDAB.qsim")
    if m == 20:
        return src
    out = os.path.join(TMP, f"circuit_n53_m{m}_trunc.qsim")
    last = 4 * m  # every cycle = 4 layers; layer 4m is a full single-qubit layer
    with open(src) as f, open(out, "w") as g:
        first = f.readline()
        g.write(first)
        for line in f:
            if line.strip() and int(line.split()[0]) <= last:
                g.write(line)
    return out


def google_amplitudes(k):
    rows = open(os.path.join(REF, "examples", "amplitudes_n30_m14_s0_e0_pEFGH_10000.txt")).read().split("\n")
    rows = [r.split() for r in rows if r.strip()][:k]
    return [r[0] for r in rows], np.array([float(r[1]) + 1j * float(r[2]) for r in rows])


def correlated_bitstrings(n, k, seed=0):
    """2^k bitstrings: k open qubit positions take all combinations, the rest are fixed."""
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 2, size=n)
    open_pos = np.sort(rng.choice(n, size=k, replace=False))
    out = []
    for v in range(1 << k):
        b = base.copy()
        for t, p in enumerate(open_pos):
            b[p] = (v >> (k - 1 - t)) & 1
        out.append("".join(map(str, b)))
    return out


def random_bitstrings(n, count, seed=0):
    rng = np.random.RandomState(seed)
    seen = set()
    while len(seen) < count:
        seen.add("".join(map(str, rng.randint(0, 2, size=n))))
    return sorted(seen)


KAT_N12 = {  # tests/test_circuits.py:25-31
    "100001000001": 0.0198028199 + 1j * (0.0106442748),
    "000101111011": 0.00497586094 + 1j * (-0.0245072283),
    "011000101100": -0.00853562169 + 1j * (-0.00701293815),
    "111001100001": -0.0100137182 + 1j * (0.0147468708),
    "001110110000": 0.00681955926 + 1j * (0.0106616206),
}

CONFIGS = {
    # name: (circuit fn, bitstrings fn, prepare kwargs, slice ids to run (None = all), run c128?)
    "n12_full": (n12_qsim, lambda: [], dict(sc_target=30, trials=4, iters=10), None, True),
    "n12_sparse5": (n12_qsim, lambda: list(KAT_N12), dict(sc_target=30, trials=4, iters=10), None, True),
    "n12_sparse64_sc9": (n12_qsim, lambda: random_bitstrings(12, 64, 1), dict(sc_target=9, trials=4, iters=10), None, True),
    "n12_sparse100_sc8": (n12_qsim, lambda: random_bitstrings(12, 100, 2), dict(sc_target=8, trials=4, iters=10), None, True),
    "n12_sparse256c_sc10": (n12_qsim, lambda: correlated_bitstrings(12, 8, 3), dict(sc_target=10, trials=4, iters=10), None, True),
    # "_own": the scheme is compiled from the reference's contraction tree by artensor_b200/scheme.py
    # (SURVEY.md 8-f1); the expected outputs are still the REFERENCE executor's, run on that scheme
    "n12_full_own": (n12_qsim, lambda: [], dict(sc_target=30, trials=4, iters=10), None, True),
    "n12_sparse100_sc8_own": (n12_qsim, lambda: random_bitstrings(12, 100, 2), dict(sc_target=8, trials=4, iters=10), None, True),
    # leaves built in float64 then cast (from_circuit_file(leaf_precision="double")); the case
    # also records the float64 state-vector amplitudes of the bitstrings (TensorNetworkCircuit.state_vec)
    "n12_sparse5_f64leaves": (n12_qsim, lambda: list(KAT_N12), dict(sc_target=30, trials=4, iters=10), None, True),
    "n12_sparse64_sc9_f64leaves": (n12_qsim, lambda: random_bitstrings(12, 64, 1), dict(sc_target=9, trials=4, iters=10), None, True),
    "n30_sparse64_sc26": (n30_qsim, lambda: google_amplitudes(64)[0], dict(sc_target=26, trials=4, iters=5), None, False),
    "n30_full": (n30_qsim, lambda: [], dict(sc_target=30, trials=4, iters=5), [0], False),
    # BASELINE config 2 sharded over its first 3 output qubits: 8 shards x 4 regular
    # slices; slice id = (shard, regular slice), the expected outputs are the reference executor's on that scheme
    "n30_full_shard3": (n30_qsim, lambda: [], dict(sc_target=30, trials=4, iters=5), [0, 9, 22, 31], False),
    "n30_sparse10000": (n30_qsim, lambda: google_amplitudes(10000)[0], dict(sc_target=30, trials=4, iters=5), [0], False),
    # BASELINE config 3 as a SLICED contraction (the unsliced scheme above is one slice: nothing to spread over GPUs)
    "n30_sparse10000_sc27": (n30_qsim, lambda: google_amplitudes(10000)[0], dict(sc_target=27, trials=4, iters=5), [0, 3], False),
    "n53_m12_sparse1024": (lambda: n53_qsim(12), lambda: correlated_bitstrings(53, 10, 0), dict(sc_target=30, trials=4, iters=3), [0, 1, 5], False),
    "n53_m20_sparse1024": (lambda: n53_qsim(20), lambda: correlated_bitstrings(53, 10, 0), dict(sc_target=30, trials=4, iters=3), [0], False),
}


def google_extra(name, bitstrings):
    """Google's amplitudes for the requested bitstrings (n30 cases only): golden vectors."""
    if not name.startswith("n30_sparse"):
        return {}
    strings, amps = google_amplitudes(10000)
    table = dict(zip(strings, amps))
    return {"google_amplitudes": [complex(table[b]) for b in bitstrings]}


def statevector_extra(name, circ_fn, bitstrings):
    """float64 state-vector amplitudes of the requested bitstrings (n12 "_f64leaves" cases): the truth
    no complex64 pipeline can be better than."""
    if not name.endswith("_f64leaves"):
        return {}
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        sv = artensor.TensorNetworkCircuit(circ_fn(), dtype=torch.complex128).state_vec().reshape(-1)
    finally:
        torch.set_default_dtype(old)
    return {"statevector_f64": {b: complex(sv[int(b, 2)]) for b in bitstrings}}


def validate_scheme(scheme, pattern):
    """B2: every chunked step must cover exactly next_shape[0] rows, no empty chunk."""
    if pattern != "sparse":
        return
    for k, step in enumerate(scheme):
        bi, bj = step[2]
        if len(bi) > 1:
            tot = sum(len(c) for c in bi)
            if any(len(c) == 0 for c in bi) or tot != step[4][0] or [len(c) for c in bi] != [len(c) for c in bj]:
                raise RuntimeError(f"step {k}: invalid chunking (B2): {[len(c) for c in bi]} vs rows {step[4][0]}")


def build(name):
    circ_fn, bits_fn, prep, slice_ids, want128 = CONFIGS[name]
    case_path = os.path.join(GOLD, f"{name}.case.gz")
    bitstrings = bits_fn()
    if not os.path.exists(case_path):
        t0 = time.time()
        shard_bits = int(name.rsplit("_shard", 1)[1]) if "_shard" in name else 0
        f64 = name.endswith("_f64leaves")
        if name.endswith("_own") or shard_bits or f64:
            from artensor_b200 import TensorNetworkSimulation as OwnSimulation
            sim = OwnSimulation.from_circuit_file(circ_fn(), bitstrings, leaf_precision="double" if f64 else "single")
            assert sim.scheme_compiler == "b200"
        else:
            sim = TensorNetworkSimulation.from_circuit_file(circ_fn(), bitstrings)
        sim.prepare_contraction(slicing_repeat=1, start_seed=0, **prep)
        if shard_bits:
            sim.prepare_open_qubit_shards(shard_bits)
        validate_scheme(sim.scheme, sim.pattern)
        print(f"[{name}] order search {time.time() - t0:.1f}s; steps={len(sim.scheme)} "
              f"slicing_bonds={len(sim.slicing_indices)}", flush=True)
        save_case(
            case_path, name=name, pattern=sim.pattern, leaves=sim.tensors, leaf_bonds=sim.tensor_bonds,
            scheme=sim.scheme, slicing_bonds=list(sim.slicing_indices.keys()),
            output_bonds=sim.output_bonds,
            permute_dims=getattr(sim, "permute_dims", None) if len(sim.output_bonds) else None,
            bitstrings_sorted=getattr(sim, "bitstrings_sorted", None),
            n_qubits=len(sim.final_qubits),
            extra={"prepare": {k: v for k, v in prep.items()}, "bitstrings_in": bitstrings, "n_shard_bonds": shard_bits,
                   **google_extra(name, bitstrings), **statevector_extra(name, circ_fn, bitstrings),
                   "ref_slicing_indices": {b: [(int(t), int(d)) for t, d in v] for b, v in sim.slicing_indices.items()}},
        )
    case = load_case(case_path)
    return case


def run_reference(case, slice_ids, dtype):
    func = tensor_contraction if case.pattern == "normal" else tensor_contraction_sparse
    sidx = case.slicing_indices()
    leaves = {k: v.to(dtype) for k, v in case.leaves.items()}
    outs = []
    for s in slice_ids:
        t0 = time.time()
        sl = slice_leaves(leaves, case.slicing_bonds, sidx, s)
        res = func(sl, case.scheme)
        outs.append(res)
        print(f"   slice {s} [{dtype}] {time.time() - t0:.2f}s shape={tuple(res.shape)}", flush=True)
    return outs


def sample_indices(numel, count=8192, seed=7):
    rng = np.random.RandomState(seed)
    return np.unique(rng.randint(0, numel, size=count, dtype=np.int64))


def golden(name):
    case = build(name)
    _, _, _, slice_ids, want128 = CONFIGS[name]
    exp_path = os.path.join(GOLD, f"{name}.expected.npz")
    if os.path.exists(exp_path):
        print(f"[{name}] expected exists")
        return
    if slice_ids is None:
        slice_ids = list(range(case.n_slices))
    torch.set_num_threads(os.cpu_count())
    out = {"slice_ids": np.array(slice_ids, dtype=np.int64)}
    res64 = run_reference(case, slice_ids, torch.complex64)
    numel = res64[0].numel()
    big = numel > (1 << 22)
    if big:  # store a fixed sample of entries + the squared norm (cannot commit 8 GiB)
        # entries are sampled in the executor's own output order (before permute_dims)
        idx = sample_indices(numel)
        out["sample_idx"] = idx
        out["per_slice_c64"] = np.stack([r.reshape(-1)[torch.from_numpy(idx)].numpy() for r in res64])
        out["per_slice_norm2"] = np.array([float((r.abs().double() ** 2).sum()) for r in res64])
        if name == "n30_full" and case.n_slices > 1:
            pass
    else:
        out["per_slice_c64"] = np.stack([r.contiguous().reshape(-1).numpy() for r in res64])
    out["shape"] = np.array(res64[0].shape, dtype=np.int64)
    del res64
    if want128:
        res128 = run_reference(case, slice_ids, torch.complex128)
        out["per_slice_c128"] = np.stack([r.contiguous().reshape(-1).numpy() for r in res128])
    np.savez_compressed(exp_path, **out)
    print(f"[{name}] wrote {exp_path}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for nm in names:
        golden(nm)
