"""CPU ORACLE (torch flavour) -- test / baseline infrastructure, NOT product code.

The reference's executor is a thin loop around `torch.einsum` (third-party, not vendored:
artensor/contraction.py:70 and :147-190 are the call sites; the reference pins
pytorch==1.12.1 in examples/requirements.txt:7, this image ships 2.11.0).  This module restates
that loop with the same library call on CPU tensors, so that `bench.py`'s `cpu_baseline` leg and
`bench.py --impl reference` time what the reference itself would execute on the host cores
(`/root/reference` does not exist on the GPU box, so the reference package cannot be imported
there).  Only `tests/`, `__graft_entry__.smoke()` and those two bench legs may import it.

PINNING: `tests/test_oracle.py::test_torch_oracle_matches_reference_goldens` compares it with
the outputs of the real reference recorded in tests/golden/*.expected.npz.

  run_normal   <- tensor_contraction          artensor/contraction.py:62-76
  run_sparse   <- tensor_contraction_sparse   artensor/contraction.py:132-205
  slice_stepper <- one iteration of the slice loop, simulation.py:107-114, step by step
  estimate_slice_seconds: timing model on synthetic operands (no reference counterpart; a cross-check)
"""
import time

import torch


def run_normal(tensors, scheme):
    """contraction.py:62-76.  `tensors` (dict or list of CPU tensors) is mutated."""
    out = None
    for step in scheme:
        (i, j), eq = step[0], step[1]
        tensors[i] = out = torch.einsum(eq, tensors[i], tensors[j])
    return out


def run_sparse(tensors, scheme):
    """contraction.py:132-205 without the optional rescaling: chunked batched steps are gathered
    chunk by chunk and concatenated (:140-175), single-chunk batched steps gather both sides
    (:176-179), outer steps merge their two row modes by reshape and may keep a row subset
    (:180-188), everything else is a plain einsum (:189-191)."""
    out = None
    for step in scheme:
        (i, j), eq, (rows_i, rows_j) = step[0], step[1], step[2]
        a, b = tensors[i], tensors[j]
        five = len(step) > 3
        if len(rows_i) > 1:
            parts = [torch.einsum(eq, a[rows_i[c]], b[rows_j[c]]) for c in range(len(rows_i))]
            if step[3]:
                parts = [p.reshape(step[3]) for p in parts]
            out = torch.cat(parts, dim=0)
        elif five and len(rows_i) == 1 and len(rows_j) == 1:
            out = torch.einsum(eq, a[rows_i[0]], b[rows_j[0]])
        elif five:
            out = torch.einsum(eq, a, b).reshape(step[3])
            if len(rows_i) == 1:
                out = out[rows_i[0]]
        else:
            out = torch.einsum(eq, a, b)
        tensors[i] = out
        tensors[j] = []
    return out


def contract_slices(case, slice_ids, dtype=torch.complex64):
    """simulation.py:101-114 with corrected multi-bond slicing (see tn_oracle.slice_leaves)."""
    from artensor_b200.cases import slice_leaves
    func = run_normal if case.pattern == "normal" else run_sparse
    leaves = {k: v.to(dtype) for k, v in case.leaves.items()}
    sidx = case.slicing_indices()
    acc = None
    for s in slice_ids:
        r = func(slice_leaves(leaves, case.slicing_bonds, sidx, int(s)), case.scheme)
        acc = r.clone() if acc is None else acc + r
    return acc


def slice_stepper(case, slice_id, dtype=torch.complex64):
    """ONE true slice of `case` (its real leaves, every scheme step, exactly what
    `tensor_contraction[_sparse]` would execute for that slice: simulation.py:107-114), cut into
    single scheme steps: a generator that executes step k when advanced and yields
    (k, seconds of that step); after the last step it yields (None, result tensor).  The leaf
    slicing (`select(...).clone()`, simulation.py:110-113) is charged to step 0.  `bench.py --impl
    reference` spreads the steps of one slice over its timed bench steps with this."""
    from artensor_b200.cases import slice_leaves
    sparse = case.pattern != "normal"
    t0 = time.perf_counter()
    leaves = {k: v.to(dtype) for k, v in case.leaves.items()}
    tensors = slice_leaves(leaves, case.slicing_bonds, case.slicing_indices(), int(slice_id))
    out = None
    for k, step in enumerate(case.scheme):
        if k > 0:
            t0 = time.perf_counter()
        out = (run_sparse if sparse else run_normal)(tensors, [step])
        yield k, time.perf_counter() - t0
    yield None, out


# --------------------------------------------------------------------------- timing helpers
def _shrunk_step(eq, shape_a, shape_b, max_elems):
    """Shrink one einsum step until none of its three tensors exceeds `max_elems` elements, by
    halving the extent of one label at a time (work of a pairwise einsum is linear in every
    label's extent).  The label is always taken from the currently largest tensor, preferring a
    shared row label, then a kept (left-/right-only) label, then a contracted one.  Returns
    (eq', shape_a', shape_b', scale) with scale = full work / shrunk work."""
    lhs, lo = eq.split("->")
    la, lb = lhs.split(",")
    ext = dict(zip(la, shape_a))
    ext.update(zip(lb, shape_b))
    la, lb, lo = list(la), list(lb), list(lo)

    def numel(labels):
        n = 1
        for l in labels:
            n *= ext[l]
        return n

    scale = 1.0
    while True:
        sizes = [(numel(la), la), (numel(lb), lb), (numel(lo), lo)]
        big, labels = max(sizes, key=lambda x: x[0])
        if big <= max_elems:
            break
        cands = [l for l in labels if ext[l] > 1]
        if not cands:
            break

        def prio(l):
            in_a, in_b, in_o = l in la, l in lb, l in lo
            if in_a and in_b and in_o:
                return 0
            if in_o:
                return 1
            return 2
        l = min(cands, key=lambda l: (prio(l), -ext[l]))
        new = max(1, ext[l] // 2)
        scale *= ext[l] / new
        ext[l] = new
    return ("".join(la) + "," + "".join(lb) + "->" + "".join(lo), [ext[l] for l in la], [ext[l] for l in lb], scale)


def estimate_slice_seconds(step_shapes, max_elems=1 << 22, repeats=1, budget_s=60.0):
    """Estimated CPU seconds for ONE slice of a scheme.

    `step_shapes` is a list of (eq, shape_a, shape_b): the einsum of every step with the shapes
    its operands have (gathered row counts included).  Each step runs `torch.einsum` on
    synthetic operands of exactly that shape; steps whose tensors exceed `max_elems` elements
    run on a sub-block (leading left-only modes fixed to 0) and their time is multiplied by the
    number of such sub-blocks -- einsum work is linear in those modes.  Returns
    (seconds, measured_steps, scaled_steps, wall_seconds)."""
    gen = torch.Generator().manual_seed(0)
    total = 0.0
    scaled = 0
    t_start = time.perf_counter()
    for eq, sa, sb in step_shapes:
        eq2, sa2, sb2, scale = _shrunk_step(eq, sa, sb, max_elems)
        a = torch.randn(sa2 + [2], generator=gen)
        b = torch.randn(sb2 + [2], generator=gen)
        a, b = torch.view_as_complex(a), torch.view_as_complex(b)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            torch.einsum(eq2, a, b)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        total += best * scale
        scaled += scale > 1.0
        if time.perf_counter() - t_start > budget_s:
            raise RuntimeError("cpu baseline sample exceeded its time budget")
    return total, len(step_shapes), scaled, time.perf_counter() - t_start
