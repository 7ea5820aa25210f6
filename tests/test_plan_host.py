"""Host logic of the plan compiler, checked on CPU: the operation records ContractionPlan
emits are executed by a numpy interpreter (tests/emulate.py) and compared with the oracle
and with the reference's golden outputs."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from artensor_b200 import SchemeError, partition_slices
from artensor_b200 import _native as N
from artensor_b200.backend import Arena, ContractionPlan, PlanOptions
from artensor_b200.plan import SchemeParser, parse_eq
from oracle import tn_oracle as O
import emulate

SMALL = ["n12_full", "n12_sparse5", "n12_sparse64_sc9", "n12_sparse100_sc8", "n12_sparse256c_sc10",
         "n12_full_own", "n12_sparse100_sc8_own"]   # _own: scheme compiled by artensor_b200/scheme.py


def make_plan(case, **kw):
    return ContractionPlan(case.scheme, {k: tuple(v.shape) for k, v in case.leaves.items()}, case.pattern == "sparse",
                           slicing_bonds=case.slicing_bonds, slicing_indices=case.slicing_indices(),
                           build_native=False, **kw)


@pytest.mark.parametrize("name", SMALL)
def test_lowered_plan_matches_reference(name):
    case, exp = load_golden(name)
    plan = make_plan(case)
    blob = plan.pack_leaves(case.leaves).numpy()
    ids = exp["slice_ids"]
    pick = list(ids[:3]) + [int(ids[-1])]
    for s in dict.fromkeys(int(x) for x in pick):
        k = int(np.where(ids == s)[0][0])
        got = emulate.run_plan(plan, blob, [s]).reshape(-1)
        want = exp["per_slice_c64"][k]
        assert np.abs(got - want).max() / np.abs(want).max() < 5e-6


@pytest.mark.parametrize("name", ["n12_sparse64_sc9", "n12_sparse256c_sc10"])
def test_slice_sum_and_hoisting(name):
    """Summing a slice range inside one execute == summing per-slice oracle results; hoisted
    (slice-invariant) steps give the same numbers as recomputing them per slice."""
    case, exp = load_golden(name)
    ids = list(range(5, 13))
    want = O.contract_slices(case.leaves, case.scheme, case.pattern, case.slicing_bonds, case.slicing_indices(), ids)
    for hoist in (True, False):
        plan = make_plan(case, options=PlanOptions(hoist=hoist))
        if not hoist:
            assert plan.work_summary()["hoisted_steps"] == 0
        got = emulate.run_plan(plan, plan.pack_leaves(case.leaves).numpy(), ids)
        assert np.abs(got - want).max() / np.abs(want).max() < 5e-6


@pytest.mark.parametrize("hoist", [True, False])
@pytest.mark.parametrize("name", ["n12_sparse64_sc9", "n12_sparse256c_sc10", "n12_sparse100_sc8_own"])
def test_slice_reuse_is_bit_identical_to_recomputing_every_slice(name, hoist):
    """PlanOptions.slice_reuse: over a range of consecutive slice ids a step runs again only when a sliced bond
    behind it changed.  The emulator follows the library's rule (dependencies derived from the leaf records), and
    overwrites all recycled memory with NaN after every slice: the sum must still be bit-identical to running every
    operation for every slice -- i.e. every result that is read in a later slice sits in the KEEP region."""
    case, _ = load_golden(name)
    n = 1 << len(case.slicing_bonds)
    full = make_plan(case, options=PlanOptions(hoist=hoist))
    plan = make_plan(case, options=PlanOptions(hoist=hoist, slice_reuse=True))
    assert plan.slice_reuse and plan.keep_bytes > 0 and plan.workspace_bytes >= full.workspace_bytes
    deps = emulate.slice_deps(plan)
    steps = [st for st in plan.op_steps[1] if st is not None]
    assert sum(1 for d in deps if d != (1 << plan.n_sliced) - 1) > len(steps) // 4     # there IS something to reuse
    blob = plan.pack_leaves(case.leaves).numpy()
    for lo, hi in [(0, min(n, 24)), (5, 14), (n - 3, n)]:
        want = emulate.run_plan(full, blob, range(lo, hi))
        got = emulate.run_plan(plan, blob, range(lo, hi), reuse=True, poison=True)
        assert np.array_equal(got, want), f"slices [{lo}, {hi})"


@pytest.mark.parametrize("name", ["n12_sparse64_sc9", "n12_sparse256c_sc10"])
def test_slice_reuse_with_a_keep_budget(name):
    """PlanOptions.keep_budget_bytes: when the results that must outlive a slice exceed the budget, steps are TIED to
    their reader (TNC_EINSUM_RUN_WITH_READER: they run whenever the reader does, their result recycles).  Every budget
    down to zero must give the bit-identical sum (emulator with all recycled memory poisoned after every slice), the
    KEEP region must respect the budget, and the modelled cost must grow monotonically towards contracting every step
    for every slice."""
    case, _ = load_golden(name)
    n = 1 << len(case.slicing_bonds)
    full = make_plan(case)
    free = make_plan(case, options=PlanOptions(slice_reuse=True))
    blob = free.pack_leaves(case.leaves).numpy()
    lo, hi = 3, min(n, 40)
    want = emulate.run_plan(full, blob, range(lo, hi))
    last = free.reuse_summary()["amortised_s"]
    for budget in (free.keep_bytes, free.keep_bytes // 2, free.keep_bytes // 8, 0):
        plan = make_plan(case, options=PlanOptions(slice_reuse=True, keep_budget_bytes=budget))
        assert plan.keep_bytes <= max(budget, 0) or budget >= free.keep_bytes
        assert (sum(plan.step_tied) > 0) == (budget < free.keep_bytes)
        cost = plan.reuse_summary()
        assert cost["amortised_s"] >= last * (1 - 1e-12) and cost["amortised_s"] <= cost["full_s"] * (1 + 1e-12)
        last = cost["amortised_s"]
        got = emulate.run_plan(plan, blob, range(lo, hi), reuse=True, poison=True)
        assert np.array_equal(got, want), f"budget {budget}"
    assert plan.keep_bytes == 0 and abs(last - cost["full_s"]) <= 1e-9 * cost["full_s"]      # budget 0: nothing is reused


@pytest.mark.parametrize("name", ["n53_m12_sparse1024", "n53_m20_sparse1024"])
def test_keep_planning_invariants_on_the_n53_trees(name):
    """The planner's KEEP decisions on the big trees (host only): every slice-phase result is either kept or runs
    exactly when its reader does (equal lowest slice-id bits of their run dependencies); run dependencies contain the
    step's own; a budget is respected, ties only ever ADD modelled time, and the library's own finalize check accepts
    the layout (it fails later, in cudaMalloc, on a machine without a GPU)."""
    import torch
    from artensor_b200 import _native as N
    case, _ = load_golden(name)
    S = len(case.slicing_bonds)
    free = make_plan(case, options=PlanOptions(slice_reuse=True))
    low = lambda d: min((S - 1 - b for b in d), default=None)
    for budget in (None, free.keep_bytes // 3, 0):
        plan = make_plan(case, options=PlanOptions(slice_reuse=True, keep_budget_bytes=budget))
        every = frozenset(range(S))
        kept_bytes = 0
        for i, st in enumerate(plan.steps):
            c = plan._consumer_of[i]
            reader = plan.run_deps[c] if c is not None else every
            assert plan.step_deps[i] <= plan.run_deps[i]
            if plan.step_phase[i] != N.TNC_PHASE_SLICE:
                continue
            if low(plan.run_deps[i]) != low(reader):
                kept_bytes += max(1024, (st.c.numel * 8 + 1023) // 1024 * 1024)
            if plan.step_tied[i]:
                assert plan.run_deps[i] == reader
        assert kept_bytes == plan.keep_bytes
        if budget is not None:
            assert plan.keep_bytes <= budget
        cost = plan.reuse_summary()
        assert free.reuse_summary()["amortised_s"] * (1 - 1e-12) <= cost["amortised_s"] <= cost["full_s"] * (1 + 1e-12)
        plan.slice_reuse = True
        try:
            plan._build_native(plan.ops)
            status = 0
        except N.NativeError as e:
            status = e.status
        assert status == (0 if torch.cuda.is_available() else 2), budget


def test_fit_reuse_to_memory_shrinks_the_keep_budget():
    """TensorNetworkSimulation.fit_reuse_to_memory: the KEEP budget shrinks until the reuse plan's workspace fits the
    free memory it is told about; with room to spare nothing is tied."""
    from artensor_b200 import TensorNetworkSimulation
    import artensor_b200.contraction as C
    case, _ = load_golden("n53_m12_sparse1024")
    sim = TensorNetworkSimulation.from_case(case)
    sim.plan_options = PlanOptions(slice_reuse=True)
    real = C.ContractionPlan
    C.ContractionPlan = lambda *a, **k: real(*a, **dict(k, build_native=False))      # host-side planning only
    try:
        sim.optimize_slice_order()               # the reuse order keeps ~1.8 GiB of results across slices
        roomy = sim.fit_reuse_to_memory(free_bytes=64 << 30)
        assert sum(roomy.step_tied) == 0 and sim.plan_options.keep_budget_bytes is None
        tight = sim.fit_reuse_to_memory(free_bytes=roomy.workspace_bytes - (1 << 29), margin=0)
        assert tight.workspace_bytes <= roomy.workspace_bytes - (1 << 29) and sum(tight.step_tied) > 0
        assert tight.keep_bytes < roomy.keep_bytes and sim.plan_options.keep_budget_bytes is not None
    finally:
        C.ContractionPlan = real


def test_slice_reuse_layout_check_of_the_library():
    """tnc_plan_finalize checks what TNC_OPT_SLICE_REUSE asks of the layout on the host, before its first CUDA
    call: the planner's layout passes it (on a machine without a GPU finalize then fails in cudaMalloc: status
    CUDA), the layout of a plan made WITHOUT the option, executed with it, is refused (status INVALID)."""
    import ctypes as C
    import torch
    from artensor_b200 import _native as N
    case, _ = load_golden("n12_sparse64_sc9")
    shapes = {k: tuple(v.shape) for k, v in case.leaves.items()}

    def finalize_status(layout_reuse, budget=None):
        plan = make_plan(case, options=PlanOptions(slice_reuse=layout_reuse, keep_budget_bytes=budget))
        plan.slice_reuse = True                  # the option the library sees
        try:
            plan._build_native(plan.ops)
        except N.NativeError as e:
            return e.status
        return 0
    ok = finalize_status(True)
    assert ok == (0 if torch.cuda.is_available() else 2)
    assert finalize_status(False) == 1
    # with a keep budget the planner ties steps to their readers (TNC_EINSUM_RUN_WITH_READER): the library derives the
    # same run conditions from the flags and accepts the smaller KEEP region
    for budget in (8192, 0):
        assert finalize_status(True, budget) == ok


def test_work_summary_counts():
    case, _ = load_golden("n12_sparse64_sc9")
    w = make_plan(case).work_summary()
    assert w["steps"] == 68 and 0 < w["hoisted_steps"] < 68
    assert w["exec_flops_per_slice"] < w["ref_flops_per_slice"]


def test_multi_sliced_leaf_uses_corrected_dims():
    """A leaf with two sliced bonds: all four settings must match numpy indexing on the
    un-sliced tensor (the packaged reference loop is off by one here)."""
    rng = np.random.RandomState(0)
    t0 = torch.from_numpy((rng.randn(2, 2, 2, 2) + 1j * rng.randn(2, 2, 2, 2)).astype(np.complex64))
    t1 = torch.from_numpy((rng.randn(2, 2) + 1j * rng.randn(2, 2)).astype(np.complex64))
    scheme = [((0, 1), "ab,bc->ac")]      # after slicing dims 1 and 3 of t0 it is rank 2
    sl = {"x": [(0, 1)], "y": [(0, 3)]}
    plan = ContractionPlan(scheme, {0: (2, 2, 2, 2), 1: (2, 2)}, False, slicing_bonds=["x", "y"],
                           slicing_indices=sl, build_native=False)
    blob = plan.pack_leaves({0: t0, 1: t1}).numpy()
    for s in range(4):
        bx, by = (s >> 1) & 1, s & 1
        want = t0.numpy()[:, bx, :, by] @ t1.numpy()
        got = emulate.run_plan(plan, blob, [s])
        assert np.allclose(got, want, atol=1e-6)


def test_parser_rejects_bad_schemes():
    with pytest.raises(SchemeError):
        parse_eq("ab,bc,cd->ad")
    with pytest.raises(SchemeError):
        parse_eq("aab,bc->ac")
    p = SchemeParser({0: (2, 2), 1: (2, 2)}, sparse=False)
    with pytest.raises(SchemeError, match="summed inside one operand"):
        p.parse([((0, 1), "ab,bc->c")])
    p = SchemeParser({0: (2, 3), 1: (2, 2)}, sparse=False)
    with pytest.raises(SchemeError, match="extent 2"):
        p.parse([((0, 1), "ab,bc->ac")])
    p = SchemeParser({0: (2, 2), 1: (2, 2)}, sparse=False)
    with pytest.raises(SchemeError, match="not among the leaves"):
        p.parse([((0, 5), "ab,bc->ac")])
    p = SchemeParser({0: (2, 2), 1: (2, 2), 2: (2, 2)}, sparse=False)
    with pytest.raises(SchemeError, match="already consumed"):
        p.parse([((0, 1), "ab,bc->ac"), ((2, 1), "ab,bc->ac")])


def test_parser_detects_invalid_chunking():
    """B2: index lists that do not cover the announced row count are refused up front."""
    a = torch.zeros(4, 2, dtype=torch.complex64)
    b = torch.zeros(4, 2, dtype=torch.complex64)
    good = [((0, 1), "ab,ab->a", [[torch.tensor([0, 1])], [torch.tensor([2, 3])]], None, (2,))]
    SchemeParser({0: (4, 2), 1: (4, 2)}, True).parse(good)
    bad = [((0, 1), "ab,ab->a", [[torch.tensor([0, 1])], [torch.tensor([2, 3])]], None, (3,))]
    with pytest.raises(SchemeError, match="invalid chunking"):
        SchemeParser({0: (4, 2), 1: (4, 2)}, True).parse(bad)


def test_arena_reuses_and_coalesces():
    a = Arena()
    o1, s1 = a.alloc(1000)
    o2, s2 = a.alloc(5000)
    o3, s3 = a.alloc(100)
    assert (o1, o2, o3) == (0, 1024, 1024 + 5120)
    a.release(o1, s1)
    a.release(o2, s2)
    o4, s4 = a.alloc(6000)        # fits the coalesced hole
    assert o4 == 0 and a.high == 1024 + 5120 + 1024
    a.release(o3, s3)
    o5, _ = a.alloc(10)
    assert o5 in (6144, 7168)


def test_partition_slices_covers_range_exactly():
    for n, w in [(8, 2), (7, 3), (1, 8), (0, 4), (1 << 23, 8)]:
        parts = [partition_slices(3, 3 + n, r, w) for r in range(w)]
        assert parts[0][0] == 3 and parts[-1][1] == 3 + n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def _dist_worker(rank, world, port, name, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case, _ = load_golden(name)
    plan = make_plan(case)
    lo, hi = partition_slices(0, case.n_slices, rank, world)
    part = emulate.run_plan(plan, plan.pack_leaves(case.leaves).numpy(), range(lo, hi))
    t = torch.from_numpy(np.ascontiguousarray(part))
    dist.all_reduce(torch.view_as_real(t), op=dist.ReduceOp.SUM)   # same call the product makes over NCCL
    if rank == 0:
        out_q.put(t.numpy())
    dist.destroy_process_group()


def test_two_rank_slice_partition_gloo():
    """world_size-2 gloo run of the multi-GPU recipe: block-partition the slice range, contract
    locally, one sum all-reduce of the partial amplitudes."""
    import torch.multiprocessing as mp
    name = "n12_sparse256c_sc10"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case, exp = load_golden(name)
    want = exp["per_slice_c128"].sum(axis=0)
    assert np.abs(got.reshape(-1) - want).max() / np.abs(want).max() < 5e-6


@pytest.mark.parametrize("name", ["n12_full", "n12_sparse64_sc9", "n12_sparse256c_sc10", "n12_sparse100_sc8"])
def test_tensor_core_layouts_match_reference(name):
    """Every eligible step lowered to the tensor-core form (C = [rows][m][n], scratch panels in
    the arena): the records still evaluate to the reference's result."""
    case, exp = load_golden(name)
    plan = make_plan(case, options=PlanOptions(tc_min_flops=0, tc_min_intensity=0))
    assert N.TNC_ALGO_TC in plan.step_algo
    blob = plan.pack_leaves(case.leaves).numpy()
    ids = exp["slice_ids"]
    for s in (int(ids[0]), int(ids[-1])):
        k = int(np.where(ids == s)[0][0])
        got = emulate.run_plan(plan, blob, [s]).reshape(-1)
        want = exp["per_slice_c64"][k]
        assert np.abs(got - want).max() / np.abs(want).max() < 5e-6


def test_tc_scratch_formula_matches_library():
    import ctypes as C
    from artensor_b200.backend import tc_scratch_bytes
    lib = N.load()
    case, _ = load_golden("n12_sparse256c_sc10")
    plan = make_plan(case, options=PlanOptions(tc_min_flops=0, tc_min_intensity=0))
    n = 0
    for ph in plan.ops:
        for (kind, rec), st in zip(plan.ops[ph], plan.op_steps[ph]):
            if kind == "einsum" and rec.algo == N.TNC_ALGO_TC:
                assert lib.tnc_einsum_tc_scratch_bytes(N.TNC_C64, C.byref(rec)) == tc_scratch_bytes(st) == rec.scratch_bytes
                n += 1
    assert n > 10


@pytest.mark.parametrize("name", ["n12_full", "n12_sparse64_sc9", "n12_sparse256c_sc10"])
def test_streaming_kernel_layouts_match_reference(name):
    """Every step lowered to the streaming kernel's form (C = [rows][m][n])."""
    case, exp = load_golden(name)
    plan = make_plan(case, options=PlanOptions(tc_min_flops=float("inf"), stem_min_elems=0))
    assert plan.step_algo.count(N.TNC_ALGO_STEM) > len(plan.steps) // 2
    blob = plan.pack_leaves(case.leaves).numpy()
    ids = exp["slice_ids"]
    s = int(ids[-1])
    k = int(np.where(ids == s)[0][0])
    got = emulate.run_plan(plan, blob, [s]).reshape(-1)
    want = exp["per_slice_c64"][k]
    assert np.abs(got - want).max() / np.abs(want).max() < 5e-6


def test_default_options_pick_algorithms_by_intensity():
    """n53 m20 (the bench workload): the fat GEMM goes to the pack + tcgen05 GEMM lowering, the
    stem to the streaming kernels (tcgen05 for 2 <= k <= 5, fp32 otherwise), the tiny steps to the
    generic kernel."""
    case, _ = load_golden("n53_m20_sparse1024")
    plan = make_plan(case)
    by = {a: [st for st, x in zip(plan.steps, plan.step_algo) if x == a] for a in (0, 1, 2, 3)}
    fat = max(plan.steps, key=lambda s: s.flops)
    assert fat in by[N.TNC_ALGO_TC] and fat.flops > 5e13
    assert all(s.flops >= 24 * s.bytes_c64 for s in by[N.TNC_ALGO_TC])
    streamed = by[N.TNC_ALGO_STEM] + by[N.TNC_ALGO_SKINNY]
    assert sum(s.bytes_c64 for s in streamed) > 0.7 * sum(s.bytes_c64 for s in plan.steps if s is not fat)
    for s in by[N.TNC_ALGO_SKINNY]:
        assert 2 <= len(s.k_modes) <= 6 and 1 <= len(s.n_modes) <= 7 and s.a.numel >= 1 << 20
    assert sum(s.flops for s in by[N.TNC_ALGO_SKINNY]) > 2 * sum(s.flops for s in by[N.TNC_ALGO_STEM])
    assert len(by[N.TNC_ALGO_SIMT]) > 100 and max(s.c.numel for s in by[N.TNC_ALGO_SIMT]) < 1 << 12


def test_operand_swap_keeps_results():
    """Steps whose right operand is the larger one are lowered with the roles exchanged
    (C^T = B^T A^T); the emitted records must still compute the reference's einsum, with and
    without rows on either operand."""
    from artensor_b200.backend import should_swap
    rng = np.random.RandomState(4)

    def rnd(*shape):
        return torch.from_numpy((rng.randn(*shape) + 1j * rng.randn(*shape)).astype(np.complex64))

    # plain, no rows: A rank 5, B rank 13
    eq = "abcxy,xydefghijklmn->dabecfghijklmn"
    leaves = {0: rnd(*[2] * 5), 1: rnd(*[2] * 13)}
    plan = ContractionPlan([((0, 1), eq)], {k: tuple(v.shape) for k, v in leaves.items()}, False, build_native=False)
    assert should_swap(plan.steps[0])
    rec = plan.ops[N.TNC_PHASE_ONCE][-1][1] if plan.ops[N.TNC_PHASE_ONCE][-1][0] == "einsum" else plan.ops[N.TNC_PHASE_SLICE][0][1]
    assert rec.a.rank == 13 and rec.b.rank == 5 and rec.n_m == 11 and rec.n_n == 3     # roles exchanged
    got = emulate.run_plan(plan, plan.pack_leaves(leaves).numpy(), [0]).reshape(plan.out_shape)
    want = np.einsum(eq, leaves[0].numpy().astype(np.complex128), leaves[1].numpy().astype(np.complex128))
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6
    no_swap = ContractionPlan([((0, 1), eq)], {k: tuple(v.shape) for k, v in leaves.items()}, False, build_native=False,
                              options=PlanOptions(swap_operands=False))
    got2 = emulate.run_plan(no_swap, no_swap.pack_leaves(leaves).numpy(), [0]).reshape(no_swap.out_shape)
    assert np.abs(got2 - want).max() / np.abs(want).max() < 2e-6
    # sparse, rows on the (larger) right operand only: 6 bitstring rows
    eq = "abx,Zxcdefghijkl->Zabcdefghijkl"
    leaves = {0: rnd(2, 2, 2), 1: rnd(6, *[2] * 11)}
    step = ((0, 1), eq, [[torch.tensor([0])], [torch.arange(6)]])      # contraction.py:249-254
    plan = ContractionPlan([step], {k: tuple(v.shape) for k, v in leaves.items()}, True, build_native=False)
    assert should_swap(plan.steps[0])
    got = emulate.run_plan(plan, plan.pack_leaves(leaves).numpy(), [0]).reshape(plan.out_shape)
    want = np.einsum(eq, leaves[0].numpy().astype(np.complex128), leaves[1].numpy().astype(np.complex128))
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6


def test_outer_pairs_detection_and_native_validation():
    """backend.outer_pairs: every (A row, B row) pair exactly once in any order but A-major."""
    from artensor_b200.backend import outer_pairs, full_outer
    rng = np.random.RandomState(2)
    RA, RB, m, k, n = 3, 4, 3, 2, 2
    L = "abcdefghijkl"
    la, lk, ln = L[:m], L[m:m + k], L[m + k:m + k + n]
    shapes = {0: (RA,) + (2,) * (m + k), 1: (RB,) + (2,) * (k + n)}
    eq = f"X{la}{lk},X{lk}{ln}->X{la}{ln}"

    def plan_for(ia, ib):
        step = ((0, 1), eq, [[torch.as_tensor(ia)], [torch.as_tensor(ib)]], None, tuple([len(ia)] + [2] * (m + n)))
        return ContractionPlan([step], shapes, True, build_native=False)
    order = rng.permutation(RA * RB)
    p = plan_for(order // RB, order % RB)
    assert outer_pairs(p.steps[0]) and not full_outer(p.steps[0])
    assert p.ops[N.TNC_PHASE_ONCE][-1][1].flags == N.TNC_EINSUM_OUTER_PAIRS
    got = emulate.run_plan(p, np.arange(p.leaf_blob_elems).astype(np.complex64), [0])   # the emulator checks the flag's contract
    assert got.shape[0] == RA * RB
    sub = order[:-1]
    assert not outer_pairs(plan_for(sub // RB, sub % RB).steps[0])                      # a pair missing
    dup = np.concatenate([order[:-1], order[:1]])
    assert not outer_pairs(plan_for(dup // RB, dup % RB).steps[0])                      # a pair twice



def test_plan_of_the_sliced_10000_amplitude_scheme_lowers():
    """n30 m14, 10000 bitstrings, sc_target 27 (512 slices): lowering on the host, hoisting, arena."""
    case, exp = load_golden("n30_sparse10000_sc27")
    plan = make_plan(case)
    w = plan.work_summary()
    assert plan.n_slices == 512 and w["hoisted_steps"] > 100 and w["exec_flops_per_slice"] <= w["ref_flops_per_slice"]
    assert plan.workspace_bytes < (1 << 30)
    assert plan.out_shape == (10000,)
