"""Scheme compilers (artensor_b200/scheme.py) against the reference's own.

The contraction tree comes from the reference's order search (unchanged, out of scope); both the
reference's scheme and ours are then run on the REFERENCE executors (artensor/contraction.py:62-76,
:132-205) in complex128 on the same leaves and slices, and the amplitudes must agree bitstring by
bitstring.  Needs the reference package (this container: /root/reference); the host-logic tests at
the bottom do not.
"""
import os
import sys
from copy import deepcopy

import numpy as np
import pytest
import torch

from artensor_b200 import scheme as S
from artensor_b200.cases import slice_leaves
from artensor_b200.simulation import slicing_dims

REF = os.environ.get("ARTENSOR_REFERENCE", "/root/reference")
N12 = os.path.join(REF, "tests", "circuit_n12_m14_s0_e0_pEFGH.qsim")


def reference():
    if not os.path.exists(N12):
        pytest.skip("the reference package is not present on this machine")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import artensor
    return artensor


def random_bitstrings(n, count, seed):
    rng = np.random.RandomState(seed)
    seen = set()
    while len(seen) < count:
        seen.add("".join(map(str, rng.randint(0, 2, size=n))))
    return sorted(seen)


def correlated_bitstrings(n, k, seed):
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


_prepared = {}


def prepared(bitstrings_key, bitstrings, sc_target):
    """Reference simulation with order search done (cached per configuration)."""
    ref = reference()
    key = (bitstrings_key, sc_target)
    if key not in _prepared:
        sim = ref.TensorNetworkSimulation.from_circuit_file(N12, bitstrings)
        sim.prepare_contraction(sc_target=sc_target, trials=2, iters=5, slicing_repeat=1, start_seed=0)
        _prepared[key] = sim
    return _prepared[key]


def run_on_reference_executor(ref, sim, scheme, sparse):
    """Sum over all slices with shift-corrected leaf slicing, complex128."""
    from artensor.contraction import tensor_contraction, tensor_contraction_sparse
    leaves = {i: t.to(torch.complex128) for i, t in sim.tensors.items()}
    bonds = list(sim.slicing_indices.keys())
    dims = slicing_dims(sim.tensors, sim.tensor_bonds, bonds)
    total = None
    for s in range(1 << len(bonds)):
        sl = slice_leaves(leaves, bonds, dims, s)
        out = (tensor_contraction_sparse if sparse else tensor_contraction)(sl, scheme)
        total = out.clone() if total is None else total + out
    return total


def chunks_are_valid(scheme):
    for step in scheme:
        if len(step) == 5 and step[3] is None:
            li, lj = [len(c) for c in step[2][0]], [len(c) for c in step[2][1]]
            if li != lj or 0 in li or sum(li) != step[4][0]:
                return False
    return True


def test_normal_scheme_matches_reference_scheme():
    ref = reference()
    sim = prepared("full", [], 30)
    ours, out_bonds = S.contraction_scheme(deepcopy(sim.ctree))
    theirs, their_bonds = ref.contraction_scheme(deepcopy(sim.ctree))
    assert len(ours) == len(theirs)
    assert sorted(map(str, out_bonds)) == sorted(map(str, their_bonds))
    a = run_on_reference_executor(ref, sim, ours, False)
    b = run_on_reference_executor(ref, sim, theirs, False)
    # bring ours into the reference's output mode order
    b_perm = [out_bonds.index(x) for x in their_bonds]
    a = a.permute(b_perm)
    assert a.shape == b.shape
    assert (a - b).abs().max().item() < 1e-12 * b.abs().max().item() + 1e-15


def test_normal_scheme_is_layout_friendly_and_deterministic():
    """Output modes = kept modes of the left operand in its order, then the new modes of the right
    operand in its order; letters by first appearance."""
    sim = prepared("full", [], 30)
    ours, _ = S.contraction_scheme(deepcopy(sim.ctree))
    again, _ = S.contraction_scheme(deepcopy(sim.ctree))
    assert ours == again
    for (i, j), eq in ours:
        lhs, out = eq.split("->")
        la, lb = lhs.split(",")
        kept_a = [c for c in la if c in out]
        new_b = [c for c in lb if c in out and c not in la]
        assert list(out) == kept_a + new_b
        seen = []
        for c in la + lb + out:
            if c not in seen:
                seen.append(c)
        assert seen == S.ALPHABET[:len(seen)]


@pytest.mark.parametrize("name,bits,sc", [
    ("kat5", ["100001000001", "000101111011", "011000101100", "111001100001", "001110110000"], 30),
    ("rand64", random_bitstrings(12, 64, 1), 9),
    ("rand100", random_bitstrings(12, 100, 2), 8),
    ("corr256", correlated_bitstrings(12, 8, 3), 10),
])
def test_sparse_scheme_matches_reference_scheme(name, bits, sc):
    ref = reference()
    sim = prepared(name, bits, sc)
    ours, rest, ordered = S.contraction_scheme_sparse(deepcopy(sim.ctree), bits, sc_target=sc)
    theirs, their_rest, their_ordered = ref.contraction_scheme_sparse(deepcopy(sim.ctree), bits, sc_target=sc)
    assert rest == their_rest == []
    assert sorted(ordered) == sorted(set(bits)) == sorted(their_ordered)
    assert [s[0] for s in ours] == [s[0] for s in theirs]                 # same edges, same slots
    assert [len(s) for s in ours] == [len(s) for s in theirs]             # same step kinds ...
    assert [s[3] is None for s in ours if len(s) == 5] == [s[3] is None for s in theirs if len(s) == 5]
    assert chunks_are_valid(ours)
    a = run_on_reference_executor(ref, sim, ours, True)
    mine = dict(zip(ordered, a.reshape(-1).tolist()))
    if chunks_are_valid(theirs):
        b = run_on_reference_executor(ref, sim, theirs, True)
        ref_amp = dict(zip(their_ordered, b.reshape(-1).tolist()))
        scale = max(abs(v) for v in ref_amp.values())
        assert max(abs(mine[k] - ref_amp[k]) for k in ref_amp) < 1e-12 * scale
    # and through this package's own parser / numpy oracle of the executor
    from artensor_b200.plan import SchemeParser
    steps = SchemeParser({i: tuple(t.shape) for i, t in slice_leaves(
        sim.tensors, list(sim.slicing_indices), slicing_dims(sim.tensors, sim.tensor_bonds, list(sim.slicing_indices)), 0
    ).items()}, True).parse(ours)
    assert len(steps) == len(ours)


def test_sparse_scheme_known_answers():
    """tests/test_circuits.py:25-31 of the reference (effective tolerance 3e-5)."""
    ref = reference()
    kat = {"100001000001": 0.0198028199 + 0.0106442748j, "000101111011": 0.00497586094 - 0.0245072283j,
           "011000101100": -0.00853562169 - 0.00701293815j, "111001100001": -0.0100137182 + 0.0147468708j,
           "001110110000": 0.00681955926 + 0.0106616206j}
    sim = prepared("kat5", list(kat), 30)
    ours, _, ordered = S.contraction_scheme_sparse(deepcopy(sim.ctree), list(kat), sc_target=30)
    a = run_on_reference_executor(ref, sim, ours, True).reshape(-1).tolist()
    for b, v in zip(ordered, a):
        assert abs(v - kat[b]) < 3e-5 * abs(kat[b])


def test_drop_in_simulation_uses_own_compiler_and_lowers():
    """artensor_b200.TensorNetworkSimulation.prepare_contraction: the reference's order search, this
    package's scheme compiler, this package's plan lowering (emulated on the host: the records the
    CUDA library would execute), against the reference executor on the reference's scheme."""
    ref = reference()
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emulate
    from artensor_b200 import TensorNetworkSimulation
    from artensor_b200.backend import ContractionPlan
    bits = random_bitstrings(12, 64, 1)
    sim = TensorNetworkSimulation.from_circuit_file(N12, bits)
    assert sim.scheme_compiler == "b200"
    sim.prepare_contraction(sc_target=9, trials=2, iters=5, slicing_repeat=1, start_seed=0)
    assert chunks_are_valid(sim.scheme)
    plan = ContractionPlan(sim.scheme, {i: tuple(sim.tensors[i].shape) for i in sim._ids()}, True,
                           slicing_bonds=sim.slicing_bonds, slicing_indices=sim.slicing_indices, build_native=False)
    got = emulate.run_plan(plan, plan.pack_leaves(sim.tensors).numpy(), range(plan.n_slices)).reshape(-1)
    mine = dict(zip(sim.bitstrings_sorted, got.tolist()))
    # the reference end to end (its own compiler), same order search settings -> same tree
    rsim = prepared("rand64", bits, 9)
    want = run_on_reference_executor(ref, rsim, rsim.scheme, True).reshape(-1).tolist()
    ref_amp = dict(zip(rsim.bitstrings_sorted, want))
    scale = max(abs(v) for v in ref_amp.values())
    assert max(abs(mine[k] - ref_amp[k]) for k in ref_amp) < 5e-6 * scale
    sim.scheme_compiler = "reference"
    sim.update_scheme(9, bits)
    assert [s[1] for s in sim.scheme] != []                      # the reference's compiler is still selectable
    with pytest.raises(ValueError):
        sim.scheme_compiler = "other"
        sim.update_scheme(9, bits)


@pytest.mark.parametrize("reuse", [False, True])
def test_prepare_contraction_sweep_keeps_the_cheapest_tree(reuse):
    """SURVEY.md 8-f4 as an API: the reference's order search over a grid of settings, every tree priced with this
    package's step-time model (for all 2^S slices; amortised under cross-slice reuse when asked), the cheapest kept.
    The simulation must be left prepared with exactly that tree -- and contract to the reference's amplitudes."""
    ref = reference()
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emulate
    from artensor_b200 import TensorNetworkSimulation, PlanOptions
    from artensor_b200.backend import ContractionPlan
    bits = random_bitstrings(12, 64, 1)
    sim = TensorNetworkSimulation.from_circuit_file(N12, bits)
    sim.plan_options = PlanOptions(slice_reuse=reuse)
    res = sim.prepare_contraction_sweep(sc_targets=(8, 9), alphas=(32.0,), start_seeds=(0, 3), trials=2, iters=3)
    assert len(res) == 4 and all(r["fits"] for r in res)
    assert [r["task_seconds"] for r in res] == sorted(r["task_seconds"] for r in res)
    best = res[0]
    assert len(sim.slicing_bonds) == best["sliced_bonds"] and sim._sc_target == best["sc_target"]
    plan = ContractionPlan(sim.scheme, {i: tuple(sim.tensors[i].shape) for i in sim._ids()}, True,
                           slicing_bonds=sim.slicing_bonds, slicing_indices=sim.slicing_indices,
                           options=sim.plan_options, build_native=False)
    model = plan.reuse_summary()
    per_slice = model["amortised_s"] if reuse else model["full_s"]
    assert abs(per_slice - best["seconds_per_slice"]) <= 1e-9 * per_slice       # the kept tree (and bit order) is the priced one
    got = emulate.run_plan(plan, plan.pack_leaves(sim.tensors).numpy(), range(plan.n_slices), reuse=reuse, poison=reuse).reshape(-1)
    mine = dict(zip(sim.bitstrings_sorted, got.tolist()))
    rsim = prepared("rand64", bits, 9)
    want = run_on_reference_executor(ref, rsim, rsim.scheme, True).reshape(-1).tolist()
    ref_amp = dict(zip(rsim.bitstrings_sorted, want))
    scale = max(abs(v) for v in ref_amp.values())
    assert max(abs(mine[k] - ref_amp[k]) for k in ref_amp) < 5e-6 * scale
    with pytest.raises(RuntimeError):
        sim.prepare_contraction_sweep(sc_targets=(9,), alphas=(32.0,), start_seeds=(0,), trials=1, iters=1, max_workspace_bytes=16)


# ------------------------------------------------------------------------------------------------
# host logic that needs no reference
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,rank,sc", [(100, 7, 10), (1, 40, 30), (1024, 20, 30), (7, 3, 30), (1000, 18, 20),
                                        (33, 9, 12), (256, 28, 30), (3, 30, 30)])
def test_chunk_bounds_cover_every_row_once(n, rank, sc):
    """The reference's rule (contraction.py:288-297) loses rows when N mod chunks > N // chunks
    (100 rows in 32 chunks -> 33 x 3 = 99) and emits empty chunks for one row of a large operand."""
    b = S.chunk_bounds(n, rank, sc)
    assert b[0][0] == 0 and b[-1][1] == n
    assert all(e > s for s, e in b)
    assert all(b[k][1] == b[k + 1][0] for k in range(len(b) - 1))
    budget = 2 ** (sc - 2 - rank) if sc - 2 - rank >= 0 else 1
    assert max(e - s for s, e in b) <= budget


def test_spread_restrict_roundtrip():
    rng = np.random.RandomState(0)
    width = 9
    locs = sorted(rng.choice(width, 4, replace=False).tolist())
    codes = np.arange(16)
    wide = S._spread(codes, locs, width)
    assert np.array_equal(S._restrict(wide, locs, width), codes)
    for c, w in zip(codes, wide):
        s = np.binary_repr(int(w), width)
        assert "".join(s[q] for q in locs) == np.binary_repr(int(c), 4)


def test_einsum_equation_first_appearance():
    assert S.einsum_equation([-1, "x", "y"], ["y", "z"], [-1, "x", "z"]) == "ABC,CD->ABD"
    with pytest.raises(S.SchemeError):
        S.einsum_equation(list(range(60)), [], [])


# ------------------------------------------------------------------------------------------------
# open-qubit sharding of full-amplitude contractions
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_regular,n_bits", [(0, 1), (2, 2), (3, 3)])
def test_open_qubit_shards_concatenate_to_the_full_state(n_regular, n_bits):
    """`prepare_open_qubit_shards(b)` fixes the first b output qubits per shard: every shard's sum
    over its (regular) slices, run by the REFERENCE executor in complex128, must equal the
    matching block of the un-sharded full state, with and without regular slicing.  (The n12 tree
    needs no slicing at sc_target >= 12 and below that the reference slices OPEN bonds, SURVEY.md
    4.3; so the regular sliced bonds are taken by hand: inner bonds of the tree.)"""
    ref = reference()
    from artensor.contraction import tensor_contraction
    from artensor_b200 import TensorNetworkSimulation
    sim = TensorNetworkSimulation.from_circuit_file(N12, [])
    assert isinstance(sim, ref.TensorNetworkSimulation)           # the subclass of the reference's class
    sim.prepare_contraction(sc_target=30, trials=2, iters=5, slicing_repeat=1, start_seed=0)
    assert sim.slicing_bonds == [] and len(sim.output_bonds) == 12
    if n_regular:
        tree = deepcopy(sim.ctree)
        inner = sorted(b for b, ts in tree.tn.bond_tensors.items() if len(ts) == 2)[5:5 + n_regular]
        for b in inner:
            tree.slicing(b)
        sim.ctree, sim.slicing_bonds = tree, inner
        sim.slicing_indices = slicing_dims(sim.tensors, sim.tensor_bonds, inner)
        qubit_of = {b: q for q, b in zip(np.argsort(np.argsort(sim.permute_dims)), sim.output_bonds)}
        sim.update_scheme()
        sim.permute_dims = [int(d) for d in np.argsort([qubit_of[b] for b in sim.output_bonds])]
    leaves = {i: t.to(torch.complex128) for i, t in sim.tensors.items()}

    def contract(slice_ids):
        total = None
        for s in slice_ids:
            out = tensor_contraction(slice_leaves(leaves, sim.slicing_bonds, sim.slicing_indices, s), sim.scheme)
            total = out.clone() if total is None else total + out
        return total.permute(sim.permute_dims)
    full = contract(range(1 << n_regular))                          # [2] * 12 in qubit order
    sim.prepare_open_qubit_shards(n_bits)
    assert len(sim.shard_bonds) == n_bits and sim.slicing_bonds[:n_bits] == sim.shard_bonds
    assert len(sim.slicing_bonds) == n_regular + n_bits and len(sim.output_bonds) == 12 - n_bits
    per_shard = 1 << n_regular
    blocks = [contract(range(v * per_shard, (v + 1) * per_shard)) for v in range(1 << n_bits)]
    got = torch.stack(blocks).reshape([2] * 12)
    assert (got - full).abs().max().item() < 1e-12 * full.abs().max().item()
    with pytest.raises(ValueError):
        sim.prepare_open_qubit_shards(1)                            # already sharded
    # the reuse bit order leaves the shard bonds where they are (most significant: a shard stays a contiguous slice
    # range) and permutes the regular sliced bonds only -- every shard's sum over its slices is unchanged
    before = list(sim.slicing_bonds)
    sim.optimize_slice_order()
    assert sim.slicing_bonds[:n_bits] == before[:n_bits] and sorted(sim.slicing_bonds[n_bits:]) == sorted(before[n_bits:])
    for v in (0, (1 << n_bits) - 1):
        again = contract(range(v * per_shard, (v + 1) * per_shard))
        assert (again - blocks[v]).abs().max().item() < 1e-12 * full.abs().max().item()
