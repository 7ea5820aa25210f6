"""Pins the numpy oracle (oracle/tn_oracle.py) against outputs of the REAL reference
(fixtures written by tools/gen_cases.py, which imports /root/reference), against the
reference's own known-answer table and against Google's amplitude file."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import tn_oracle as O

SMALL = ["n12_full", "n12_sparse5", "n12_sparse64_sc9", "n12_sparse100_sc8", "n12_sparse256c_sc10",
         "n12_full_own", "n12_sparse100_sc8_own"]   # _own: scheme compiled by artensor_b200/scheme.py

# tests/test_circuits.py:25-31 of the reference
KAT_N12 = {
    "100001000001": 0.0198028199 + 1j * 0.0106442748,
    "000101111011": 0.00497586094 + 1j * -0.0245072283,
    "011000101100": -0.00853562169 + 1j * -0.00701293815,
    "111001100001": -0.0100137182 + 1j * 0.0147468708,
    "001110110000": 0.00681955926 + 1j * 0.0106616206,
}


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_per_slice(name):
    case, exp = load_golden(name)
    sidx = case.slicing_indices()
    ids = exp["slice_ids"]
    pick = ids if len(ids) <= 16 else ids[:: max(1, len(ids) // 16)]
    for s in pick:
        k = int(np.where(ids == s)[0][0])
        r = O.contract_slices(case.leaves, case.scheme, case.pattern, case.slicing_bonds, sidx, [s]).reshape(-1)
        # two complex64 evaluations with different summation orders: agree to a few ulp of the scale
        assert _rel(r, exp["per_slice_c64"][k]) < 5e-6
        assert _rel(r, exp["per_slice_c128"][k]) < 5e-6


def test_oracle_known_answers_n12_sparse():
    """The reference's own KAT (effective tolerance ~3e-5: the table comes from another
    simulator and the gate tensors are unitary only to 1e-7)."""
    case, _ = load_golden("n12_sparse5")
    r = O.contract_slices(case.leaves, case.scheme, case.pattern, case.slicing_bonds, case.slicing_indices(), [0])
    for b, amp in zip(case.bitstrings_sorted, r.reshape(-1)):
        assert abs(amp - KAT_N12[b]) / abs(KAT_N12[b]) < 5e-5


def test_oracle_known_answers_n12_full():
    case, _ = load_golden("n12_full")
    r = O.contract_slices(case.leaves, case.scheme, case.pattern, case.slicing_bonds, case.slicing_indices(), [0])
    amps = np.transpose(r, case.permute_dims).reshape(-1)
    for b, amp in KAT_N12.items():
        assert abs(amps[int(b, 2)] - amp) / abs(amp) < 5e-5


def test_oracle_sliced_sum_is_slice_invariant_total():
    """Sum over all slices of a sliced scheme == amplitudes of the same bitstrings taken from
    the un-sliced full-amplitude contraction (different tree, different slicing)."""
    case, exp = load_golden("n12_sparse64_sc9")
    total = exp["per_slice_c128"].sum(axis=0)
    full, fexp = load_golden("n12_full")
    amps = np.transpose(fexp["per_slice_c128"][0].reshape(fexp["shape"]), full.permute_dims).reshape(-1)
    want = np.array([amps[int(b, 2)] for b in case.bitstrings_sorted])
    assert _rel(total, want) < 1e-5


def test_oracle_scientific_notation():
    case, exp = load_golden("n12_sparse5")
    leaves = {k: v.numpy() for k, v in case.leaves.items()}
    factor, t = O.tensor_contraction_sparse(dict(leaves), case.scheme, scientific_notation=True)
    assert _rel((t * 10.0 ** factor).reshape(-1), exp["per_slice_c64"][0]) < 5e-6


@pytest.mark.slow
def test_oracle_n30_sliced_matches_reference_and_google():
    """n30 m14, 64 bitstrings, sliced with chunked batched steps: slice 0 vs the reference,
    and the reference's 4-slice total vs Google's amplitudes (~1e-4)."""
    case, exp = load_golden("n30_sparse64_sc26")
    r = O.contract_slices(case.leaves, case.scheme, case.pattern, case.slicing_bonds, case.slicing_indices(), [0])
    assert _rel(r.reshape(-1), exp["per_slice_c64"][0]) < 5e-6
    google = dict(zip(case.extra["bitstrings_in"], case.extra["google_amplitudes"]))
    total = exp["per_slice_c64"].sum(axis=0)
    want = np.array([google[b] for b in case.bitstrings_sorted])
    rel = np.abs(total - want) / np.abs(want)
    assert np.median(rel) < 2e-4 and rel.max() < 2e-3


def test_oracles_on_the_sliced_10000_amplitude_scheme():
    """n30 m14, Google's 10000 bitstrings at sc_target 27 (512 slices): slice 3 through the numpy and
    the torch oracle against the reference executor's output (row tables of ~10^4 entries, subset
    outer steps, a batched step with a 2^11-long contraction)."""
    from oracle import tn_oracle_torch as OT
    case, exp = load_golden("n30_sparse10000_sc27")
    k = int(np.where(exp["slice_ids"] == 3)[0][0])
    r = O.contract_slices(case.leaves, case.scheme, case.pattern, case.slicing_bonds, case.slicing_indices(), [3])
    assert _rel(r.reshape(-1), exp["per_slice_c64"][k]) < 5e-6
    rt = OT.contract_slices(case, [3]).reshape(-1).numpy()
    assert _rel(rt, exp["per_slice_c64"][k]) < 5e-6


@pytest.mark.parametrize("name", SMALL)
def test_torch_oracle_matches_reference_goldens(name):
    """The torch-on-CPU restatement (the CPU baseline that bench.py times) is pinned the same way."""
    from oracle import tn_oracle_torch as OT
    case, exp = load_golden(name)
    ids = exp["slice_ids"]
    for s in list(ids[:2]) + [ids[-1]]:
        k = int(np.where(ids == s)[0][0])
        r = OT.contract_slices(case, [int(s)]).reshape(-1).numpy()
        assert _rel(r, exp["per_slice_c64"][k]) < 5e-6


def test_torch_oracle_step_shrinking_is_linear():
    from oracle import tn_oracle_torch as OT
    eq, sa, sb, scale = OT._shrunk_step("abcde,efc->abdf", [2] * 5, [2, 2, 2], 8)
    assert scale == 4 and eq == "abcde,efc->abdf" and sa == [1, 1, 2, 2, 2] and sb == [2, 2, 2]
    # a shared row label is halved first, and only while the largest tensor is too large
    eq, sa, sb, scale = OT._shrunk_step("zab,zbc->zac", [64, 2, 2], [64, 2, 2], 64)
    assert scale == 4 and sa == [16, 2, 2] and sb == [16, 2, 2]
    secs, n, scaled, _ = OT.estimate_slice_seconds([("abcde,efc->abdf", [2] * 5, [2, 2, 2])], max_elems=8)
    assert n == 1 and scaled == 1 and secs > 0


def test_open_qubit_shard_is_a_block_of_the_unsharded_slice():
    """n30 m14 full amplitude sharded over its first 3 output qubits (fixture n30_full_shard3,
    SURVEY.md 8e): slice id 0 = (shard 0, regular slice 0), contracted by the torch oracle, must be
    the block of the UNSHARDED slice 0 whose shard qubits are 0 -- checked on the entries of the
    reference's recorded output of that slice (fixture n30_full) that fall into the block."""
    from oracle import tn_oracle_torch
    full, fexp = load_golden("n30_full")
    case, exp = load_golden("n30_full_shard3")
    n_sh = int(case.extra["n_shard_bonds"])
    assert n_sh == 3 and case.slicing_bonds[n_sh:] == full.slicing_bonds and int(fexp["slice_ids"][0]) == 0
    got = tn_oracle_torch.contract_slices(case, [0]).reshape(-1).numpy()
    want_own = exp["per_slice_c64"][list(exp["slice_ids"]).index(0)]
    assert _rel(got[exp["sample_idx"]], want_own) < 1e-5              # the reference executor on the sharded scheme
    # entries of the unsharded slice: flat index -> bit of every output bond
    idx = fexp["sample_idx"]
    n_out = len(full.output_bonds)
    bit = {b: (idx >> (n_out - 1 - d)) & 1 for d, b in enumerate(full.output_bonds)}
    in_block = np.ones(len(idx), dtype=bool)
    for b in case.slicing_bonds[:n_sh]:
        in_block &= bit[b] == 0                                       # shard 0: every shard qubit is 0
    pos = np.zeros(len(idx), dtype=np.int64)
    for d, b in enumerate(case.output_bonds):
        pos |= bit[b] << (len(case.output_bonds) - 1 - d)
    assert in_block.sum() > 500
    assert _rel(got[pos[in_block]], fexp["per_slice_c64"][0][in_block]) < 1e-5


@pytest.mark.parametrize("name", ["n12_sparse5_f64leaves", "n12_sparse64_sc9_f64leaves"])
def test_f64_built_leaves_reproduce_the_state_vector(name):
    """Leaves built in float64 and cast once: the oracle in complex128 on those
    leaves must reproduce the circuit's float64 state vector (recorded in the case) to the
    precision of the complex64 leaves themselves, ~1e-7 of the rms amplitude."""
    case, exp = load_golden(name)
    sv = case.extra["statevector_f64"]
    want = np.array([sv[b] for b in case.bitstrings_sorted])
    got = O.contract_slices(case.leaves, case.scheme, case.pattern, case.slicing_bonds, case.slicing_indices(),
                            range(case.n_slices), dtype=np.complex128).reshape(-1)
    rms = np.sqrt(np.mean(np.abs(want) ** 2))
    assert np.abs(got - want).max() < 5e-7 * rms
    assert np.abs(exp["per_slice_c128"].sum(axis=0) - got).max() < 1e-12 * rms       # and the reference executor agrees
