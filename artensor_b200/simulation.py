"""Drop-in for the contraction half of artensor/simulation.py.

`TensorNetworkSimulation.contraction` (simulation.py:90-117) and the copy of the slice loop in
`tensor_network_contraction` (simulation.py:198-213) are replaced by one call into the native
executor that walks the slice range on the GPU.  Circuit parsing, `_simplify` and the order
search are the reference's own code, imported lazily from the `artensor` package and used
unchanged (`prepare_contraction`); the scheme is compiled from the reference's contraction tree by
this package's own compilers (`scheme.py`, same tuple format; `scheme_compiler = "reference"`
selects the reference's).  A simulation can also be rebuilt from a frozen case file
(`from_case`), which needs no reference.

Differences from the reference, all deliberate:
  * leaf slicing fixes every sliced bond of a tensor at once (the packaged loop mis-indexes
    tensors with >= 2 sliced bonds, simulation.py:110-113);
  * slices can be restricted to a range and are partitioned over the ranks of a
    torch.distributed process group, followed by ONE sum-reduce of the partial amplitudes;
  * `device` must be a CUDA device; there is no CPU path.
"""
from copy import deepcopy

import numpy as np
import torch

from . import contraction as _c
from . import scheme as _scheme
from .backend import PlanOptions
from .plan import SchemeError


def _reference():
    try:
        import artensor  # noqa: F401
        return artensor
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError(
            "this entry point needs the reference package `artensor` for circuit building / order "
            "search (they are used unchanged); install it or load a frozen case with "
            "TensorNetworkSimulation.from_case()") from e


def check_bitstrings(bitstrings):
    """simulation.py:14-23"""
    if len(bitstrings):
        return 'sparse', len(np.unique(bitstrings))
    return 'normal', 1


def get_bond_tensors(tensor_bonds):
    """simulation.py:25-31"""
    bond_tensors = {}
    for i, bonds in tensor_bonds.items():
        for b in bonds:
            bond_tensors.setdefault(b, set()).add(i)
    return bond_tensors


def slicing_dims(tensors, tensor_bonds, slicing_bonds):
    """{bond: [(tid, dim)]} with dim counted on the actual un-sliced tensor (the hidden
    bitstring-batch dim of sparse final-qubit leaves included)."""
    out = {}
    for bond in slicing_bonds:
        lst = []
        for tid, bonds in tensor_bonds.items():
            if bond in bonds:
                hidden = tensors[tid].dim() - len(bonds)
                lst.append((tid, bonds.index(bond) + hidden))
        out[bond] = lst
    return out


def partition_slices(begin, end, rank, world):
    """Contiguous block partition of [begin, end) over `world` ranks."""
    n = end - begin
    lo = begin + (n * rank) // world
    hi = begin + (n * (rank + 1)) // world
    return lo, hi


class TensorNetworkSimulation:
    def __init__(self, tensors, tensor_bonds, bond_dims, final_qubits, bitstrings, pattern, max_bitstrings) -> None:
        self.tensors = tensors
        self.tensor_bonds = tensor_bonds
        self.bond_dims = bond_dims
        self.final_qubits = final_qubits
        self.bitstrings = bitstrings
        self.pattern = pattern
        self.max_bitstrings = max_bitstrings
        self.plan_options = PlanOptions()
        self.scheme_compiler = "b200"      # or "reference": contraction.py:23-59 / :208-342 unchanged
        self._plan_cache = {}

    # ---- planning: the reference's order search, unchanged (simulation.py:47-88) ----
    def prepare_contraction(self, sc_target=30, trials=6, iters=20, betas=np.linspace(0.1, 10, 100),
                            slicing_repeat=4, start_seed=0, alpha=32.0):
        ref = _reference()
        bond_tensors = get_bond_tensors(self.tensor_bonds)
        betas = np.linspace(3.0, 21.0, 61)   # simulation.py:52 overrides the argument
        order_slicing, slicing_bonds, self.ctree = ref.find_order(
            self.tensor_bonds, self.bond_dims, self.final_qubits, 0, self.max_bitstrings,
            sc_target=sc_target, trials=trials, iters=iters, betas=betas, start_seed=start_seed,
            slicing_repeat=slicing_repeat, alpha=alpha)
        self.slicing_bonds = list(slicing_bonds)
        self.slicing_indices = slicing_dims(self.tensors, self.tensor_bonds, self.slicing_bonds)
        self.update_scheme(sc_target, self.bitstrings)
        self.permute_dims = None
        if len(self.output_bonds) > 0:
            bond_inds = []
            for x in range(len(self.output_bonds)):
                assert len(bond_tensors[self.output_bonds[x]]) == 1
                tensor_id = next(iter(bond_tensors[self.output_bonds[x]]))
                assert tensor_id in self.final_qubits
                bond_inds.append(list(self.final_qubits).index(tensor_id))
            self.permute_dims = tuple(int(d) for d in np.argsort(bond_inds))
            if self.pattern == 'sparse':
                self.permute_dims = [0] + [dim + 1 for dim in self.permute_dims]

    def update_scheme(self, sc_target=30, bitstrings=[]):
        """simulation.py:79-88.  The tree is compiled by artensor_b200.scheme (layout-friendly mode
        orders, reproducible strings, chunking that covers every row: SURVEY.md 4.3-B2/B5) unless
        `scheme_compiler == "reference"`."""
        if self.scheme_compiler not in ("b200", "reference"):
            raise ValueError(f"scheme_compiler {self.scheme_compiler!r}: expected 'b200' or 'reference'")
        comp = _scheme if self.scheme_compiler == "b200" else _reference()
        if self.pattern == 'normal':
            self.scheme, self.output_bonds = comp.contraction_scheme(deepcopy(self.ctree))
            self.tensor_contraction_func = _c.tensor_contraction
        else:
            self.scheme, self.output_bonds, self.bitstrings_sorted = comp.contraction_scheme_sparse(
                deepcopy(self.ctree), bitstrings, sc_target=sc_target)
            self.tensor_contraction_func = _c.tensor_contraction_sparse
            assert len(self.bitstrings_sorted) <= self.max_bitstrings
        self._plan_cache.clear()

    # ---- the hot path ----
    def plan(self, mode="c64"):
        """Compiled plan for a compute mode: "c64" (fp32-accurate) or "chalf" (reduced-precision
        tensor-core products, see contraction._DTYPES)."""
        options = _c.mode_options(mode, self.plan_options)
        key = repr(options)
        if key not in self._plan_cache:
            self._plan_cache[key] = _c.ContractionPlan(
                self.scheme, {i: tuple(self.tensors[i].shape) for i in self._ids()},
                self.pattern == 'sparse', slicing_bonds=self.slicing_bonds,
                slicing_indices=self.slicing_indices, dtype="c64", options=options)
        return self._plan_cache[key]

    def _ids(self):
        return self.tensors.keys() if isinstance(self.tensors, dict) else range(len(self.tensors))

    def contraction(self, tensors=None, dtype=torch.complex64, device='cuda', slice_range=None, group=None,
                    reduce_result=True):
        """Sum of the contraction over slices (simulation.py:90-117).

        dtype:       torch.complex64 (fp32-accurate) or torch.complex32 (reduced-precision
                     complex-half tensor-core mode); the result is complex64 in both.
        slice_range: (begin, end) subset of slice ids, default all 2^S.
        group:       torch.distributed process group (or True for the default group): the slice
                     range is block-partitioned over its ranks and the partial amplitude tensors are
                     summed with one all-reduce (NCCL over NVLink when the tensors are CUDA).
        """
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError("artensor_b200 executes on CUDA devices only (no CPU fallback); got device=%r" % (device,))
        if dtype not in _c._DTYPES:
            raise RuntimeError(f"artensor_b200: unsupported dtype {dtype}; supported: {list(_c._DTYPES)}")
        src = self.tensors if tensors is None else tensors
        ids = list(self._ids())
        plan = self.plan(_c._DTYPES[dtype])
        begin, end = (0, plan.n_slices) if slice_range is None else slice_range
        if group is not None:
            import torch.distributed as dist
            pg = None if group is True else group
            begin, end = partition_slices(begin, end, dist.get_rank(pg), dist.get_world_size(pg))
        with torch.cuda.device(device):
            blob = plan.pack_leaves({i: src[i] for i in ids}, device=device)
            collect_tensor = torch.zeros(plan.out_shape, dtype=torch.complex64, device=device)
            ws = _c.get_workspace(device, plan.workspace_bytes)
            plan.execute(blob, collect_tensor, begin, end, ws, torch.cuda.current_stream(device).cuda_stream)
            if group is not None and reduce_result:
                import torch.distributed as dist
                dist.all_reduce(torch.view_as_real(collect_tensor), op=dist.ReduceOp.SUM,
                                group=None if group is True else group)
        if len(self.output_bonds) > 0 and self.permute_dims is not None:
            collect_tensor = collect_tensor.permute(self.permute_dims)
        return collect_tensor

    # ---- constructors ----
    @classmethod
    def from_circuit_file(cls, circuit_filename, bitstrings=[]):
        ref = _reference()
        return cls.from_tn_circuit(ref.TensorNetworkCircuit(circuit_filename), bitstrings)

    @classmethod
    def from_tn_circuit(cls, circ, bitstrings=[]):
        """simulation.py:135-148, with the reference's own network simplification."""
        ref = _reference()
        pattern, max_bitstrings = check_bitstrings(bitstrings)
        tensors, tensor_bonds, bond_dims, final_qubits = circ.to_numerical_tn()
        numerical_tn = ref.NumericalTensorNetwork(tensors, tensor_bonds, bond_dims, final_qubits)
        tensor_bonds_reorder, final_qubit_inds = numerical_tn._simplify(pattern)
        tensors = {i: numerical_tn.tensors[j] for i, j in enumerate(numerical_tn.tensors.keys())}
        return cls(tensors, tensor_bonds_reorder, bond_dims, final_qubit_inds, bitstrings, pattern, max_bitstrings)

    @classmethod
    def from_case(cls, case):
        """Rebuild a prepared simulation from a frozen case (artensor_b200.cases); no reference needed."""
        sim = cls(dict(case.leaves), case.leaf_bonds, None, None, case.extra.get("bitstrings_in", []),
                  case.pattern, len(case.bitstrings_sorted) if case.bitstrings_sorted else 1)
        sim.scheme = case.scheme
        sim.output_bonds = case.output_bonds
        sim.permute_dims = case.permute_dims
        sim.bitstrings_sorted = case.bitstrings_sorted
        sim.slicing_bonds = list(case.slicing_bonds)
        sim.slicing_indices = case.slicing_indices()
        sim.tensor_contraction_func = _c.tensor_contraction if case.pattern == 'normal' else _c.tensor_contraction_sparse
        return sim


def tensor_network_contraction(tensors, tensor_bonds, bond_dims, final_qubits, bitstrings=[], sc_target=31,
                               trial_num=8, alpha=0.0, dtype=torch.complex64, device='cuda'):
    """simulation.py:151-213 with the same signature and return value (collect_tensor, bitstrings)."""
    ref = _reference()
    pattern, max_bitstrings = check_bitstrings(bitstrings)
    numerical_tn = ref.NumericalTensorNetwork(tensors, tensor_bonds, bond_dims, final_qubits)
    tensor_bonds_reorder, final_qubit_inds = numerical_tn._simplify(pattern)
    leaves = {i: numerical_tn.tensors[j] for i, j in enumerate(numerical_tn.tensors.keys())}
    sim = TensorNetworkSimulation(leaves, tensor_bonds_reorder, numerical_tn.bond_dims, final_qubit_inds,
                                  bitstrings, pattern, max_bitstrings)
    # simulation.py:162-166: trials=trial_num, iters=50, start_seed=0
    sim.prepare_contraction(sc_target=sc_target, trials=trial_num, iters=50, start_seed=0, alpha=alpha)
    if pattern == 'sparse':
        assert len(sim.bitstrings_sorted) == max_bitstrings
    result = sim.contraction(dtype=dtype, device=device)
    return result, (sim.bitstrings_sorted if pattern == 'sparse' else bitstrings)


def quantum_circuit_simulation(circuit_filename, bitstrings=[], sc_target=31, trial_num=8, alpha=0.0,
                               dtype=torch.complex64, device='cuda'):
    """simulation.py:216-225"""
    ref = _reference()
    circ = ref.TensorNetworkCircuit(circuit_filename)
    tensors, tensor_bonds, bond_dims, final_qubits = circ.to_numerical_tn()
    return tensor_network_contraction(tensors, tensor_bonds, bond_dims, final_qubits, bitstrings, sc_target,
                                      trial_num, alpha, dtype, device)
