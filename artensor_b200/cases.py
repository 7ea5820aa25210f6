"""Case files: a contraction problem frozen to disk (leaves + scheme + slicing info).

The reference keeps the contraction tree, scheme and slicing indices in memory only
(`artensor/simulation.py:53-67`), so the minutes-long order search has to be repeated in
every process.  A *case* freezes everything the numerical executor consumes:

  leaves            {tensor id: complex64 ndarray}   (what `simulation.py:92-99` would upload)
  leaf_bonds        {tensor id: [bond, ...]}         un-sliced bond list of every leaf
  scheme            list of step tuples in the reference's own format
                    (`artensor/contraction.py:53` normal, `:327-335` sparse)
  slicing_bonds     [bond, ...]  in the order the slice id enumerates them (MSB first,
                    `simulation.py:108`: np.binary_repr(s, S))
  output_bonds, permute_dims, bitstrings_sorted      as set by `prepare_contraction`
  pattern           'normal' | 'sparse'

Index tensors inside sparse steps are stored as numpy int64 and re-materialised as CPU
int64 torch tensors on load, which is what `contraction_scheme_sparse` emits
(`contraction.py:249-283`).  Nothing in here depends on the reference package.
"""
import gzip
import pickle

import numpy as np
import torch

FORMAT_VERSION = 1


def _step_to_plain(step):
    edge, eq = step[0], step[1]
    out = {"edge": (int(edge[0]), int(edge[1])), "eq": str(eq), "len": len(step)}
    if len(step) >= 3:
        out["batch"] = [[np.asarray(t, dtype=np.int64) for t in side] for side in step[2]]
    if len(step) >= 5:
        out["rshape"] = None if step[3] is None else tuple(int(x) for x in step[3])
        out["next_shape"] = tuple(int(x) for x in step[4])
    return out


def _step_from_plain(p):
    edge, eq = tuple(p["edge"]), p["eq"]
    if p["len"] == 2:
        return (edge, eq)
    batch = [[torch.from_numpy(np.ascontiguousarray(a)) for a in side] for side in p["batch"]]
    if p["len"] == 3:
        return (edge, eq, batch)
    return (edge, eq, batch, p["rshape"], p["next_shape"])


def scheme_to_plain(scheme):
    return [_step_to_plain(s) for s in scheme]


def scheme_from_plain(plain):
    return [_step_from_plain(p) for p in plain]


def save_case(path, *, name, pattern, leaves, leaf_bonds, scheme, slicing_bonds,
              output_bonds, permute_dims, bitstrings_sorted, n_qubits, extra=None):
    case = {
        "format": FORMAT_VERSION,
        "name": name,
        "pattern": pattern,
        "n_qubits": int(n_qubits),
        "leaves": {int(k): np.ascontiguousarray(v.detach().cpu().numpy() if torch.is_tensor(v) else v)
                   for k, v in leaves.items()},
        "leaf_bonds": {int(k): list(v) for k, v in leaf_bonds.items()},
        "scheme": scheme_to_plain(scheme),
        "slicing_bonds": list(slicing_bonds),
        "output_bonds": list(output_bonds),
        "permute_dims": None if permute_dims is None else [int(x) for x in permute_dims],
        "bitstrings_sorted": None if bitstrings_sorted is None else [str(b) for b in bitstrings_sorted],
        "extra": extra or {},
    }
    with gzip.open(path, "wb", compresslevel=6) as f:
        pickle.dump(case, f, protocol=4)
    return case


class Case:
    """A loaded case.  `leaves` are CPU complex64 torch tensors keyed by tensor id."""

    def __init__(self, raw):
        if raw.get("format") != FORMAT_VERSION:
            raise ValueError(f"unsupported case format {raw.get('format')}")
        self.name = raw["name"]
        self.pattern = raw["pattern"]
        self.n_qubits = raw["n_qubits"]
        self.leaves = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in raw["leaves"].items()}
        self.leaf_bonds = raw["leaf_bonds"]
        self.scheme = scheme_from_plain(raw["scheme"])
        self.slicing_bonds = raw["slicing_bonds"]
        self.output_bonds = raw["output_bonds"]
        self.permute_dims = raw["permute_dims"]
        self.bitstrings_sorted = raw["bitstrings_sorted"]
        self.extra = raw["extra"]

    @property
    def n_slices(self):
        return 1 << len(self.slicing_bonds)

    def slicing_indices(self):
        """{bond: [(tid, dim)]} with `dim` taken on the UN-sliced tensor, the same
        convention as `simulation.py:60-65` (dims of the actual tensor, so the hidden
        bitstring-batch dim of sparse final-qubit leaves is accounted for)."""
        out = {}
        for bond in self.slicing_bonds:
            lst = []
            for tid, bonds in self.leaf_bonds.items():
                if bond in bonds:
                    hidden = self.leaves[tid].dim() - len(bonds)
                    lst.append((tid, bonds.index(bond) + hidden))
            out[bond] = lst
        return out


def load_case(path):
    with gzip.open(path, "rb") as f:
        raw = pickle.load(f)
    return Case(raw)


def slice_leaves(leaves, slicing_bonds, slicing_indices, slice_id):
    """Leaf tensors of slice `slice_id`: every sliced bond fixed to its bit.

    Semantics of the slice loop in `simulation.py:107-113`, except that all sliced dims of
    one tensor are indexed simultaneously on the un-sliced tensor.  The packaged loop applies
    `select` sequentially with un-sliced dims and is off by one for tensors carrying two or
    more sliced bonds; `examples/sycamore.ipynb` cell 11 does it this way.
    """
    S = len(slicing_bonds)
    per_tensor = {}
    for x, bond in enumerate(slicing_bonds):
        bit = (slice_id >> (S - 1 - x)) & 1  # np.binary_repr(s, S)[x]
        for tid, dim in slicing_indices[bond]:
            per_tensor.setdefault(tid, {})[dim] = bit
    out = dict(leaves) if isinstance(leaves, dict) else list(leaves)
    for tid, dims in per_tensor.items():
        t = leaves[tid]
        idx = tuple(dims.get(d, slice(None)) for d in range(t.dim()))
        out[tid] = t[idx].clone()
    return out
