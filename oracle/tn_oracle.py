"""CPU ORACLE -- test infrastructure, NOT product code.

A numpy restatement of artensor's numerical contraction executor.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module; the product path (`artensor_b200/`) never does and has no CPU fallback.

What is restated (reference file:line, relative to /root/reference):
  * `tensor_contraction`          artensor/contraction.py:62-76
  * `tensor_contraction_sparse`   artensor/contraction.py:132-205
  * the slice loop                artensor/simulation.py:103-117 (and its copy :198-213)
The arithmetic of the reference is `torch.einsum` (third-party, not vendored; the reference
pins pytorch==1.12.1 in examples/requirements.txt:7, this image has 2.11.0).  For two operands
torch.einsum is "permute to [batch, M, K] x [batch, K, N], reshape, bmm, reshape back"; that is
what `einsum_pair` below does with numpy (`np.matmul` on complex arrays).

PINNING: the oracle is checked in `tests/test_oracle.py` against golden outputs produced by
running the real reference in the build container (`tools/gen_cases.py`, fixtures under
`tests/golden/`), against the reference's own known-answer table
(`tests/test_circuits.py:25-31`) and against Google's amplitude file for the n30 circuit
(`examples/amplitudes_n30_m14_s0_e0_pEFGH_10000.txt`, sampled into the fixtures).

Deliberate divergence: leaf slicing fixes all sliced dims of a tensor at once on the un-sliced
tensor.  The packaged loop (`simulation.py:110-113`) applies `select` sequentially with
un-sliced dim numbers and is off by one for tensors with >= 2 sliced bonds;
`examples/sycamore.ipynb` cell 11 slices correctly and is what this follows.
"""
import numpy as np


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def einsum_pair(eq, a, b):
    """Two-operand einsum the way torch.einsum lowers it (contraction.py:70 and :147-190 call
    sites): classify labels, permute both operands to [batch, M, K] / [batch, K, N], matmul,
    reshape, permute to the requested output order."""
    lhs, lo = eq.split("->")
    la, lb = lhs.split(",")
    assert len(la) == a.ndim and len(lb) == b.ndim, (eq, a.shape, b.shape)
    ext = {}
    for lab, e in list(zip(la, a.shape)) + list(zip(lb, b.shape)):
        assert ext.setdefault(lab, e) == e, f"extent mismatch for {lab} in {eq}"
    batch = [l for l in la if l in lb and l in lo]
    con = [l for l in la if l in lb and l not in lo]
    m = [l for l in la if l not in lb]
    n = [l for l in lb if l not in la]
    assert all(l in lo for l in m + n), f"label summed inside one operand: {eq}"
    pa = [la.index(l) for l in batch + m + con]
    pb = [lb.index(l) for l in batch + con + n]
    size = lambda labs: int(np.prod([ext[l] for l in labs], dtype=np.int64)) if labs else 1
    B, M, N, K = size(batch), size(m), size(n), size(con)
    a2 = np.ascontiguousarray(np.transpose(a, pa)).reshape(B, M, K)
    b2 = np.ascontiguousarray(np.transpose(b, pb)).reshape(B, K, N)
    c2 = np.matmul(a2, b2)
    cur = batch + m + n
    c = c2.reshape([ext[l] for l in cur])
    return np.transpose(c, [cur.index(l) for l in lo])


def tensor_contraction(tensors, scheme):
    """contraction.py:62-76: tensors[i] = einsum(eq, tensors[i], tensors[j]) for every step;
    returns the last tensors[i].  Mutates `tensors` like the reference does."""
    i = None
    for s in scheme:
        i, j = s[0]
        tensors[i] = einsum_pair(s[1], _np(tensors[i]), _np(tensors[j]))
    return tensors[i]


def tensor_contraction_sparse(tensors, contraction_scheme, scientific_notation=False):
    """contraction.py:132-205.  Same branch structure as the reference:
    A (:140-175) chunked batched step, results concatenated along dim 0 in chunk order;
    B (:176-179) single-chunk batched step; C (:180-188) outer step + reshape + optional row
    subset; D (:189-191) plain step.  Optional max-abs rescaling (:197-200)."""
    factor = 0.0
    i = None
    for step in contraction_scheme:
        i, j = step[0]
        eq = step[1]
        batch_i, batch_j = step[2]
        ti, tj = _np(tensors[i]), _np(tensors[j])
        if len(batch_i) > 1:
            chunks = []
            for k in range(len(batch_i)):
                r = einsum_pair(eq, ti[_np(batch_i[k])], tj[_np(batch_j[k])])
                if step[3]:
                    r = r.reshape(step[3])
                chunks.append(r)
            tensors[j] = []
            tensors[i] = np.concatenate(chunks, axis=0)
        elif len(step) > 3 and len(batch_i) == len(batch_j) == 1:
            tensors[i] = einsum_pair(eq, ti[_np(batch_i[0])], tj[_np(batch_j[0])])
        elif len(step) > 3:
            r = einsum_pair(eq, ti, tj).reshape(step[3])
            if len(batch_i) == 1:
                r = r[_np(batch_i[0])]
            tensors[i] = r
            tensors[j] = []
        else:
            tensors[i] = einsum_pair(eq, ti, tj)
            tensors[j] = []
        if scientific_notation:
            nf = np.abs(tensors[i]).max()
            tensors[i] = tensors[i] / nf
            factor += np.log10(nf)
    if scientific_notation:
        return factor, tensors[i]
    return tensors[i]


def slice_leaves(leaves, slicing_bonds, slicing_indices, slice_id):
    """simulation.py:108-113 with the B1 fix (see module docstring).  `slicing_indices` is
    {bond: [(tid, dim on the un-sliced tensor)]}; bit x of the slice id (MSB first, as
    np.binary_repr(s, S)) fixes bond slicing_bonds[x]."""
    S = len(slicing_bonds)
    per_tensor = {}
    for x, bond in enumerate(slicing_bonds):
        bit = (slice_id >> (S - 1 - x)) & 1
        for tid, dim in slicing_indices[bond]:
            per_tensor.setdefault(tid, {})[dim] = bit
    out = dict(leaves) if isinstance(leaves, dict) else list(leaves)
    for tid, dims in per_tensor.items():
        t = _np(leaves[tid])
        out[tid] = np.ascontiguousarray(t[tuple(dims.get(d, slice(None)) for d in range(t.ndim))])
    return out


def contract_slices(leaves, scheme, pattern, slicing_bonds, slicing_indices, slice_ids,
                    dtype=np.complex64):
    """simulation.py:101-114: sum of the executor's result over the given slice ids."""
    func = tensor_contraction if pattern == "normal" else tensor_contraction_sparse
    leaves = {k: _np(v).astype(dtype) for k, v in (leaves.items() if isinstance(leaves, dict) else enumerate(leaves))}
    acc = None
    for s in slice_ids:
        r = func(slice_leaves(leaves, slicing_bonds, slicing_indices, int(s)), scheme)
        acc = r.copy() if acc is None else acc + r
    return acc
