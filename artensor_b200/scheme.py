"""Scheme compilers: contraction tree -> the step tuples the executors consume.

Replacements for `contraction_scheme` (artensor/contraction.py:23-59) and
`contraction_scheme_sparse` (artensor/contraction.py:208-342).  They take the reference's
`ContractionTree` (its order search is unchanged and out of scope) and emit the SAME tuple
format, so a scheme from here also runs on the reference's executors
(`contraction.py:62-76`, `:132-205`) -- that is how tests/test_scheme_compiler.py checks them --
and a scheme from the reference still runs on this package's executor.

What differs from the reference, deliberately:

  * mode order of every intermediate is `[kept modes of the left operand, in its order] +
    [new modes of the right operand, in its order]`, i.e. the layout a GEMM writes, instead of the
    iteration order of a Python `set` (`contraction.py:33-40,49`);
  * einsum letters are assigned by first appearance, so the strings do not change with
    PYTHONHASHSEED (`contraction.py:17-18` enumerates a `set` of strings);
  * chunks of a batched step cover every row exactly once and are never empty
    (`contraction.py:288-297` can drop rows or emit empty chunks);
  * rows of a batched step are ordered with a STABLE sort (the reference's quicksort order of
    ties is an implementation detail of numpy), so the returned bitstring order is reproducible;
  * which bonds are contracted is looked up in a bond -> carriers map instead of a scan over all
    tensors per bond (`contraction.py:226-235`).

Host-side planning only: no arithmetic happens here.
"""
from math import ceil, log2

import numpy as np
import torch

from .plan import SchemeError

# contraction.py:9-10 uses A-Y + a-y; Z and z are appended (torch.einsum accepts them)
ALPHABET = [chr(c) for c in list(range(65, 90)) + list(range(97, 122))] + ["Z", "z"]

ROWS_LEFT, ROWS_RIGHT, ROWS_SHARED = -1, -2, -3      # labels of the bitstring-row modes (contraction.py:303-325)


def einsum_equation(left, right, out):
    """Two-operand einsum string for label lists; letters in order of first appearance."""
    letter = {}
    for lab in list(left) + list(right) + list(out):
        if lab not in letter:
            if len(letter) == len(ALPHABET):
                raise SchemeError(f"a step with more than {len(ALPHABET)} distinct labels cannot be written as an einsum string")
            letter[lab] = ALPHABET[len(letter)]
    spell = lambda labs: "".join(letter[x] for x in labs)
    return f"{spell(left)},{spell(right)}->{spell(out)}"


def _is_leaf(v):
    return not (v.left and v.right)


def contraction_scheme(ctree):
    """Normal (full-amplitude) mode, contraction.py:23-59: returns (scheme, output_bonds).

    Steps are `((i, j), eq)`; the result of a step replaces slot i, which is always the operand
    whose subtree holds the larger intermediate (`mark_rep_tensor`, contraction_tree.py:305-314),
    so the stem tensor stays in one slot.  Children are emitted larger subtree first."""
    ctree.mark_rep_tensor()
    root = ctree.tree[ctree.all_tensors]
    labels = {}            # id(vertex) -> bond list in the dim order of the tensor the executor holds
    scheme = []
    stack = [(root, False)]
    while stack:
        v, children_done = stack.pop()
        if _is_leaf(v):
            labels[id(v)] = list(ctree.tn.tensor_bonds[v.rep_tensor])
            continue
        if not children_done:
            big, small = (v.left, v.right) if v.left.sc > v.right.sc else (v.right, v.left)
            stack += [(v, True), (small, False), (big, False)]
            continue
        if v.rep_tensor == v.left.rep_tensor:
            keep, other = v.left, v.right
        elif v.rep_tensor == v.right.rep_tensor:
            keep, other = v.right, v.left
        else:
            raise SchemeError("contraction tree: a vertex's representative tensor is not one of its children's")
        a, b = labels.pop(id(keep)), labels.pop(id(other))
        wanted = set(v.contain_bonds)
        in_a = set(a)
        out = [x for x in a if x in wanted] + [x for x in b if x in wanted and x not in in_a]
        if set(out) != wanted:
            raise SchemeError("contraction tree: a vertex keeps a bond neither child carries")
        scheme.append(((keep.rep_tensor, other.rep_tensor), einsum_equation(a, b, out)))
        labels[id(v)] = out
    return scheme, labels[id(root)]


# ------------------------------------------------------------------------------------------------
# sparse-state mode
# ------------------------------------------------------------------------------------------------
def _spread(codes, locs, width):
    """Partial-bitstring codes over `len(locs)` qubits (first qubit = most significant bit) ->
    the same bits at positions `locs` of a `width`-bit code."""
    codes = np.asarray(codes, dtype=np.int64)
    out = np.zeros_like(codes)
    n = len(locs)
    for t, loc in enumerate(locs):
        out |= ((codes >> (n - 1 - t)) & 1) << (width - 1 - loc)
    return out


def _restrict(codes, locs, width):
    """Inverse of `_spread`: the bits at positions `locs` of `width`-bit codes, packed."""
    codes = np.asarray(codes, dtype=np.int64)
    out = np.zeros_like(codes)
    n = len(locs)
    for t, loc in enumerate(locs):
        out |= ((codes >> (width - 1 - loc)) & 1) << (n - 1 - t)
    return out


def _row_of(codes, wanted):
    """Index in `codes` (distinct values) of every entry of `wanted`."""
    order = np.argsort(codes, kind="stable")
    pos = np.searchsorted(codes[order], wanted)
    if np.any(pos >= len(codes)) or np.any(codes[order][np.minimum(pos, len(codes) - 1)] != wanted):
        raise SchemeError("a requested bitstring is not among the rows of an operand")
    return order[pos]


def chunk_bounds(n_rows, rank, sc_target):
    """[begin, end) row ranges of a batched step: every gathered operand of a chunk holds at most
    2^(sc_target - 2) amplitudes where that is possible (contraction.py:288), every row is in exactly
    one chunk, and no chunk is empty."""
    if n_rows <= 0:
        return []
    n_chunks = 2 ** max(0, ceil(log2(n_rows) + rank - (sc_target - 2)))
    length = max(1, -(-n_rows // n_chunks))
    return [(s, min(s + length, n_rows)) for s in range(0, n_rows, length)]


def contraction_scheme_sparse(ctree, bitstrings=None, sc_target=31):
    """Sparse-state mode, contraction.py:208-342: returns (scheme, remaining bonds, bitstrings in
    the row order of the result).

    Final-qubit leaves are `[bit value, in bond]` tensors whose first dim enumerates bit values
    (tensor_network.py:143-145); every tensor that contains final qubits carries one leading
    row mode, and `rows[t] = (final-qubit positions covered, partial-bitstring code of every row)`.
    Steps: 3-tuple `((i, j), eq, batch_seq)` when at most one operand has rows, 5-tuple
    `(..., rshape, next_shape)` when both have: `rshape` set = outer step (all row pairs, i-major,
    then the optional subset `batch_seq[0][0]`), `rshape` None = batched step (gathered row pairs,
    one chunk per entry of `batch_seq[0]` / `batch_seq[1]`)."""
    order = ctree.tree_order_dfs()
    bonds = ctree.tn.tensor_bonds                      # mutated, as in the reference (callers pass a copy)
    final_qubits = ctree.tn.final_qubits
    if isinstance(final_qubits, (set, frozenset)):
        final_qubits = sorted(final_qubits)
    final_qubits = list(final_qubits)
    n_fq = len(final_qubits)
    rows = {t: ([final_qubits.index(t)], np.array([0, 1], dtype=np.int64)) if t in final_qubits else ([], None)
            for t in bonds}
    carriers = {}
    for t, bl in bonds.items():
        for b in bl:
            carriers.setdefault(b, set()).add(t)
    requested = None                                   # full-width codes of the requested bitstrings
    if bitstrings is not None and len(bitstrings):
        requested = np.array([int(s, 2) for s in bitstrings], dtype=np.int64)

    scheme = []
    i = None
    for edge in order:
        i, j = edge
        bi, bj = list(bonds[i]), list(bonds[j])
        in_i = set(bi)
        contracted = {b for b in bj if b in in_i and carriers[b] <= {i, j}}
        kept_i = [b for b in bi if b not in contracted]
        new_i = kept_i + [b for b in bj if b not in contracted and b not in in_i]
        for b in bj:
            carriers[b].discard(j)
            if b not in contracted:
                carriers[b].add(i)
        for b in contracted:
            carriers[b].discard(i)
        bonds[i], bonds[j] = new_i, []

        (pos_i, codes_i), (pos_j, codes_j) = rows[i], rows[j]
        merged = sorted(pos_i + pos_j)
        rshape = None
        shared = False
        if not merged:
            batch_seq = [[torch.tensor([0])], [torch.tensor([0])]]
            codes = None
        elif not pos_j:
            batch_seq = [[torch.arange(len(codes_i))], [torch.tensor([0])]]
            codes = codes_i
        elif not pos_i:
            batch_seq = [[torch.tensor([0])], [torch.arange(len(codes_j))]]
            codes = codes_j
        else:
            if requested is None:
                raise SchemeError("sparse scheme: two operands carry bitstring rows but no bitstrings were given")
            width = len(merged)
            loc_i, loc_j = [merged.index(q) for q in pos_i], [merged.index(q) for q in pos_j]
            wanted = np.unique(_restrict(requested, merged, n_fq))
            if len(wanted) == 2 ** width or width + len(new_i) <= sc_target:
                # outer step: every (row of i, row of j) pair, i-major; then keep the wanted ones
                codes = (_spread(codes_i, loc_i, width)[:, None] + _spread(codes_j, loc_j, width)[None, :]).reshape(-1)
                if len(wanted) != len(codes):
                    keep = np.sort(_row_of(codes, wanted))
                    codes = codes[keep]
                    batch_seq = [[torch.from_numpy(keep)], []]
                else:
                    batch_seq = [[], []]
                rshape = (-1,) + (2,) * len(new_i)
            else:
                # batched step: gather the row pair of every wanted partial bitstring, ordered by the
                # row of the operand with more rows (its gather is then a monotone walk)
                ri = _row_of(codes_i, _restrict(wanted, loc_i, width))
                rj = _row_of(codes_j, _restrict(wanted, loc_j, width))
                perm = np.argsort(ri if len(codes_i) > len(codes_j) else rj, kind="stable")
                ri, rj, codes = ri[perm], rj[perm], wanted[perm]
                bounds = chunk_bounds(len(codes), max(len(bi), len(bj)), sc_target)
                batch_seq = [[torch.from_numpy(ri[s:e].copy()) for s, e in bounds],
                             [torch.from_numpy(rj[s:e].copy()) for s, e in bounds]]
                shared = True
        left = ([ROWS_SHARED] if shared else [ROWS_LEFT]) + bi if pos_i else bi
        right = ([ROWS_SHARED] if shared else [ROWS_RIGHT]) + bj if pos_j else bj
        out = new_i
        if shared:
            out = [ROWS_SHARED] + out
        else:
            out = ([ROWS_LEFT] if pos_i else []) + ([ROWS_RIGHT] if pos_j else []) + out
        eq = einsum_equation(left, right, out)
        if pos_i and pos_j:
            scheme.append((edge, eq, batch_seq, rshape, (len(codes),) + (2,) * len(new_i)))
        else:
            scheme.append((edge, eq, batch_seq))
        rows[i] = (merged, codes)
        rows[j] = ([], None)

    if i is None:
        raise SchemeError("empty contraction order")
    final_codes = rows[i][1]
    ordered = [] if final_codes is None else [np.binary_repr(int(c), n_fq) for c in final_codes]
    return scheme, bonds[i], ordered
