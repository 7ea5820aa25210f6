"""Plan compiler, front end: a reference-format contraction scheme -> flat step records.

Input is exactly what `artensor.contraction.contraction_scheme` (`contraction.py:23-59`) or
`contraction_scheme_sparse` (`contraction.py:208-342`) emit and what the reference executors
(`contraction.py:62-76`, `:132-205`) consume.  This module only *describes* each step; it does
no arithmetic.  Everything in a circuit tensor network is bits:

  * every bond has extent 2 (`circuit.py:117,130`), so a rank-r tensor is a flat 2^r array
    and a mode order is an assignment of modes to address bits;
  * in sparse mode a tensor may additionally carry ONE leading "row" mode (the bitstring
    batch, `contraction.py:219-220`) of arbitrary extent, always kept outermost.

Each step is normalised to

    C[b][m..., n..., h...] = sum_k  A[ra[b]][m..., k..., h...] * B[rb[b]][k..., n..., h...]

where b enumerates output rows, (ra[b], rb[b]) are source rows (or None when the operand has
no row mode), m/n/k/h are sets of bit modes (left-only, right-only, contracted, shared-kept).
The three sparse step kinds of `tensor_contraction_sparse` reduce to a choice of (ra, rb):

  plain   (`contraction.py:189-191`)   rows (if any) pass through:      ra = arange / None
  outer   (`contraction.py:180-188`)   rows are i-major pairs:          ra = b // Rj, rb = b % Rj,
                                       then the optional row subset     b -> remain[b]
  batched (`contraction.py:140-179`)   gathered pairs, chunks concatenated in order
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


class SchemeError(ValueError):
    pass


@dataclass
class TensorInfo:
    """Logical description of one tensor slot: `modes` are globally unique mode ids in the
    reference's (logical) dim order, excluding the row mode; `rows` is the extent of the
    leading row mode or None."""
    modes: Tuple[int, ...]
    rows: Optional[int] = None

    @property
    def rank(self):
        return len(self.modes)

    @property
    def numel(self):
        return (self.rows or 1) << len(self.modes)


@dataclass
class Step:
    index: int
    i: int                      # slot of the left operand == output slot
    j: int                      # slot of the right operand
    kind: str                   # 'plain' | 'outer' | 'batched'
    a: TensorInfo
    b: TensorInfo
    c: TensorInfo
    m_modes: Tuple[int, ...]    # modes only in A (kept)
    n_modes: Tuple[int, ...]    # modes only in B (kept)
    k_modes: Tuple[int, ...]    # contracted
    h_modes: Tuple[int, ...]    # in A, B and C (hyper-edge batch; absent in circuit TNs)
    ra: Optional[np.ndarray]    # int64[nb] source row in A per output row, or None
    rb: Optional[np.ndarray]
    out_rows: Optional[int]     # nb when C has a row mode, else None
    chunks: int = 1             # how many chunks the reference used (informational)

    @property
    def nb(self):
        return self.out_rows or 1

    @property
    def flops(self):
        """8*B*M*N*K (complex MAC = 8 real flops); K=0 outer products count 6*M*N."""
        mn = self.nb << (len(self.m_modes) + len(self.n_modes) + len(self.h_modes))
        if len(self.k_modes) == 0:
            return 6 * mn
        return 8 * mn << len(self.k_modes)

    @property
    def bytes_c64(self):
        """8*(|A|+|B|+|C|) with gathered row counts."""
        ra = self.nb if self.ra is not None else 1
        rb = self.nb if self.rb is not None else 1
        if self.kind == 'plain' or self.kind == 'outer':
            ra = self.a.rows or 1
            rb = self.b.rows or 1
        return 8 * ((ra << self.a.rank) + (rb << self.b.rank) + (self.nb << self.c.rank))


def parse_eq(eq):
    try:
        lhs, out = eq.split('->')
        la, lb = lhs.split(',')
    except ValueError:
        raise SchemeError(f"not a two-operand einsum equation: {eq!r}")
    for s in (la, lb, out):
        if len(set(s)) != len(s):
            raise SchemeError(f"repeated label inside one operand is not supported: {eq!r}")
    return la, lb, out


def _shape_of(t):
    return tuple(int(x) for x in t.shape)


class SchemeParser:
    """Walks a scheme once, tracking the logical mode list of every tensor slot."""

    def __init__(self, leaf_shapes: Dict[int, Tuple[int, ...]], sparse: bool):
        self.sparse = sparse
        self._next_mode = 0
        self.slots: Dict[int, Optional[TensorInfo]] = {}
        self.leaf_info: Dict[int, TensorInfo] = {}
        self.leaf_shapes = dict(leaf_shapes)
        self._row_hint: Dict[int, bool] = {}

    def _new_modes(self, n):
        out = tuple(range(self._next_mode, self._next_mode + n))
        self._next_mode += n
        return out

    def _leaf(self, tid, has_row):
        """Materialise slot `tid` from its leaf shape at first use."""
        if tid in self.slots:
            info = self.slots[tid]
            if info is None:
                raise SchemeError(f"tensor {tid} was already consumed")
            return info
        if tid not in self.leaf_shapes:
            raise SchemeError(f"scheme refers to tensor {tid} which is not among the leaves")
        shape = self.leaf_shapes[tid]
        if has_row:
            if len(shape) < 1:
                raise SchemeError(f"tensor {tid} needs a row mode but is a scalar")
            rows, bits = shape[0], shape[1:]
        else:
            rows, bits = None, shape
        if any(e != 2 for e in bits):
            raise SchemeError(f"tensor {tid}: every bond must have extent 2, got shape {shape}")
        info = TensorInfo(self._new_modes(len(bits)), rows)
        self.slots[tid] = info
        self.leaf_info[tid] = info
        return info

    def parse(self, scheme) -> List[Step]:
        steps = []
        for idx, s in enumerate(scheme):
            if len(s) not in (2, 3, 5):
                raise SchemeError(f"step {idx}: unexpected tuple length {len(s)}")
            if not self.sparse and len(s) != 2:
                raise SchemeError(f"step {idx}: sparse step tuple passed to the normal executor")
            if self.sparse and len(s) == 2:
                raise SchemeError(f"step {idx}: normal step tuple passed to the sparse executor")
            (i, j), eq = s[0], s[1]
            la, lb, lo = parse_eq(eq)
            a = self._leaf(i, self._leaf_has_row(i, s, 0))
            b = self._leaf(j, self._leaf_has_row(j, s, 1))
            steps.append(self._one(idx, s, i, j, la, lb, lo, a, b))
        return steps

    # -- which leaves carry a row mode ------------------------------------------------
    def _leaf_has_row(self, tid, step, side):
        """Decided at the first use of a leaf (results inherit it from their operands).

        `contraction_scheme_sparse` gives a tensor a leading row mode iff it contains final
        qubits (`contraction.py:219-220`).  Seen from the scheme alone: a leaf carries rows iff
        it enters a 5-tuple step (both operands carry rows, `contraction.py:327-333`), or the
        step's index list for its side enumerates more than one row (`contraction.py:249-254`
        emit arange(rows) for the row side and [0] for the other), or its dim-0 extent is not 2
        (cannot be a bond)."""
        if tid in self.slots or not self.sparse:
            return False
        shape = self.leaf_shapes.get(tid)
        if shape is None:
            return False
        if len(step) == 5:
            return True
        if len(shape) >= 1 and shape[0] != 2:
            return True
        lists = step[2][side]
        return len(lists) == 1 and len(lists[0]) > 1

    # -- one step ------------------------------------------------------------------------
    def _one(self, idx, s, i, j, la, lb, lo, a: TensorInfo, b: TensorInfo) -> Step:
        a_has_row, b_has_row = a.rows is not None, b.rows is not None
        if len(la) != a.rank + a_has_row or len(lb) != b.rank + b_has_row:
            raise SchemeError(
                f"step {idx}: equation {s[1]!r} does not match operand ranks "
                f"({a.rank}+{int(a_has_row)}, {b.rank}+{int(b_has_row)})")
        five = len(s) == 5
        if five and not (a_has_row and b_has_row):
            raise SchemeError(f"step {idx}: 5-tuple step but an operand has no row mode")
        row_a = la[0] if a_has_row else None
        row_b = lb[0] if b_has_row else None
        bits_a = la[1:] if a_has_row else la
        bits_b = lb[1:] if b_has_row else lb
        lab2mode = {}
        for lab, mode in zip(bits_a, a.modes):
            lab2mode[lab] = mode
        k_modes, h_modes, n_labels = [], [], []
        b_mode_of = {}
        for lab, mode in zip(bits_b, b.modes):
            b_mode_of[lab] = mode
        out_set = set(lo)
        shared = [lab for lab in bits_a if lab in b_mode_of]
        for lab in shared:
            if lab in out_set:
                h_modes.append(lab)
            else:
                k_modes.append(lab)
        m_labels = [lab for lab in bits_a if lab not in b_mode_of]
        n_labels = [lab for lab in bits_b if lab not in lab2mode]
        for lab in m_labels + n_labels:
            if lab not in out_set:
                raise SchemeError(f"step {idx}: label {lab!r} is summed inside one operand (unsupported)")
        # row handling --------------------------------------------------------------
        kind, ra, rb, out_rows, chunks = 'plain', None, None, None, 1
        out_row_labels = 0
        if five:
            bi, bj = s[2]
            if s[3] is None:          # batched: shared row label (-3)
                kind = 'batched'
                if row_a != row_b or not lo or lo[0] != row_a:
                    raise SchemeError(f"step {idx}: batched step without a shared leading row label")
                if len(bi) != len(bj) or len(bi) < 1:
                    raise SchemeError(f"step {idx}: batched step with mismatched chunk lists")
                ra = np.concatenate([np.asarray(c, dtype=np.int64).reshape(-1) for c in bi])
                rb = np.concatenate([np.asarray(c, dtype=np.int64).reshape(-1) for c in bj])
                if len(ra) != len(rb):
                    raise SchemeError(f"step {idx}: batched step with mismatched index lists")
                chunks = len(bi)
                out_rows = int(len(ra))
                out_row_labels = 1
            else:                     # outer: (-1, -2) merged by reshape, optional subset
                kind = 'outer'
                if len(lo) < 2 or lo[0] != row_a or lo[1] != row_b:
                    raise SchemeError(f"step {idx}: outer step must emit both row labels first")
                full = np.arange(a.rows * b.rows, dtype=np.int64)
                if len(bi) == 1:
                    full = np.asarray(bi[0], dtype=np.int64).reshape(-1)
                elif len(bi) > 1:
                    raise SchemeError(f"step {idx}: outer step with a chunked index list")
                ra, rb = full // b.rows, full % b.rows
                out_rows = int(len(full))
                out_row_labels = 2
            if out_rows and (ra.max() >= a.rows or rb.max() >= b.rows or ra.min() < 0 or rb.min() < 0):
                raise SchemeError(f"step {idx}: row index out of range")
            if len(s[4]) and int(s[4][0]) != out_rows:
                raise SchemeError(
                    f"step {idx}: index lists cover {out_rows} rows but the scheme expects {s[4][0]} "
                    f"(invalid chunking, see SURVEY 4.3-B2)")
        else:
            if a_has_row and b_has_row:
                raise SchemeError(f"step {idx}: both operands carry rows but the step is not a 5-tuple")
            if a_has_row:
                if not lo or lo[0] != row_a:
                    raise SchemeError(f"step {idx}: row label of the left operand must lead the output")
                ra, out_rows, out_row_labels = np.arange(a.rows, dtype=np.int64), a.rows, 1
            elif b_has_row:
                if not lo or lo[0] != row_b:
                    raise SchemeError(f"step {idx}: row label of the right operand must lead the output")
                rb, out_rows, out_row_labels = np.arange(b.rows, dtype=np.int64), b.rows, 1
        # output modes in the reference's logical order -------------------------------
        out_bits = lo[out_row_labels:]
        c_modes = []
        new_n = {}
        for lab in out_bits:
            if lab in lab2mode:
                c_modes.append(lab2mode[lab])
            elif lab in b_mode_of:
                c_modes.append(b_mode_of[lab])
            else:
                raise SchemeError(f"step {idx}: output label {lab!r} appears in no operand")
        if len(out_bits) != len(m_labels) + len(n_labels) + len(h_modes):
            raise SchemeError(f"step {idx}: output labels do not match the kept operand labels")
        c = TensorInfo(tuple(c_modes), out_rows)
        # shared-kept labels keep A's identity; make B's alias resolvable
        st = Step(
            index=idx, i=i, j=j, kind=kind, a=a, b=b, c=c,
            m_modes=tuple(lab2mode[l] for l in m_labels),
            n_modes=tuple(b_mode_of[l] for l in n_labels),
            k_modes=tuple(lab2mode[l] for l in k_modes),
            h_modes=tuple(lab2mode[l] for l in h_modes),
            ra=ra, rb=rb, out_rows=out_rows, chunks=chunks,
        )
        st.k_modes_b = tuple(b_mode_of[l] for l in k_modes)
        st.h_modes_b = tuple(b_mode_of[l] for l in h_modes)
        self.slots[i] = c
        self.slots[j] = None
        return st
