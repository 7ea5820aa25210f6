"""TEST INFRASTRUCTURE: a numpy interpreter of the native operation records.

Lets the CPU test-suite check the whole host side (scheme parser, hoisting, layouts, arena
offsets, leaf table, row tables) against the oracle without a GPU: it executes exactly the
records `ContractionPlan` would hand to libtnc_b200.so, with the semantics documented in
include/tnc_b200.h.  Never used by the product path.
"""
import numpy as np

from artensor_b200 import _native as N


def _view(arena, t):
    n = t.rows << t.rank
    return arena[t.offset // 8: t.offset // 8 + n]


def _deposit(idx, positions):
    out = np.zeros_like(idx)
    for i, p in enumerate(positions):
        out |= ((idx >> i) & 1) << int(p)
    return out


def slice_deps(plan):
    """Slice-id bits (LSB-based) behind every SLICE-phase operation, derived from the leaf records the way
    tnc_plan_finalize does for TNC_OPT_SLICE_REUSE (leaf loads and the accumulate: every bit)."""
    S = plan.n_sliced
    at, deps = {}, []
    for kind, rec in plan.ops[N.TNC_PHASE_SLICE]:
        d = (1 << S) - 1
        if kind == "leaves":
            for L in rec:
                at[L.dst.offset] = sum(1 << (S - 1 - L.sliced_bond[q]) for q in range(L.n_sliced))
        elif kind == "einsum":
            d = at.get(rec.a.offset, 0) | at.get(rec.b.offset, 0)
            at[rec.c.offset] = d
        elif kind == "permute":
            d = at.get(rec.src.offset, 0)
            at[rec.dst.offset] = d
        deps.append(d)
    # TNC_EINSUM_RUN_WITH_READER: such an operation runs whenever the (later) operation that reads its result does
    ops = plan.ops[N.TNC_PHASE_SLICE]
    for n in range(len(ops) - 1, -1, -1):
        kind, rec = ops[n]
        if kind != "einsum" or not (rec.flags & N.TNC_EINSUM_RUN_WITH_READER):
            continue
        deps[n] = (1 << S) - 1
        for m in range(n + 1, len(ops)):
            k2, r2 = ops[m]
            if (k2 == "einsum" and rec.c.offset in (r2.a.offset, r2.b.offset)) or \
               (k2 == "permute" and r2.src.offset == rec.c.offset) or (k2 == "accum" and r2.src.offset == rec.c.offset):
                deps[n] = deps[m]
                break
    return deps


def run_plan(plan, leaf_blob, slice_ids, reuse=False, poison=False):
    """Execute plan.ops on a host arena; returns the accumulator (logical result order).
    reuse: TNC_OPT_SLICE_REUSE semantics over the (consecutive) slice_ids -- after the first slice an operation
    runs only when a slice-id bit behind it changed.  poison: after every slice, everything outside the ONCE and
    KEEP regions is overwritten with NaN (what recycled memory may hold by the time a later slice reads it)."""
    assert plan.dtype == N.TNC_C64
    arena = np.zeros(plan.workspace_bytes // 8, dtype=np.complex64)
    out = np.zeros(int(np.prod(plan.out_shape)) if plan.out_shape else 1, dtype=np.complex64)
    S = plan.n_sliced

    def run(kind, rec, sid):
        if kind == "leaves":
            for L in rec:
                dst = _view(arena, L.dst)
                e = np.arange(L.dst.rows << L.dst.rank, dtype=np.int64)
                row, eb = e >> L.dst.rank, e & ((1 << L.dst.rank) - 1)
                so = _deposit(eb, [L.keep_pos[i] for i in range(L.dst.rank)])
                for s in range(L.n_sliced):
                    bit = (sid >> (S - 1 - L.sliced_bond[s])) & 1
                    so |= bit << int(L.sliced_pos[s])
                dst[:] = leaf_blob[L.src_offset + (row << L.src_rank) + so]
        elif kind == "einsum":
            E = rec
            if E.algo == N.TNC_ALGO_TC:
                # the scratch panels must not overlap any operand of the step, and the output
                # must have the layout the tensor-core GEMM writes (n modes lowest)
                lo, hi = E.scratch_offset, E.scratch_offset + E.scratch_bytes
                assert hi <= plan.arena_bytes and lo % 1024 == 0
                for t in (E.a, E.b, E.c):
                    assert t.offset + ((t.rows << t.rank) * 8) <= lo or t.offset >= hi, "scratch overlaps an operand"
                assert sorted(E.n_c[i] for i in range(E.n_n)) == list(range(E.n_n))
                assert E.n_k >= 1 and E.n_n >= 1 and E.n_h == 0
            if E.algo == N.TNC_ALGO_SKINNY:
                assert sorted(E.n_c[i] for i in range(E.n_n)) == list(range(E.n_n)) and E.n_h == 0
                assert 2 <= E.n_k <= 6 and 1 <= E.n_n <= 7 and E.n_m >= 7 and not (E.n_k == 6 and E.n_n > 6)
                assert (E.rows_b == N.TNC_ROWS_NONE or E.nb == 1 or (E.rows_a == N.TNC_ROWS_NONE and E.rows_b == N.TNC_ROWS_IDENTITY)
                        or (E.flags & N.TNC_EINSUM_OUTER_ROWS))
            if E.algo == N.TNC_ALGO_STEM:
                assert sorted(E.n_c[i] for i in range(E.n_n)) == list(range(E.n_n)) and E.n_h == 0
                assert (8 << (E.n_k + E.n_n)) + (4 << E.n_k) <= 60 * 1024
            if E.flags & N.TNC_EINSUM_OUTER_ROWS:
                assert E.nb == E.a.rows * E.b.rows and E.rows_a != N.TNC_ROWS_NONE and E.rows_b != N.TNC_ROWS_NONE
            A, B, Cv = _view(arena, E.a), _view(arena, E.b), _view(arena, E.c)
            e = np.arange(E.nb << E.c.rank, dtype=np.int64)
            row, cb = e >> E.c.rank, e & ((1 << E.c.rank) - 1)
            oa = np.zeros_like(e)
            ob = np.zeros_like(e)
            for i in range(E.n_m):
                oa |= ((cb >> int(E.m_c[i])) & 1) << int(E.m_a[i])
            for i in range(E.n_n):
                ob |= ((cb >> int(E.n_c[i])) & 1) << int(E.n_b[i])
            for i in range(E.n_h):
                bit = (cb >> int(E.h_c[i])) & 1
                oa |= bit << int(E.h_a[i])
                ob |= bit << int(E.h_b[i])

            def rows(mode):
                if mode == N.TNC_ROWS_NONE:
                    return np.zeros_like(row)
                if mode == N.TNC_ROWS_IDENTITY:
                    return row
                return plan.tables[mode].astype(np.int64)[row]
            if E.flags & N.TNC_EINSUM_OUTER_ROWS:
                assert np.array_equal(rows(E.rows_a), row // E.b.rows) and np.array_equal(rows(E.rows_b), row % E.b.rows)
            if E.flags & N.TNC_EINSUM_OUTER_PAIRS:      # every (A row, B row) pair exactly once, any order
                assert E.nb == E.a.rows * E.b.rows and not (E.flags & N.TNC_EINSUM_OUTER_ROWS)
                pairs = (rows(E.rows_a) * E.b.rows + rows(E.rows_b))[::1 << E.c.rank]
                assert len(np.unique(pairs)) == E.nb
            oa += rows(E.rows_a) << E.a.rank
            ob += rows(E.rows_b) << E.b.rank
            k = np.arange(1 << E.n_k, dtype=np.int64)
            ka = _deposit(k, [E.k_a[i] for i in range(E.n_k)])
            kb = _deposit(k, [E.k_b[i] for i in range(E.n_k)])
            acc = np.zeros(len(e), dtype=np.complex64)
            step = max(1, (1 << 22) // len(k))
            for lo in range(0, len(e), step):
                hi = min(len(e), lo + step)
                acc[lo:hi] = (A[oa[lo:hi, None] + ka[None, :]] * B[ob[lo:hi, None] + kb[None, :]]).sum(axis=1)
            Cv[:] = acc
        elif kind == "permute":
            P = rec
            src, dst = _view(arena, P.src), _view(arena, P.dst)
            e = np.arange(P.src.rows << P.src.rank, dtype=np.int64)
            q = e & ((1 << P.src.rank) - 1)
            s = _deposit(q, [P.perm[i] for i in range(P.src.rank)])
            dst[:] = src[((e >> P.src.rank) << P.src.rank) + s]
        elif kind == "accum":
            A = rec
            src = _view(arena, A.src)
            e = np.arange(A.src.rows << A.src.rank, dtype=np.int64)
            q = e & ((1 << A.src.rank) - 1)
            d = _deposit(q, [A.out_pos[i] for i in range(A.src.rank)])
            out[((e >> A.src.rank) << A.src.rank) + d] += src
        else:
            raise ValueError(kind)

    for kind, rec in plan.ops[N.TNC_PHASE_ONCE]:
        run(kind, rec, 0)
    deps = slice_deps(plan) if reuse else None
    prev = None
    for sid in slice_ids:
        first = prev is None
        for n, (kind, rec) in enumerate(plan.ops[N.TNC_PHASE_SLICE]):
            if reuse and not first and kind in ("einsum", "permute") and not (deps[n] & (int(sid) ^ prev)):
                continue
            run(kind, rec, int(sid))
        prev = int(sid)
        if poison:
            lo, hi = plan.recycled_range
            arena[lo // 8: hi // 8] = np.nan
    return out.reshape(plan.out_shape)
