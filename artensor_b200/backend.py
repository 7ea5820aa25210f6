"""Plan compiler, back end: parsed steps -> operations of the native library.

Host logic only (numpy + ctypes struct filling); the arithmetic happens in libtnc_b200.so.
What is decided here:

  * the leaf table: where each leaf sits in the packed leaf blob and how its sliced bonds
    are fixed per slice id (replaces `select(ind, bit).clone()`, artensor/simulation.py:110-113);
  * hoisting: steps that depend on no sliced bond run once per execute call instead of once
    per slice (the reference recomputes them for every slice, simulation.py:107-114);
  * the physical mode order ("layout") of every intermediate, and which algorithm runs a step;
  * arena offsets with liveness-based reuse (the reference lets torch allocate per einsum).
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as N
from .plan import SchemeParser, Step, TensorInfo, SchemeError

ALIGN = 1024
REGION_KEEP = 2     # third arena region (after ONCE and SLICE): slice-phase intermediates that outlive a slice (slice_reuse)


class Arena:
    """First-fit free-list allocator run at plan time; the high-water mark is the workspace."""

    def __init__(self):
        self.free: List[Tuple[int, int]] = []   # (offset, size), sorted, coalesced
        self.top = 0
        self.high = 0

    def alloc(self, nbytes):
        size = max(ALIGN, (nbytes + ALIGN - 1) // ALIGN * ALIGN)
        for idx, (off, sz) in enumerate(self.free):
            if sz >= size:
                if sz == size:
                    self.free.pop(idx)
                else:
                    self.free[idx] = (off + size, sz - size)
                return off, size
        # grow: extend a trailing free block if there is one
        if self.free and self.free[-1][0] + self.free[-1][1] == self.top:
            off, sz = self.free.pop()
            self.top = off + size
        else:
            off = self.top
            self.top += size
        self.high = max(self.high, self.top)
        return off, size

    def release(self, off, size):
        self.free.append((off, size))
        self.free.sort()
        merged = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self.free = merged


@dataclass
class Buf:
    """A tensor resident in the arena: `pos[mode]` is the bit position of each mode.
    Buffers of slice-invariant data live in the ONCE region, everything else in the SLICE
    region that starts where the ONCE region ends (bases are fixed after allocation); with
    PlanOptions.slice_reuse, slice-phase results that must survive until a LATER slice (their
    consumer depends on more sliced bonds than they do) get a region of their own, KEEP, where
    nothing is ever laid over them."""
    region: int           # N.TNC_PHASE_ONCE, N.TNC_PHASE_SLICE or REGION_KEEP
    rel_offset: int
    size: int
    info: TensorInfo
    pos: Dict[int, int]
    is_leaf: bool
    base: List[int] = None   # shared [once_base, slice_base, keep_base], filled in at the end
    deps: frozenset = frozenset()   # sliced bonds (indices into slicing_bonds) the contents depend on

    @property
    def dependent(self):
        return self.region != N.TNC_PHASE_ONCE

    def tensor(self):
        return N.TncTensor(self.base[self.region] + self.rel_offset, self.info.rank, self.info.rows or 1)


def logical_positions(info: TensorInfo):
    r = info.rank
    return {m: r - 1 - i for i, m in enumerate(info.modes)}


@dataclass
class PlanOptions:
    """Algorithm selection per step:
      * tcgen05 GEMM for compute-bound steps: flops >= tc_min_flops and arithmetic intensity
        (flops / algorithmic bytes) >= tc_min_intensity;
      * the streaming tcgen05 kernel ("skinny": A read in place, small right operand resident in
        shared memory) for steps of its shape class whose left operand has >= skinny_min_elems
        amplitudes -- it replaces both the pack + GEMM lowering and the fp32 streaming kernel there;
      * the streaming fp32 kernel for the other large steps (HBM-bound "stem" steps);
      * the generic kernel for the hundreds of tiny steps (< stem_min_elems output elements).
    tc_precision: operand precision of the tensor-core steps (include/tnc_b200.h tnc_tc_precision):
      "3xf16" / "3xtf32" are fp32-accurate split products (complex64 mode), "f16" is the
      reduced-precision complex-half mode."""
    tc_min_flops: float = 1e8
    tc_min_intensity: float = 24.0
    stem_min_elems: int = 1 << 12
    skinny_min_elems: int = 1 << 20
    # rows of <= 8 outputs leave the tensor-core kernel latency-bound (measured on n53 m20: 1.4-3.0
    # TB/s against 2.9-4.8 TB/s of the fp32 streaming kernel, whose FMA load is light there)
    skinny_min_n: int = 4
    swap_operands: bool = True
    hoist: bool = True
    tc_precision: str = "3xf16"
    # Replay the slice phase as one CUDA graph per slice (TNC_OPT_CUDA_GRAPH).  None = automatic:
    # on for sliced plans whose slice is light enough (< graph_max_flops executed flops) for the
    # launch gaps of its ~100 kernels to matter (n53 m12: 3.3 ms per slice over 140 launches)
    cuda_graph: Optional[bool] = None
    graph_max_flops: float = 5e12
    # tensor-core operands written by a streaming / GEMM kernel get their largest magnitude (the fp16 operand scale)
    # reduced by that kernel's epilogue instead of a separate pass over the tensor (TNC_OPT_FUSE_AMAX); results are
    # bit-identical either way
    fuse_amax: bool = True
    # Cross-slice reuse (TNC_OPT_SLICE_REUSE): inside one execute call over consecutive slice ids, a step is
    # contracted again only when a sliced bond it depends on changed since the previous slice -- the reference's
    # loop (simulation.py:107-114) recomputes the whole tree for every slice although most of a deep tree depends
    # on a few of the sliced bonds only.  Results are bit-identical; intermediates whose consumer depends on more
    # bonds than they do keep a region of their own for the whole call (a larger workspace).
    slice_reuse: bool = False
    # Upper bound (bytes) for the KEEP region of slice_reuse, None = unbounded.  When the results that would have to be
    # kept exceed it, the planner TIES some of them to their reader instead (TNC_EINSUM_RUN_WITH_READER): such a step is
    # contracted again whenever its reader is, its result recycles -- less memory, more recomputation; the ties that
    # cost the least modelled time per byte freed are chosen first.
    keep_budget_bytes: Optional[int] = None

    def __post_init__(self):
        if self.tc_precision not in N.TC_PRECISIONS:
            raise ValueError(f"tc_precision {self.tc_precision!r}: expected one of {sorted(N.TC_PRECISIONS)}")


def tc_scratch_bytes(st: Step):
    """Mirror of tnc_einsum_tc_scratch_bytes(): hi/lo panels of A' and of the expanded B'."""
    al = lambda x: (x + 1023) // 1024 * 1024
    nb_a = st.nb if st.ra is not None else 1
    nb_b = st.nb if st.rb is not None else 1
    if full_outer(st) or outer_pairs(st):
        nb_a, nb_b = st.a.rows, st.b.rows
    a_panel = al((nb_a << (len(st.m_modes) + len(st.k_modes))) * 8)
    b_panel = al((nb_b << (len(st.n_modes) + len(st.k_modes))) * 16)
    return 2 * a_panel + 2 * b_panel + 1024      # + the amax / barrier words of the step


def full_outer(st: Step):
    """An outer step that keeps every (row of A, row of B) pair, A-major."""
    return (st.kind == "outer" and st.a.rows is not None and st.b.rows is not None and
            st.nb == st.a.rows * st.b.rows and np.array_equal(st.ra, np.arange(st.nb) // st.b.rows) and
            np.array_equal(st.rb, np.arange(st.nb) % st.b.rows))


def outer_pairs(st: Step):
    """A step whose output rows are every (row of A, row of B) pair exactly once, in an order other
    than A-major: the reference's batched ("cat") steps when the wanted bitstrings are the full
    product of the operands' rows (contraction.py:272-300 sorts the pairs by the row of the larger
    operand).  TNC_EINSUM_OUTER_PAIRS lets the tensor-core path pack each operand row once."""
    if st.ra is None or st.rb is None or st.a.rows is None or st.b.rows is None or st.nb < 2:
        return False
    if st.nb != st.a.rows * st.b.rows or full_outer(st):
        return False
    pair = np.asarray(st.ra, dtype=np.int64) * st.b.rows + np.asarray(st.rb, dtype=np.int64)
    return len(np.unique(pair)) == st.nb


def tc_eligible(st: Step, precision="3xtf32"):
    """Mirror of the checks in tc_gemm_create(): the fp16 panels need rows of >= 16 bytes (TMA),
    i.e. at least 4 complex k per row."""
    min_k = 1 if precision == "3xtf32" else 2
    return len(st.k_modes) >= min_k and len(st.n_modes) >= 1 and len(st.h_modes) == 0


def tc_uses_3m(st: Step, precision="3xf16"):
    """Mirror of the 3M (Karatsuba) selection in tc_gemm_create(): the 3xF16 precision, whole 64-k blocks
    (k >= 6 bits), >= 128 complex columns (n >= 7 bits), whole 256-row pair tiles, B's rows not
    folded into N.  Such a step issues 0.75 real tensor-core products per useful complex one
    (2.25 with the hi/lo split) instead of 1 (3)."""
    if precision != "3xf16":
        return False
    if os.environ.get("TNC_EXPERIMENTS", "0") not in ("", "0") and "0" in (os.environ.get("TNC_TC_3M"), os.environ.get("TNC_TC_2CTA")):
        return False                      # experiment knobs of the library (include/tnc_b200.h)
    k, n, m = len(st.k_modes), len(st.n_modes), len(st.m_modes)
    if not tc_eligible(st, precision) or k < 6 or n < 7 or m < 7:
        return False
    outer = full_outer(st) or outer_pairs(st)
    if outer:
        return False                      # B's rows are folded into N (n >= 4): plain [row of B][n][k] panel
    fold_rows = st.rb is None or st.nb == 1
    nb_a = st.nb if st.ra is not None else 1
    rows = (nb_a << m) if fold_rows else (1 << m)
    return rows % 256 == 0


def stem_eligible(st: Step):
    """Mirror of stem_supported() in csrc/stem.cu: B[k][n] and the k offsets must fit shared memory."""
    k, n = len(st.k_modes), len(st.n_modes)
    return len(st.h_modes) == 0 and k <= 12 and n <= 12 and (8 << (k + n)) + (4 << k) <= 60 * 1024


def operand_elems(st: Step):
    """Amplitudes the step reads from each operand (gathered rows counted)."""
    ea = (st.nb if st.ra is not None else 1) << st.a.rank
    eb = (st.nb if st.rb is not None else 1) << st.b.rank
    return ea, eb


def should_swap(st: Step, min_elems=1 << 12):
    """C = A.B is lowered as C^T = B^T.A^T when the right operand is the larger one: every kernel
    streams (or packs without expansion) its LEFT operand and keeps the right one small -- the
    tensor-core lowering even doubles the right operand (B' expansion).  The output layout is the
    planner's choice anyway, so exchanging the roles costs nothing.  Outer steps keep their order
    (their row pairs are A-major by construction, contraction.py:180-185)."""
    if st.kind == "outer" or len(st.h_modes) != 0:
        return False
    ea, eb = operand_elems(st)
    if max(ea, eb) < min_elems:
        return False
    return eb > ea or (eb == ea and len(st.n_modes) > len(st.m_modes))


def swapped(st: Step):
    """The same step with the operand roles exchanged (a <-> b, m <-> n, ra <-> rb)."""
    import copy
    s = copy.copy(st)
    s.a, s.b = st.b, st.a
    s.m_modes, s.n_modes = st.n_modes, st.m_modes
    s.k_modes, s.k_modes_b = st.k_modes_b, st.k_modes
    s.ra, s.rb = st.rb, st.ra
    s.is_swapped = True
    return s


def skinny_fold(st: Step):
    """Mirror of skinny_fold() in csrc/skinny.cu: rows of the right operand one launch handles
    (1: a single right operand; > 1: its rows are folded into N; 0: not supported)."""
    if st.rb is None or st.nb == 1:
        return 0 if (st.ra is None and st.nb != 1) else 1
    fold = 0
    if st.ra is None and st.kind == "plain":
        fold = st.nb
    elif full_outer(st):
        fold = st.b.rows
    k, n = len(st.k_modes), len(st.n_modes)
    if fold < 1 or k > 5 or n < 2 or (fold << (n + 1)) > 256:
        return 0
    return fold


def skinny_eligible(st: Step, precision, min_n=1):
    """Mirror of skinny_supported() in csrc/skinny.cu; `min_n` is the planner's own cut-off
    (PlanOptions.skinny_min_n) on the outputs per row of A read (folded rows count)."""
    k, n, m = len(st.k_modes), len(st.n_modes), len(st.m_modes)
    if precision == "3xtf32" or len(st.h_modes) != 0:
        return False
    fold = skinny_fold(st)
    if fold == 0:
        return False
    # rows of K >= 32 carry enough bytes on the A side to run one output bit shorter
    low = max(1, min_n - 1) if k >= 5 else max(1, min_n)
    if not (2 <= k <= 6 and 1 <= n <= 7 and m >= 7) or (k == 6 and n > 6):
        return False
    return (fold << n) >= (1 << low)


class ContractionPlan:
    """A compiled scheme.  Immutable after construction; execute() is stream-ordered."""

    def __init__(self, scheme, leaf_shapes: Dict[int, Tuple[int, ...]], sparse: bool, *,
                 slicing_bonds=(), slicing_indices=None, dtype="c64", options: Optional[PlanOptions] = None,
                 build_native=True):
        self.options = options or PlanOptions()
        self.sparse = sparse
        if dtype != "c64":
            raise ValueError(f"dtype {dtype!r}: tensors are stored as complex64 (the complex-half mode is "
                             f"PlanOptions.tc_precision='f16', an operand precision)")
        self.dtype = N.TNC_C64
        self.elem_bytes = 8
        self.slicing_bonds = list(slicing_bonds)
        self.n_sliced = len(self.slicing_bonds)
        if self.n_sliced > 62:
            raise SchemeError(f"{self.n_sliced} sliced bonds: slice ids do not fit 63 bits")
        slicing_indices = slicing_indices or {}
        # per leaf: {dim (un-sliced tensor dim): bond index x}
        self.leaf_sliced: Dict[int, Dict[int, int]] = {}
        for x, bond in enumerate(self.slicing_bonds):
            for tid, dim in slicing_indices[bond]:
                self.leaf_sliced.setdefault(int(tid), {})[int(dim)] = x
        self.full_leaf_shapes = {int(k): tuple(int(e) for e in v) for k, v in leaf_shapes.items()}
        sliced_shapes = {}
        for tid, shp in self.full_leaf_shapes.items():
            dims = self.leaf_sliced.get(tid, {})
            for d in dims:
                if d < 0 or d >= len(shp) or shp[d] != 2:
                    raise SchemeError(f"leaf {tid}: sliced dim {d} invalid for shape {shp}")
            sliced_shapes[tid] = tuple(e for d, e in enumerate(shp) if d not in dims)
        parser = SchemeParser(sliced_shapes, sparse)
        self.steps: List[Step] = parser.parse(scheme)
        if not self.steps:
            raise SchemeError("empty scheme")
        self.leaf_info = parser.leaf_info
        self.result_info = self.steps[-1].c
        self.result_slot = self.steps[-1].i
        self.out_shape = tuple(([self.result_info.rows] if self.result_info.rows is not None else []) +
                               [2] * self.result_info.rank)
        self._lib = None
        self._handle = None
        self._build = build_native
        self._lower()

    # ------------------------------------------------------------------ lowering
    def _lower(self):
        arenas = {N.TNC_PHASE_ONCE: Arena(), N.TNC_PHASE_SLICE: Arena(), REGION_KEEP: Arena()}
        base = [0, 0, 0]
        reuse = self.slice_reuse = bool(self.options.slice_reuse and self.n_sliced > 0)
        # sliced bonds behind every step's result, and behind the step that consumes it (the accumulate runs for
        # every slice: it "depends" on every bond)
        every = frozenset(range(self.n_sliced))
        slot_deps = {tid: frozenset(self.leaf_sliced.get(tid, {}).values()) for tid in self.leaf_info}
        slot_producer: Dict[int, int] = {}
        self.step_deps: List[frozenset] = []
        consumer_deps: List[frozenset] = []
        for idx, st in enumerate(self.steps):
            d = slot_deps[st.i] | slot_deps[st.j]
            for t in (st.i, st.j):
                if t in slot_producer:
                    consumer_deps[slot_producer[t]] = d
            self.step_deps.append(d)
            consumer_deps.append(every)
            slot_deps[st.i] = d
            slot_producer[st.i] = idx
            del slot_deps[st.j]
            slot_producer.pop(st.j, None)
        # per step, before anything is allocated: operand roles, phase, algorithm
        pre = []
        self.step_phase, self.step_algo = [], []
        for idx, orig in enumerate(self.steps):
            st = orig
            is_swapped = bool(self.options.swap_operands and self.dtype == N.TNC_C64 and should_swap(orig))
            if is_swapped:
                st = swapped(orig)
            phase = N.TNC_PHASE_SLICE if (self.step_deps[idx] or not self.options.hoist) else N.TNC_PHASE_ONCE
            algo = N.TNC_ALGO_SIMT
            if self.dtype == N.TNC_C64:
                o = self.options
                if st.a.numel >= o.skinny_min_elems and skinny_eligible(st, o.tc_precision, o.skinny_min_n):
                    algo = N.TNC_ALGO_SKINNY
                elif st.flops >= o.tc_min_flops and st.flops >= o.tc_min_intensity * st.bytes_c64 and tc_eligible(st, o.tc_precision):
                    algo = N.TNC_ALGO_TC
                elif st.c.numel >= o.stem_min_elems and stem_eligible(st):
                    algo = N.TNC_ALGO_STEM
            pre.append((st, is_swapped, phase, algo))
            self.step_phase.append(phase)
            self.step_algo.append(algo)
        # slice_reuse: when each step runs (run_deps: the sliced bonds whose change makes it run -- its own, or its
        # reader's when it is tied to it) and which results must outlive a slice (kept)
        self.step_tied = [False] * len(self.steps)
        self.run_deps = list(self.step_deps)
        consumer_of = [None] * len(self.steps)
        prod = {}
        for idx, st in enumerate(self.steps):
            for t in (st.i, st.j):
                if t in prod:
                    consumer_of[prod[t]] = idx
            prod[st.i] = idx
            prod.pop(st.j, None)
        self._consumer_of = consumer_of
        kept = [False] * len(self.steps)
        if reuse:
            kept = self._plan_keep(consumer_of, every)
        # leaf blob layout (elements), in ascending tensor id order over the leaves the scheme uses
        self.leaf_order = sorted(self.leaf_info)
        self.leaf_src_offset = {}
        off = 0
        for tid in self.leaf_order:
            self.leaf_src_offset[tid] = off
            off += int(np.prod(self.full_leaf_shapes[tid], dtype=np.int64)) if self.full_leaf_shapes[tid] else 1
        self.leaf_blob_elems = off

        bufs: Dict[int, Buf] = {}
        leaf_bufs = {N.TNC_PHASE_ONCE: [], N.TNC_PHASE_SLICE: []}
        for tid in self.leaf_order:
            info = self.leaf_info[tid]
            region = N.TNC_PHASE_SLICE if (tid in self.leaf_sliced or not self.options.hoist) else N.TNC_PHASE_ONCE
            o, sz = arenas[region].alloc(info.numel * self.elem_bytes)
            buf = Buf(region, o, sz, info, logical_positions(info), True, base,
                      frozenset(self.leaf_sliced.get(tid, {}).values()))
            bufs[tid] = buf
            leaf_bufs[region].append((tid, buf))

        pending = []        # (phase, step, A, B, C, algo) in scheme order
        for idx, orig in enumerate(self.steps):
            A, B = bufs[orig.i], bufs[orig.j]
            st, is_swapped, phase, algo = pre[idx]
            if is_swapped:
                A, B = B, A
            assert phase == (N.TNC_PHASE_SLICE if (A.dependent or B.dependent) else N.TNC_PHASE_ONCE)
            cpos = self._choose_layout(st, A, B, algo)
            # slice_reuse: a result whose reader runs more often than it does is read again in later slices
            # without being recomputed -- it gets memory nothing else is ever laid over (KEEP); a result that
            # runs exactly when its reader does is recomputed whenever it is read and recycles as before
            region = REGION_KEEP if kept[idx] else phase
            o, sz = arenas[region].alloc(st.c.numel * self.elem_bytes)
            Cb = Buf(region, o, sz, st.c, cpos, False, base, self.run_deps[idx])
            scratch = None
            if algo == N.TNC_ALGO_TC:
                # packed operand panels live only while the step runs
                so, ssz = arenas[phase].alloc(tc_scratch_bytes(st))
                arenas[phase].release(so, ssz)
                scratch = (so, ssz)
            pending.append((phase, st, A, B, Cb, algo, scratch, self.step_tied[idx]))
            for old in (A, B):
                # a buffer dies with its consumer unless it is a leaf (reloaded / kept) or a
                # slice-invariant result consumed inside the slice loop (needed by every slice)
                if not old.is_leaf and old.region == phase:
                    arenas[phase].release(old.rel_offset, old.size)
            bufs[orig.i] = Cb
            del bufs[orig.j]
        base[N.TNC_PHASE_ONCE] = 0
        base[N.TNC_PHASE_SLICE] = arenas[N.TNC_PHASE_ONCE].high
        base[REGION_KEEP] = base[N.TNC_PHASE_SLICE] + arenas[N.TNC_PHASE_SLICE].high
        self.keep_bytes = arenas[REGION_KEEP].high
        # bytes [lo, hi) that are recycled within a slice (everything the SLICE arena holds except the leaves)
        leaf_hi = max([b.rel_offset + b.size for _, b in leaf_bufs[N.TNC_PHASE_SLICE]], default=0)
        self.recycled_range = (base[N.TNC_PHASE_SLICE] + leaf_hi, base[REGION_KEEP])
        self.arena_bytes = max(base[REGION_KEEP] + arenas[REGION_KEEP].high, ALIGN)
        # the library keeps its own words (amax words reduced by producing kernels, the slice-id word of graph
        # replay) in a tail behind the arena: tnc_plan_workspace_bytes
        self.workspace_bytes = (self.arena_bytes + ALIGN - 1) // ALIGN * ALIGN + N.TNC_WORKSPACE_TAIL_BYTES
        exec_flops = sum(st.flops for st, ph in zip(self.steps, self.step_phase) if ph == N.TNC_PHASE_SLICE)
        self.cuda_graph = (self.n_sliced >= 1 and exec_flops < self.options.graph_max_flops and not reuse
                           if self.options.cuda_graph is None else bool(self.options.cuda_graph))

        ops = {N.TNC_PHASE_ONCE: [], N.TNC_PHASE_SLICE: []}
        for phase in ops:
            if leaf_bufs[phase]:
                ops[phase].append(("leaves", [self._leaf_record(tid, b) for tid, b in leaf_bufs[phase]]))
        self.tables: List[np.ndarray] = []
        op_steps = {ph: [None] * len(ops[ph]) for ph in ops}     # Step behind each op (None: not an einsum)
        for phase, st, A, B, Cb, algo, scratch, tied in pending:
            rec = self._einsum_record(st, A, B, Cb, algo)
            if tied:
                rec.flags |= N.TNC_EINSUM_RUN_WITH_READER
            if scratch is not None:
                rec.scratch_offset = base[phase] + scratch[0]
                rec.scratch_bytes = scratch[1]
            ops[phase].append(("einsum", rec))
            op_steps[phase].append(st)
        final = bufs[self.result_slot]
        acc = N.TncAccum()
        acc.src = final.tensor()
        r = final.info.rank
        out_pos = [0] * r
        for i, m in enumerate(final.info.modes):
            out_pos[final.pos[m]] = r - 1 - i
        acc.out_pos = N.bits(out_pos)
        # the accumulate runs per slice even when nothing is sliced (one "slice")
        ops[N.TNC_PHASE_SLICE].append(("accum", acc))
        op_steps[N.TNC_PHASE_SLICE].append(None)
        self.ops = ops
        self.op_steps = op_steps
        if self._build:
            self._build_native(ops)

    def _plan_keep(self, consumer_of, every):
        """slice_reuse: which step results must outlive a slice.  A step runs when a bond of its run_deps changed:
        its own dependencies, or -- when it is TIED to its reader -- its reader's.  A result is KEPT (a region of its
        own) when its reader runs on other occasions than it does (over consecutive slice ids: when the lowest
        slice-id bits behind the two differ).  Without a budget nothing is tied.  With
        `keep_budget_bytes`, while the kept results exceed it, the kept result whose tie costs the least modelled
        time per byte is tied to its reader, together with the producers that ran in step with it (they keep
        running in step with it, so they stay un-kept)."""
        n = len(self.steps)
        slice_phase = [ph == N.TNC_PHASE_SLICE for ph in self.step_phase]
        size = [max(ALIGN, (st.c.numel * self.elem_bytes + ALIGN - 1) // ALIGN * ALIGN) for st in self.steps]   # as Arena.alloc rounds
        producers = [[] for _ in range(n)]          # step-produced inputs of every step
        for p, c in enumerate(consumer_of):
            if c is not None:
                producers[c].append(p)
        cost = self.step_seconds_model()
        bit = {b: self.n_sliced - 1 - b for b in range(self.n_sliced)}      # bond index -> slice-id bit

        def run_deps(tied):
            R = [None] * n
            for idx in range(n - 1, -1, -1):
                c = consumer_of[idx]
                R[idx] = (R[c] if c is not None else every) if tied[idx] else self.step_deps[idx]
            return R

        def low(deps):
            # going from slice id s - 1 to s flips the bits 0 .. ctz(s): a step runs exactly when its LOWEST bit is
            # among them, so two steps run on the same slices iff their lowest bits agree (None: only in the first slice)
            return min((bit[b] for b in deps), default=None)

        def kept_of(R):
            return [slice_phase[i] and low(R[i]) != low(R[consumer_of[i]] if consumer_of[i] is not None else every)
                    for i in range(n)]

        def amortised(R):
            return sum(cost[i] * 2.0 ** -min(bit[b] for b in R[i]) for i in range(n) if slice_phase[i] and R[i])

        tied = [False] * n
        R = run_deps(tied)
        kept = kept_of(R)
        budget = self.options.keep_budget_bytes
        while budget is not None and sum(size[i] for i in range(n) if kept[i]) > budget:
            base_cost = amortised(R)
            best = None
            for c in range(n):
                if not kept[c]:
                    continue
                trial = list(tied)
                stack = [c]
                while stack:                          # c and the producers that ran in step with it
                    x = stack.pop()
                    trial[x] = True
                    stack.extend(p for p in producers[x] if slice_phase[p] and not kept[p] and not trial[p])
                R2 = run_deps(trial)
                k2 = kept_of(R2)
                freed = sum(size[i] for i in range(n) if kept[i]) - sum(size[i] for i in range(n) if k2[i])
                if freed <= 0:
                    continue
                score = (amortised(R2) - base_cost) / freed
                if best is None or score < best[0]:
                    best = (score, trial, R2, k2)
            if best is None:
                break
            _, tied, R, kept = best
        self.step_tied, self.run_deps = tied, R
        return kept

    def _leaf_record(self, tid, buf: Buf):
        shp = self.full_leaf_shapes[tid]
        info = buf.info
        has_row = info.rows is not None
        nbits = len(shp) - (1 if has_row else 0)
        sliced = self.leaf_sliced.get(tid, {})
        rec = N.TncLeaf()
        rec.src_offset = self.leaf_src_offset[tid]
        rec.dst = buf.tensor()
        rec.src_rank = nbits
        rec.n_sliced = len(sliced)
        if len(sliced) > N.TNC_MAX_SLICED:
            raise SchemeError(f"leaf {tid} carries {len(sliced)} sliced bonds (max {N.TNC_MAX_SLICED})")
        sp, sb = [], []
        keep_dims = []
        for d in range(len(shp)):
            if has_row and d == 0:
                if d in sliced:
                    raise SchemeError(f"leaf {tid}: the row mode cannot be sliced")
                continue
            bitdim = d - (1 if has_row else 0)
            srcpos = nbits - 1 - bitdim
            if d in sliced:
                sp.append(srcpos)
                sb.append(sliced[d])
            else:
                keep_dims.append(srcpos)
        rec.sliced_pos = N.Sliced(*sp) if sp else N.Sliced()
        rec.sliced_bond = N.Sliced(*sb) if sb else N.Sliced()
        # kept dims in logical order are info.modes; destination position comes from buf.pos
        keep_pos = [0] * info.rank
        for mode, srcpos in zip(info.modes, keep_dims):
            keep_pos[buf.pos[mode]] = srcpos
        rec.keep_pos = N.bits(keep_pos)
        return rec

    def _choose_layout(self, st: Step, A: Buf, B: Buf, algo):
        """Physical mode order of the step's output.  The generic kernel writes the reference's
        logical order.  The tensor-core GEMM writes C[rows][m][n] with the right-only modes in
        the low positions; inside each group the modes keep the order they have in their
        operand, which keeps the pack kernels' reads in long contiguous runs."""
        if algo == N.TNC_ALGO_SIMT:
            return logical_positions(st.c)
        pos = {}
        for i, m in enumerate(sorted(st.n_modes, key=lambda m: B.pos[m])):
            pos[m] = i
        nn = len(st.n_modes)
        for i, m in enumerate(sorted(st.m_modes, key=lambda m: A.pos[m])):
            pos[m] = nn + i
        return pos


    def _table(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.int32)
        self.tables.append(arr)
        return len(self.tables) - 1

    def _rows_mode(self, r, nb):
        if r is None:
            return N.TNC_ROWS_NONE
        if len(r) == nb and np.array_equal(r, np.arange(nb)):
            return N.TNC_ROWS_IDENTITY
        return self._table(r)

    def _einsum_record(self, st: Step, A: Buf, B: Buf, Cb: Buf, algo):
        e = N.TncEinsum()
        e.a, e.b, e.c = A.tensor(), B.tensor(), Cb.tensor()
        e.nb = st.nb
        e.rows_a = self._rows_mode(st.ra, st.nb)
        e.rows_b = self._rows_mode(st.rb, st.nb)
        e.n_m, e.n_n, e.n_k, e.n_h = len(st.m_modes), len(st.n_modes), len(st.k_modes), len(st.h_modes)
        e.m_a = N.bits(A.pos[m] for m in st.m_modes)
        e.m_c = N.bits(Cb.pos[m] for m in st.m_modes)
        e.n_b = N.bits(B.pos[m] for m in st.n_modes)
        e.n_c = N.bits(Cb.pos[m] for m in st.n_modes)
        korder = sorted(range(len(st.k_modes)), key=lambda i: A.pos[st.k_modes[i]])
        e.k_a = N.bits(A.pos[st.k_modes[i]] for i in korder)
        e.k_b = N.bits(B.pos[st.k_modes_b[i]] for i in korder)
        e.h_a = N.bits(A.pos[m] for m in st.h_modes)
        e.h_b = N.bits(B.pos[m] for m in st.h_modes_b)
        e.h_c = N.bits(Cb.pos[m] for m in st.h_modes)
        e.algo = algo
        e.flags = N.TNC_EINSUM_OUTER_ROWS if full_outer(st) else (N.TNC_EINSUM_OUTER_PAIRS if outer_pairs(st) else 0)
        return e

    def _build_native(self, ops):
        lib = N.load()
        handle = C.c_void_p()
        N.check(lib.tnc_plan_create(self.dtype, self.n_sliced, C.byref(handle)))
        self._lib, self._handle = lib, handle
        try:
            N.check(lib.tnc_plan_set_option(handle, N.TNC_OPT_TC_PRECISION, N.TC_PRECISIONS[self.options.tc_precision]))
            N.check(lib.tnc_plan_set_option(handle, N.TNC_OPT_CUDA_GRAPH, 1 if self.cuda_graph else 0))
            N.check(lib.tnc_plan_set_option(handle, N.TNC_OPT_FUSE_AMAX, 1 if self.options.fuse_amax else 0))
            N.check(lib.tnc_plan_set_option(handle, N.TNC_OPT_SLICE_REUSE, 1 if self.slice_reuse else 0))
            for t in self.tables:
                tid = C.c_int32()
                N.check(lib.tnc_plan_add_table(handle, t.ctypes.data_as(C.POINTER(C.c_int32)), len(t), C.byref(tid)))
            for phase in (N.TNC_PHASE_ONCE, N.TNC_PHASE_SLICE):
                for kind, rec in ops[phase]:
                    if kind == "leaves":
                        arr = (N.TncLeaf * len(rec))(*rec)
                        N.check(lib.tnc_plan_add_leaves(handle, phase, arr, len(rec)))
                    elif kind == "einsum":
                        N.check(lib.tnc_plan_add_einsum(handle, phase, C.byref(rec)))
                    elif kind == "permute":
                        N.check(lib.tnc_plan_add_permute(handle, phase, C.byref(rec)))
                    elif kind == "accum":
                        N.check(lib.tnc_plan_add_accum(handle, phase, C.byref(rec)))
            N.check(lib.tnc_plan_finalize(handle, self.arena_bytes))
            assert lib.tnc_plan_workspace_bytes(handle) == self.workspace_bytes
        except Exception:
            lib.tnc_plan_destroy(handle)
            self._handle = None
            raise

    def __del__(self):
        if getattr(self, "_handle", None) is not None and self._lib is not None:
            self._lib.tnc_plan_destroy(self._handle)
            self._handle = None

    # ------------------------------------------------------------------ host-side helpers
    @property
    def n_slices(self):
        return 1 << self.n_sliced

    def pack_leaves(self, leaves, device=None):
        """Flatten the leaves the scheme uses into one contiguous complex64 array (the leaf blob).
        `leaves` is a list or dict of torch tensors as the reference takes them
        (simulation.py:92-99 list, :168-171 dict).  Host leaves are gathered into one pinned
        staging buffer and cross to `device` in a single copy; the reference moves every leaf
        with its own `.to(device)` (simulation.py:93-99)."""
        import torch
        parts = []
        for tid in self.leaf_order:
            t = leaves[tid]
            if tuple(t.shape) != self.full_leaf_shapes[tid]:
                raise SchemeError(f"leaf {tid}: shape {tuple(t.shape)} differs from the planned {self.full_leaf_shapes[tid]}")
            parts.append(t.reshape(-1) if t.dtype == torch.complex64 else t.reshape(-1).to(torch.complex64))
        on_host = all(p.device.type == "cpu" for p in parts)
        if device is None or not on_host:
            blob = torch.cat(parts) if len(parts) > 1 else parts[0].clone()
            blob = blob.to(torch.complex64)
            return blob if device is None else blob.to(device)
        # pinned staging buffers: a ring of two, each guarded by the CUDA event recorded behind its
        # last host->device copy -- the copy is asynchronous, so the buffer must not be refilled
        # while an earlier call's copy may still be queued behind a long execute
        pin = torch.device(device).type == "cuda" and torch.cuda.is_available()
        ring = getattr(self, "_stages", None)
        if ring is None:
            ring = self._stages = {"next": 0, "slots": [None, None]}
        i = ring["next"]
        ring["next"] = (i + 1) % len(ring["slots"])
        slot = ring["slots"][i]
        if slot is None or slot[0].numel() != self.leaf_blob_elems:
            slot = ring["slots"][i] = [torch.empty(self.leaf_blob_elems, dtype=torch.complex64, pin_memory=pin), None]
        stage, event = slot
        if event is not None:
            event.synchronize()
        torch.cat(parts, out=stage)
        blob = stage.to(device, non_blocking=True)
        if pin:
            slot[1] = torch.cuda.Event()
            slot[1].record(torch.cuda.current_stream(device))
        return blob

    def execute(self, leaf_blob, out, slice_begin, slice_end, workspace, stream_ptr):
        """Enqueue the contraction of slices [slice_begin, slice_end) on the given stream;
        `out` (complex64, logical result order) is accumulated into."""
        rc = self._lib.tnc_plan_execute(self._handle, leaf_blob.data_ptr(), int(slice_begin), int(slice_end),
                                        out.data_ptr(), workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                        stream_ptr)
        N.check(rc)

    def profile(self, leaf_blob, out, slice_id, workspace, stream_ptr):
        """Per-operation device times (ms) of the ONCE phase and of one slice: two flat lists with
        TNC_PROFILE_SLOTS entries per element of self.ops[phase] (slot 0 = whole operation, then
        its launches).  Measurement aid for bench.py; synchronises."""
        n0, n1 = len(self.ops[N.TNC_PHASE_ONCE]), len(self.ops[N.TNC_PHASE_SLICE])
        SL = N.TNC_PROFILE_SLOTS
        a0, a1 = (C.c_float * max(n0 * SL, 1))(), (C.c_float * max(n1 * SL, 1))()
        N.check(self._lib.tnc_plan_profile(self._handle, leaf_blob.data_ptr(), int(slice_id), out.data_ptr(),
                                           workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                           stream_ptr, a0, a1))
        return list(a0)[:n0 * SL], list(a1)[:n1 * SL]

    def fused_amax_operands(self, phase=None):
        """Tensor-core operands (of `phase`; default: of both phases) whose amax the producing kernel reduces
        instead of a separate pass over the tensor."""
        phases = (N.TNC_PHASE_ONCE, N.TNC_PHASE_SLICE) if phase is None else (phase,)
        return sum(int(self._lib.tnc_plan_num_fused_amax(self._handle, ph)) for ph in phases)

    @property
    def last_launches(self):
        return int(self._lib.tnc_plan_last_launches(self._handle))

    # ------------------------------------------------------------------ cross-slice reuse
    def step_seconds_model(self):
        """Rough B200 time of every step (seconds): tensor-bound GEMM steps at the measured useful rate of the
        split-precision product, everything else at the streaming kernels' in-slice HBM rate, plus a launch.
        Only the RATIOS matter: it ranks sliced bonds for `reuse_bond_order` and prices `reuse_summary`."""
        out = []
        for st, algo in zip(self.steps, self.step_algo):
            rate = 5.5e14 if self.options.tc_precision != "f16" else 1.3e15
            t = max(st.flops / rate, st.bytes_c64 / 4.5e12) if algo == N.TNC_ALGO_TC else st.bytes_c64 / 4.0e12
            if algo == N.TNC_ALGO_TC:
                t += (st.a.numel + st.b.numel) * 20 / 5.0e12          # operand panels: 8 bytes read, 12 written
            out.append(t + 3e-6)
        return out

    def reuse_bond_order(self, fixed_high=0):
        """Order of the sliced bonds (most significant slice-id bit first, as `slicing_bonds` is read) that makes
        consecutive slice ids share the most work under PlanOptions.slice_reuse: over a long range a step runs once
        per 2^p slices, p = the lowest slice-id bit among the bonds it depends on, so the bonds that the expensive
        steps do NOT depend on should be the fastest-changing bits.  Greedy from the least significant bit up: give
        the next bit to the bond that the cheapest set of not-yet-charged steps depends on.  The first `fixed_high`
        bonds keep their (most significant) places."""
        cost = self.step_seconds_model()
        live = [(c, d) for c, d in zip(cost, self.step_deps) if d]
        remaining = set(range(fixed_high, self.n_sliced))
        low_first = []
        while remaining:
            best = min(sorted(remaining), key=lambda b: sum(c for c, d in live if b in d))
            low_first.append(best)
            remaining.discard(best)
            live = [(c, d) for c, d in live if best not in d]
        return list(range(fixed_high)) + low_first[::-1]

    def reuse_summary(self, bond_order=None):
        """Modelled seconds per slice: every step for every slice, against the amortised cost over a long range of
        consecutive slice ids with slice_reuse (a step runs once per 2^(lowest bit it depends on) slices).
        `bond_order`: a permutation of range(n_sliced), most significant first (default: the plan's own)."""
        cost = self.step_seconds_model()
        order = list(range(self.n_sliced)) if bond_order is None else list(bond_order)
        bit = {b: self.n_sliced - 1 - i for i, b in enumerate(order)}
        full = sum(c for c, ph in zip(cost, self.step_phase) if ph == N.TNC_PHASE_SLICE)
        # the plan's own order: what really runs (ties of keep_budget_bytes included); another order: the dependencies
        deps = self.run_deps if bond_order is None else self.step_deps
        amortised = sum(c * 2.0 ** -min(bit[b] for b in d) for c, d in zip(cost, deps) if d)
        return {"full_s": full, "amortised_s": amortised}

    def work_summary(self):
        """Executed vs reference-equivalent work per slice."""
        tot_f = tot_b = dep_f = dep_b = 0
        for st, ph in zip(self.steps, self.step_phase):
            tot_f += st.flops
            tot_b += st.bytes_c64
            if ph == N.TNC_PHASE_SLICE:
                dep_f += st.flops
                dep_b += st.bytes_c64
        return {"steps": len(self.steps), "ref_flops_per_slice": tot_f, "ref_bytes_per_slice": tot_b,
                "exec_flops_per_slice": dep_f, "exec_bytes_per_slice": dep_b,
                "hoisted_steps": sum(1 for p in self.step_phase if p == N.TNC_PHASE_ONCE)}
