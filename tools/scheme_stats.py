#!/usr/bin/env python
"""Print the per-step shape table of a case (rows, m/n/k bits, flops, bytes)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from artensor_b200.cases import load_case
from artensor_b200.plan import SchemeParser

def main(path, min_log2=16):
    case = load_case(path)
    sidx = case.slicing_indices()
    shapes = {}
    nsl = {}
    for b, lst in sidx.items():
        for tid, d in lst:
            nsl[tid] = nsl.get(tid, 0) + 1
    for tid, t in case.leaves.items():
        shp = list(t.shape)
        for _ in range(nsl.get(tid, 0)):
            shp.remove(2) if False else None
        # sliced dims removed
        dims = sorted([d for b, lst in sidx.items() for (tt, d) in lst if tt == tid], reverse=True)
        for d in dims:
            shp.pop(d)
        shapes[tid] = tuple(shp)
    steps = SchemeParser(shapes, case.pattern == 'sparse').parse(case.scheme)
    sliced = set(nsl)
    dep = {}
    tot_f = tot_b = 0
    dep_f = dep_b = 0
    print(f"{case.name}: {len(steps)} steps, {len(case.slicing_bonds)} sliced bonds")
    for st in steps:
        d = dep.get(st.i, st.i in sliced) or dep.get(st.j, st.j in sliced)
        dep[st.i] = d
        tot_f += st.flops; tot_b += st.bytes_c64
        if d: dep_f += st.flops; dep_b += st.bytes_c64
        sz = max(st.a.numel, st.b.numel, st.c.numel)
        if sz >= (1 << min_log2):
            print(f"{st.index:4d} {st.kind:7s} dep={int(d)} i={st.i:4d} j={st.j:4d} rowsA={st.a.rows} rowsB={st.b.rows} rowsC={st.c.rows} "
                  f"rA={st.a.rank:2d} rB={st.b.rank:2d} rC={st.c.rank:2d} m={len(st.m_modes):2d} n={len(st.n_modes):2d} k={len(st.k_modes):2d} "
                  f"h={len(st.h_modes)} chunks={st.chunks} flops={st.flops:.3e} bytes={st.bytes_c64:.3e} AI={st.flops/st.bytes_c64:.1f}")
    print(f"total flops {tot_f:.3e} bytes {tot_b:.3e}; slice-dependent flops {dep_f:.3e} bytes {dep_b:.3e}")
    return steps

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
