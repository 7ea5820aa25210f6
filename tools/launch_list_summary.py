#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file launches.csv):
time and launch count per kernel of this library, shares of the total.

    python tools/launch_list_summary.py gpurun_out/launches.csv > profiles/rNN_ncu_launch_list.txt"""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ours, other = defaultdict(lambda: [0.0, 0]), defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        if len(r) <= val_i:
            continue
        v = float(r[val_i].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[unit_i].replace("nsecond", "ns").replace("usecond", "us").replace("msecond", "ms").replace("second", "s"), 1e-6)
        name = re.sub(r"\(.*$", "", r[name_i]).strip()
        tgt = ours if "tnc::" in name else other
        tgt[name][0] += ms
        tgt[name][1] += 1
    total = sum(v[0] for v in ours.values())
    for name, (ms, n) in sorted(ours.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:10.3f} ms {100 * ms / total:5.1f}% n={n:5d} {name}")
    print(f"total {total:.3f} ms over {sum(v[1] for v in ours.values())} launches")
    for name, (ms, n) in sorted(other.items(), key=lambda kv: -kv[1][0]):
        print(f"   (not ours) {ms:10.3f} ms n={n:5d} {name[:100]}")


if __name__ == "__main__":
    main()
