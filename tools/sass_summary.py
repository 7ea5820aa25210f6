#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native code path
(B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP
(TMA), UTCBAR (tcgen05.commit), SYNCS (mbarrier), ELECT, packed fp32x2 math, and the legacy
HMMA / warp-uniformisation loops that must be absent.  Runs where cuobjdump is (no GPU needed):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "artensor_b200", "libtnc_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "ELECT",
        "FADD2", "FFMA2", "FMUL2", "FFMA", "HMMA", "BRA.U.ANY", "LDG", "STG", "LDS", "STS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts, total, cur = collections.OrderedDict(), collections.Counter(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not (m and cur):
            continue
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA.2CTA")):
                counts[cur][k] += 1
                total[k] += 1
    print(f"SASS summary of {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes), sm_100a\n")
    print("whole library: " + ", ".join(f"{k} {total[k]}" for k in KEYS if total[k]))
    print("HMMA (legacy mma.sync path): %d;  BRA.U.ANY (warp-uniformisation loops around tcgen05/TMA issue): %d\n"
          % (total["HMMA"], total["BRA.U.ANY"]))
    for name, c in counts.items():
        if not any(c[k] for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "FADD2", "FFMA2")):
            continue
        short = demangle(name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("tnc::", "")
        short = re.sub(r"^void ", "", short)
        short = short[:short.index(">(") + 1] if ">(" in short else short.split("(")[0]
        print(f"{short}\n    " + ", ".join(f"{k} {c[k]}" for k in ["instructions"] + KEYS if c[k]))


if __name__ == "__main__":
    sys.exit(main())
