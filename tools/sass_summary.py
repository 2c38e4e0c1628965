#!/usr/bin/env python
"""Per-kernel SASS opcode summary of librr_b200.so (runs without a GPU): the tensor-core / TMA /
bulk-copy mnemonics that prove which hardware paths each kernel uses.

    python tools/sass_summary.py > profiles/r2_sass_opcodes.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "radiant-rag_b200" / "librr_b200.so"
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "ATOMS", "ATOMG",
         "RED", "POPC", "LOP3", "DFMA", "DADD", "IDP", "LDGSTS", "MATCH"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode summary of librr_b200.so (sm_100a)\n")
    print("`cuobjdump -sass radiant-rag_b200/librr_b200.so`, static instruction counts per kernel; columns are the "
          "mnemonics that identify the hardware path (UTCIMMA = tcgen05.mma kind::i8, UTMALDG = TMA tensor load, "
          "UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, SYNCS = mbarrier, ATOMS / ATOMG / RED = atomics).\n")
    cols = [c for c in WATCH if any(k[c] for k in kernels.values())]
    print("| kernel | total | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    tot = collections.Counter()
    for (name, cnt), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("rr::", "")
        print(f"| `{short}` | {sum(cnt.values())} | " + " | ".join(str(cnt[c]) if cnt[c] else "" for c in cols) + " |")
        tot.update(cnt)
    print("| **all kernels** | %d | " % sum(tot.values()) + " | ".join(str(tot[c]) for c in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
