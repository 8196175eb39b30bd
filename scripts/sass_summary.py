#!/usr/bin/env python
"""Counts of the Blackwell-native SASS mnemonics per kernel of the shipped library (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores, UBLKCP = bulk copies.

usage: sass_summary.py [lib.so] > profiles/<tag>_sass_summary.md
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "fragnet_b200/_lib/libfragnet_b200.so"
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            for p in PAT:
                if op == p:
                    counts[cur][p] += 1
    print("# SASS evidence: Blackwell-native instructions per kernel of `libfragnet_b200.so`\n")
    print("`cuobjdump -sass fragnet_b200/_lib/libfragnet_b200.so`, instruction counts per kernel (static code, not executed "
          "counts).  `UTCHMMA` = `tcgen05.mma` (kind::tf32 / f16), `LDTM` = `tcgen05.ld`, `UTMALDG` / `UTMASTG` = TMA tensor "
          "load / store (`cp.async.bulk.tensor`), `UBLKCP` = 1-D bulk copy, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier "
          "operations.  No `HMMA` (legacy `mma.sync`) anywhere: the dense path is tcgen05 only.\n")
    print("| kernel | " + " | ".join(PAT) + " |\n|---|" + "---:|" * len(PAT))
    for k, c in counts.items():
        if any(c[p] for p in PAT[:-1]):
            print(f"| `{k[:70]}` | " + " | ".join(str(c[p]) if c[p] else "" for p in PAT) + " |")
    n_tc = sum(1 for c in counts.values() if c["UTCHMMA"])
    print(f"\n{len(counts)} kernels in the library; {n_tc} issue tcgen05 MMAs; "
          f"{sum(1 for c in counts.values() if c['HMMA'])} use legacy HMMA.")


if __name__ == "__main__":
    main()
