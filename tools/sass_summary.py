"""Per-kernel counts of the SASS mnemonics that prove Blackwell-native code (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP, mbarrier -> SYNCS, ex2.approx -> MUFU.EX2) in the built libraries.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "UTCBAR", "MUFU.EX2", "FFMA2", "FADD2", "HMMA", "LDGSTS", "DFMA", "ATOM", "RED"]


def summarise(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(CUtensorMap_st.*", "(...)", name)[:150]
            per[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[name]["total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    per[name][p] += 1
    return per


if __name__ == "__main__":
    for lib in ("spadot_b200/libspadot_b200.so", "spadot_b200/libot_b200.so"):
        path = os.path.join(ROOT, lib)
        print(f"== {lib}  (cuobjdump -sass, sm_100a)")
        per = summarise(path)
        tot = collections.Counter()
        for name, c in per.items():
            tot.update(c)
            hot = {p: c[p] for p in PATTERNS if c[p]}
            print(f"{c['total']:6d} instr  {name}\n              {hot}")
        print("-- library totals:", {p: tot[p] for p in PATTERNS if tot[p]}, "\n")
