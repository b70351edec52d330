#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libmiekki_b200.so (cuobjdump -sass), written as a
markdown table: the evidence that the scan stages rows with the bulk-copy engine (UBLKCP) behind
mbarriers (SYNCS), that its arithmetic is 3-input logic (LOP3) and that the sketch kernel's bucket
minimum is a 64-bit reduction at L2 (RED / ATOMG).  Needs no GPU.

    python benchmarks/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "miekki_b200", "libmiekki_b200.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LOP3", "LDS", "LDG", "STG", "REDG", "ATOMG", "ATOMS", "IMAD", "SHF",
         "FLO", "POPC", "SHFL", "BAR", "LDL", "STL"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
        return out.split("\n")[:len(names)]
    except Exception:
        return names


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    names = demangle(list(kernels))
    print("# SASS opcode histogram per kernel (`cuobjdump -sass miekki_b200/libmiekki_b200.so`)\n")
    print("Architectures in the file: %s (no PTX fallback is embedded).  Counts are static instructions.\n"
          % ", ".join(sorted(arch)))
    print("| kernel | total | " + " | ".join(WATCH) + " |")
    print("|---|---:|" + "---:|" * len(WATCH))
    for (mangled, cnt), nice in zip(kernels.items(), names):
        short = nice.replace("(anonymous namespace)::", "")
        short = re.sub(r"^void ", "", re.sub(r"\((?!anonymous).*$", "", short))
        print("| `%s` | %d | " % (short, cnt["_total"]) + " | ".join(str(cnt.get(w, 0)) for w in WATCH) + " |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
