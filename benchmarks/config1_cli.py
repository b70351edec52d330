#!/usr/bin/env python
"""BASELINE.json config 1 through both command lines on the same box: N synthetic 5 Mbp genomes
(one FASTA file each, SURVEY.md 8d generator), R synthetic 10 kbp reads, `-k 31 -h 17 -s 0`,
all host threads.  The two `elapsed time:` lines each binary prints (main.cpp:211,234) are
file -> index seconds and reads -> hit lines seconds.  Prints one JSON line.

The reference's genome ids depend on thread timing when -t > 1, so hit
lines are compared as multisets only when --check is given (reference rerun at -t 1)."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miekki_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=100)
ap.add_argument("--reads", type=int, default=1000)
ap.add_argument("--read-len", type=int, default=10_000)
ap.add_argument("--genome-len", type=int, default=5_000_000)
ap.add_argument("--h", type=int, default=17)
ap.add_argument("--k", type=int, default=31)
ap.add_argument("--check", action="store_true")
ap.add_argument("--no-reference", action="store_true")
ap.add_argument("--repeat", type=int, default=2, help="runs of our CLI (the first one pages the files in)")
a = ap.parse_args()

CLI = os.path.join(ROOT, "miekki_b200", "cli", "miekki")
REF = os.path.join(ROOT, "oracle", "_ref", "Miekki")
nproc = os.cpu_count() or 8


def elapsed(out):
    return [float(x) for x in re.findall(r"elapsed time: ([0-9.eE+-]+)s", out)]


def run(binary, args, cwd):
    t0 = time.perf_counter()
    r = subprocess.run([binary] + [str(x) for x in args], cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit("%s failed:\n%s\n%s" % (binary, r.stdout[-2000:], r.stderr[-2000:]))
    for line in r.stderr.splitlines():
        if line.startswith("[timing]"):
            print(os.path.basename(binary), line, file=sys.stderr)
    return elapsed(r.stdout), time.perf_counter() - t0


with tempfile.TemporaryDirectory() as d:
    names = []
    cache = {}
    for g in range(a.genomes):
        s = synth.genome(g, a.genome_len)
        p = os.path.join(d, "genome%d.fa" % g)
        synth.write_fasta(p, ">genome%d" % g, s)
        names.append(p)
        if g < 16:
            cache[g] = s
    with open(os.path.join(d, "genomes.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    rng = np.random.default_rng(2_000_000)
    src = rng.integers(min(a.genomes, 16), size=a.reads)
    pos = rng.integers(a.genome_len - a.read_len, size=a.reads)
    with open(os.path.join(d, "reads.fa"), "wb") as f:
        for r in range(a.reads):
            g, p = int(src[r]), int(pos[r])
            f.write(b">read%d_g%d_p%d\n" % (r, g, p) + cache[g][p:p + a.read_len] + b"\n")
    common = ["-l", "genomes.txt", "-a", "reads.fa", "-k", a.k, "-h", a.h, "-s", 0]
    gbp = a.genomes * a.genome_len / 1e9
    kbp = a.reads * a.read_len / 1e3
    rec = {"workload": "C1: %d x %.1f Mbp FASTA files, -k %d -h %d, %d reads of %d bp, -s 0" %
                       (a.genomes, a.genome_len / 1e6, a.k, a.h, a.reads, a.read_len), "host_threads": nproc}
    ours = []
    for _ in range(max(1, a.repeat)):
        t, wall = run(CLI, common + ["-t", nproc, "-o", "gpu_hits.txt"], d)
        ours.append({"build_s": t[0], "query_s": t[1], "wall_s": wall, "build_gbp_per_s": gbp / t[0],
                     "query_kbp_per_s": kbp / t[1]})
    rec["miekki_b200"] = ours
    if not a.no_reference:
        t, wall = run(REF, common + ["-t", nproc, "-o", "ref_hits.txt"], d)
        rec["reference"] = {"build_s": t[0], "query_s": t[1], "wall_s": wall, "build_gbp_per_s": gbp / t[0],
                            "query_kbp_per_s": kbp / t[1]}
        if a.check:
            run(REF, common + ["-t", 1, "-o", "ref_hits_t1.txt"], d)
            x = open(os.path.join(d, "ref_hits_t1.txt")).read()
            y = open(os.path.join(d, "gpu_hits.txt")).read()
            rec["hit_lines_identical"] = x == y
            rec["hit_lines_multiset_equal"] = Counter(x.split("\n")) == Counter(y.split("\n"))
    print(json.dumps(rec))
