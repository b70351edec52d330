#!/usr/bin/env python
"""Exact mode (-e) through the `miekki` command line at scale (BASELINE config 5 in part): N genome
FASTA files of 5 Mbp in /dev/shm, R error-free 1 kbp reads, `miekki -l list -a reads -e -h 20`.
The second `elapsed time:` line is reads -> exact-mode lines (approximate query of every read,
then the true k-mer intersection with every genome that got a candidate).  Prints one JSON line;
the files are generated on the GPU and written by this script (not timed)."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miekki_b200  # noqa: E402
from miekki_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=2000)
ap.add_argument("--reads", type=int, default=20000)
ap.add_argument("--genome-len", type=int, default=5_000_000)
ap.add_argument("--gpus", type=int, default=1)
a = ap.parse_args()
CLI = os.path.join(ROOT, "miekki_b200", "cli", "miekki")
SEED = 0x5EED_B200
nproc = os.cpu_count() or 8


def main():
    d = tempfile.mkdtemp(prefix="miekki_exact_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        ix = miekki_b200.Miekki(k=31, h=10)
        names = []
        for g0 in range(0, a.genomes, 64):
            m = min(64, a.genomes - g0)
            b = ix.synth(SEED, g0, m, a.genome_len)
            for i in range(m):
                p = os.path.join(d, "g%d.fa" % (g0 + i))
                with open(p, "wb") as f:
                    f.write(b">genome%d\n" % (g0 + i) + b.download(i, a.genome_len) + b"\n")
                names.append(p)
            b.free()
        ix.close()
        with open(os.path.join(d, "list.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        with open(os.path.join(d, "reads.fa"), "wb") as f:
            for b0 in range(0, a.reads, 10_000):
                m = min(10_000, a.reads - b0)
                r, gs, ps = synth.cb_reads(SEED, a.genomes, a.genome_len, m, 1000, 0.0, block=b0 // 10_000)
                for i in range(m):
                    f.write(b">read%d_g%d\n" % (b0 + i, int(gs[i])) + r[i].tobytes() + b"\n")
        env = dict(os.environ, MIEKKI_TIMING="1")
        out = {}
        for tag in ("first run", "second run"):
            t0 = time.perf_counter()
            r = subprocess.run([CLI, "-l", "list.txt", "-a", "reads.fa", "-e", "-h", "20", "-t", str(nproc), "-o", "exact.txt",
                                "--gpus", str(a.gpus)], cwd=d, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                sys.exit("miekki failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
            el = [float(x) for x in re.findall(r"elapsed time: ([0-9.eE+-]+)s", r.stdout)]
            lines = sum(1 for _ in open(os.path.join(d, "exact.txt")))
            out[tag] = {"build_s": el[0], "query_and_exact_s": el[1], "process_wall_s": time.perf_counter() - t0,
                        "exact_lines": lines}
        print(json.dumps({"genomes": a.genomes, "reads": a.reads, "host_cores": nproc, "gpus": a.gpus, **out}))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
