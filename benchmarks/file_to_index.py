#!/usr/bin/env python
"""File -> index through the `miekki` command line at scale: N genome FASTA
files of 5 Mbp in /dev/shm (plain, and a gzip-1 subset), `miekki -l list -t <all cores>`; the
first `elapsed time:` line is file -> index seconds.  MIEKKI_TIMING=1 gives the phases on stderr:
how long the host spent parsing a wave of files ("parse wave": open, inflate, getline, append)
and how long mk_index_add took for it (host packing + H2D + sketch), the two running overlapped.
The files are generated on the GPU (counter-based genomes, miekki_b200/synth.py) and written by
this script; that part is not timed.  Prints one JSON line."""
import argparse
import gzip
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

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=1000)
ap.add_argument("--gz-genomes", type=int, default=200)
ap.add_argument("--genome-len", type=int, default=5_000_000)
ap.add_argument("--h", type=int, default=20)
ap.add_argument("--line", type=int, default=0, help="FASTA line width (0: the whole sequence on one line)")
a = ap.parse_args()

CLI = os.path.join(ROOT, "miekki_b200", "cli", "miekki")
nproc = os.cpu_count() or 8
SEED = 0x5EED_B200


def run(list_file, cwd, n):
    env = dict(os.environ, MIEKKI_TIMING="1")
    t0 = time.perf_counter()
    r = subprocess.run([CLI, "-l", list_file, "-k", "31", "-h", str(a.h), "-t", str(nproc), "-o", "out.txt"], cwd=cwd,
                       capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        sys.exit("miekki failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    build_s = float(re.findall(r"elapsed time: ([0-9.eE+-]+)s", r.stdout)[0])
    phases = {}
    for what, s in re.findall(r"\[timing\] (.*?) ([0-9.eE+-]+) s", r.stderr):
        phases.setdefault(what, []).append(float(s))
    return {"genomes": n, "build_s": build_s, "process_wall_s": wall, "gbp_per_s": n * a.genome_len / build_s / 1e9,
            "phases_s": {k: {"calls": len(v), "sum": round(sum(v), 3), "max": round(max(v), 3)} for k, v in phases.items()}}


def main():
    d = tempfile.mkdtemp(prefix="miekki_f2i_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        ix = miekki_b200.Miekki(k=31, h=10)
        plain, gz = [], []
        t0 = time.perf_counter()
        for g0 in range(0, a.genomes, 64):
            m = min(64, a.genomes - g0)
            b = ix.synth(SEED, g0, m, a.genome_len)
            for i in range(m):
                seq = b.download(i, a.genome_len)
                p = os.path.join(d, "g%d.fa" % (g0 + i))
                body = seq if a.line <= 0 else b"\n".join(seq[j:j + a.line] for j in range(0, len(seq), a.line))
                with open(p, "wb") as f:
                    f.write(b">genome%d\n" % (g0 + i) + body + b"\n")
                plain.append(p)
                if g0 + i < a.gz_genomes:
                    pz = p + ".gz"
                    with gzip.open(pz, "wb", compresslevel=1) as f:
                        f.write(b">genome%d\n" % (g0 + i) + body + b"\n")
                    gz.append(pz)
            b.free()
        ix.close()
        gen_s = time.perf_counter() - t0
        out = {"host_cores": nproc, "h": a.h, "genome_len": a.genome_len, "fasta_line_width": a.line,
               "files_written_s": round(gen_s, 1)}
        with open(os.path.join(d, "plain.txt"), "w") as f:
            f.write("\n".join(plain) + "\n")
        run(os.path.join(d, "plain.txt"), d, len(plain))          # first run pages the files in and warms the driver
        out["plain"] = run(os.path.join(d, "plain.txt"), d, len(plain))
        if gz:
            with open(os.path.join(d, "gz.txt"), "w") as f:
                f.write("\n".join(gz) + "\n")
            out["gzip"] = run(os.path.join(d, "gz.txt"), d, len(gz))
        print(json.dumps(out))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
