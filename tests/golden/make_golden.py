#!/usr/bin/env python
"""Generates the golden vectors in this directory by running the UNMODIFIED
reference binary (oracle/_ref/Miekki, built by `make -C oracle ref`).

Run from the repo root, in the container that has /root/reference:

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own; these
files are what pins the oracle (oracle/miekki_oracle.c) and, through it and
directly, the CUDA path.  All reference runs use `-t 1` so genome ids are list
order and Bloom byte values are deterministic.

Cases
  caseA  14 genomes x 100 kbp (8 are relatives of two others), -k 31 -h 12; inputs
         committed (gz FASTA, one
         multi-record genome with a short record, one with N / lowercase runs).
         Outputs: full dump contents, hit lines at -s 200 / -s 0 / -s 5000,
         exact-mode lines.
  caseB  same genomes, -k 21 -h 10 -b 34 (other k, h and Bloom geometry).
  caseD  70 genomes x 40 kbp (5 founders, 65 relatives at 0.3 % .. 3.5 %), -k 31 -h 12: ids cross
         the 32-genome groups of a row and the 64-genome build chunk; lists hold more than ten
         candidates spread over all of them.  Inputs regenerated from seeds (sha256 recorded).
  caseC  14 related genomes x 1.5 Mbp at -h 16: every sketch saturated -> genome_size 0
         (quirk G2) -> empty hit lists at -s 200 and the all-ties heap order at
         -s 0.  Inputs are regenerated from seeds (sha256 recorded); the dump is
         kept as hashes plus sampled rows.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from miekki_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = orc.RefBinary()


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_ref(cwd, args):
    out = REF.run(args, cwd=cwd, timeout=1200)
    return out


def dump_to_npz(dump_path, npz_path):
    d = orc.parse_dump(dump_path)
    nz = np.flatnonzero(d.bloom)
    np.savez_compressed(
        npz_path, k=d.k, h=d.h, nbm=d.nbm, nbmant=d.nbmant, n=d.n, b=d.b, bloom_bits=d.bloom_bits,
        threshold=d.threshold, compressed=d.compressed, rows=d.rows, genome_size=d.genome_size,
        sketch_size=d.sketch_size, bloom_idx=nz.astype(np.uint32), bloom_val=d.bloom[nz])
    return d


def mutate_special(seq: bytes, rng) -> bytes:
    """N runs, lowercase runs, stray IUPAC letters (quirk G8)."""
    a = bytearray(seq)
    for _ in range(6):
        p = int(rng.integers(0, len(a) - 200))
        a[p:p + 40] = b"N" * 40
    for _ in range(6):
        p = int(rng.integers(0, len(a) - 400))
        a[p:p + 150] = bytes(a[p:p + 150]).lower()
    for _ in range(20):
        a[int(rng.integers(len(a)))] = ord("R")
    # also inside the first k-1 characters of a later test read, see reads below
    return bytes(a)


def make_case_a_inputs(d):
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(77)
    L = 100_000
    genomes = [synth.genome(g, L) for g in range(6)]
    genomes[5] = mutate_special(genomes[5], rng)
    # 8 relatives of genomes 0 and 1 (0.5 % .. 8 % substitutions): more than ten
    # genomes clear the thresholds for reads cut from 0/1, so the bounded heap
    # replaces and ties (equal match counts) occur (quirk G5)
    for j, rate in enumerate((0.005, 0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.08)):
        src = np.frombuffer(genomes[j % 2], np.uint8)
        genomes.append(synth.substitute(src, rate, rng).tobytes())
    names = []
    for g, s in enumerate(genomes):
        name = "gA%d.fa.gz" % g
        with gzip.GzipFile(os.path.join(d, name), "wb", mtime=0) as f:
            if g == 4:
                # three records; the middle one is shorter than k (quirk G16 in
                # exact mode; quirk G9: records are concatenated when sketching)
                f.write(b">gA4 rec1\n" + s[:40_000] + b"\n>gA4 short\n" + s[40_000:40_020] +
                        b"\n>gA4 rec3\n")
                rest = s[40_020:]
                for i in range(0, len(rest), 70):      # wrapped lines
                    f.write(rest[i:i + 70] + b"\n")
            else:
                f.write(b">gA%d\n" % g + s + b"\n")
        names.append(name)
    with open(os.path.join(d, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    # reads
    reads = synth.sample_reads(genomes, 24, 2000, sub_rate=0.02, block=0)
    reads += synth.sample_reads(genomes[:2], 16, 3000, sub_rate=0.01, block=2)
    reads += synth.sample_reads(genomes, 8, 1000, sub_rate=0.0, block=1)
    g5 = genomes[5]
    reads.append((">special_window_g5", g5[1000:6000]))
    reads.append((">lower_whole", genomes[1][5000:7000].lower()))
    reads.append((">prefix_has_N", b"ACGTN" + genomes[2][7005:9000]))
    reads.append((">prefix_lower", genomes[2][9000:9020].lower() + genomes[2][9020:11000]))
    reads.append((">exactly_k", genomes[0][100:131]))
    reads.append((">k_plus_1", genomes[0][100:132]))
    reads.append((">shorter_than_k", genomes[0][100:120]))          # skipped by the reference
    reads.append((">random_unrelated", synth.genome(999, 3000)))
    reads.append((">starts_with_lower", b"a" + genomes[3][501:2500]))  # -e skips it (Miekki.cpp:736)
    synth.write_reads(os.path.join(d, "reads.fa"), reads)
    return genomes, reads


def case_a():
    d = os.path.join(HERE, "caseA")
    make_case_a_inputs(d)
    with tempfile.TemporaryDirectory() as tmp:
        for s in (200, 0, 5000):
            run_ref(d, ["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-t", 1, "-s", s,
                        "-o", os.path.join(tmp, "hits.txt"), "-d", os.path.join(tmp, "dump.gz")])
            shutil.copy(os.path.join(tmp, "hits.txt"), os.path.join(d, "hits_s%d.txt" % s))
            if s == 200:
                dump_to_npz(os.path.join(tmp, "dump.gz"), os.path.join(d, "dump.npz"))
        run_ref(d, ["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-t", 1, "-e",
                    "-o", os.path.join(tmp, "exact.txt")])
        shutil.copy(os.path.join(tmp, "exact.txt"), os.path.join(d, "exact.txt"))
        # exact mode with a low threshold so that more pairs are compared
        run_ref(d, ["-l", "list.txt", "-a", "reads.fa", "-k", 31, "-h", 12, "-t", 1, "-e", "-s", 0,
                    "-o", os.path.join(tmp, "exact0.txt")])
        shutil.copy(os.path.join(tmp, "exact0.txt"), os.path.join(d, "exact_s0.txt"))


def case_a_whole_files():
    """-A: whole-file queries (Miekki.cpp:487-514, :592-612) and -A -e (:616-645, :763-788)."""
    d = os.path.join(HERE, "caseA")
    qd = os.path.join(d, "qfiles")
    os.makedirs(qd, exist_ok=True)
    genomes = [H_genome(os.path.join(d, "gA%d.fa.gz" % g)) for g in range(14)]
    rng = np.random.default_rng(79)
    specs = [(0, 1000, 31000), (1, 20000, 52000), (3, 500, 9000), (7, 0, 100000), (5, 900, 7000)]
    names = []
    for i, (g, a, b) in enumerate(specs):
        seq = synth.substitute(np.frombuffer(genomes[g][a:b], np.uint8), 0.01, rng).tobytes()
        name = "qfiles/q%d.fa" % i
        with open(os.path.join(d, name), "wb") as f:
            f.write(b">whole%d from gA%d\n" % (i, g))
            for o in range(0, len(seq), 60):
                f.write(seq[o:o + 60] + b"\n")
        names.append(name)
    with open(os.path.join(d, "qfiles/short.fa"), "wb") as f:      # shorter than k: no output line
        f.write(b">short\nACGTACGTAC\n")
    names.append("qfiles/short.fa")
    with open(os.path.join(d, "qfiles/unrelated.fa"), "wb") as f:  # no hits: -A prints nothing
        f.write(b">unrelated\n" + synth.genome(998, 4000) + b"\n")
    names.append("qfiles/unrelated.fa")
    with open(os.path.join(d, "alist.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    with tempfile.TemporaryDirectory() as tmp:
        run_ref(d, ["-l", "list.txt", "-A", "alist.txt", "-k", 31, "-h", 12, "-t", 1,
                    "-o", os.path.join(tmp, "a.txt")])
        shutil.copy(os.path.join(tmp, "a.txt"), os.path.join(d, "hits_A.txt"))
        run_ref(d, ["-l", "list.txt", "-A", "alist.txt", "-k", 31, "-h", 12, "-t", 1, "-e",
                    "-o", os.path.join(tmp, "ae.txt")])
        shutil.copy(os.path.join(tmp, "ae.txt"), os.path.join(d, "exact_A.txt"))


def H_genome(path):
    with gzip.open(path, "rb") as f:
        return b"".join(l for l in f.read().split(b"\n") if l[:1] != b">")


def case_b():
    a = os.path.join(HERE, "caseA")
    d = os.path.join(HERE, "caseB")
    os.makedirs(d, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        run_ref(a, ["-l", "list.txt", "-a", "reads.fa", "-k", 21, "-h", 10, "-b", 34, "-t", 1, "-s", 50,
                    "-o", os.path.join(tmp, "hits.txt"), "-d", os.path.join(tmp, "dump.gz")])
        shutil.copy(os.path.join(tmp, "hits.txt"), os.path.join(d, "hits_s50.txt"))
        dump_to_npz(os.path.join(tmp, "dump.gz"), os.path.join(d, "dump.npz"))


def case_c_genomes(L=1_500_000):
    """14 relatives of one ancestor: every read matches all of them, all with
    intersection 0 -> the all-ties order of the bounded heap decides the list.
    (tests/ regenerate the genomes with this function and check the sha256.)"""
    rng = np.random.default_rng(78)
    base = np.frombuffer(synth.genome(100, L), np.uint8)
    return [synth.substitute(base, 0.002 * g, rng).tobytes() for g in range(14)]


def case_c():
    d = os.path.join(HERE, "caseC")
    os.makedirs(d, exist_ok=True)
    L = 1_500_000
    genomes = case_c_genomes(L)
    reads = synth.sample_reads(genomes, 16, 5000, sub_rate=0.01, block=7)
    synth.write_reads(os.path.join(d, "reads.fa"), reads)
    meta = {"recipe": "see make_golden.py:case_c", "genome_len": L,
            "genome_sha256": [hashlib.sha256(s).hexdigest() for s in genomes]}
    with tempfile.TemporaryDirectory() as tmp:
        names = []
        for g, s in enumerate(genomes):
            p = os.path.join(tmp, "gC%d.fa" % g)
            synth.write_fasta(p, ">gC%d" % g, s)
            names.append(p)
        with open(os.path.join(tmp, "list.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        for s in (200, 0):
            run_ref(tmp, ["-l", "list.txt", "-a", os.path.join(d, "reads.fa"), "-k", 31, "-h", 16,
                          "-t", 1, "-s", s, "-o", "hits.txt", "-d", "dump.gz"])
            shutil.copy(os.path.join(tmp, "hits.txt"), os.path.join(d, "hits_s%d.txt" % s))
        dd = orc.parse_dump(os.path.join(tmp, "dump.gz"))
        nz = np.flatnonzero(dd.bloom)
        meta.update(rows_sha256=sha(dd.rows), genome_size=[int(x) for x in dd.genome_size],
                    sketch_size=[int(x) for x in dd.sketch_size],
                    bloom_nonzero=int(len(nz)), bloom_idx_sha256=sha(nz.astype(np.uint32)),
                    bloom_val_sha256=sha(dd.bloom[nz]), k=dd.k, h=dd.h, b=dd.b)
        np.save(os.path.join(d, "rows_sample.npy"), dd.rows[::1024].copy())
    with open(os.path.join(d, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


def case_d_genomes(L=40_000):
    rng = np.random.default_rng(81)
    founders = [synth.genome(200 + g, L) for g in range(5)]
    genomes = list(founders)
    for j in range(65):                       # families of 14: more candidates than the 10 heap slots
        src = np.frombuffer(founders[j % 5], np.uint8)
        genomes.append(synth.substitute(src, 0.003 + 0.0005 * j, rng).tobytes())
    return genomes


def case_d():
    d = os.path.join(HERE, "caseD")
    os.makedirs(d, exist_ok=True)
    genomes = case_d_genomes()
    reads = synth.sample_reads(genomes, 48, 1500, sub_rate=0.01, block=11)
    reads += synth.sample_reads(genomes[:5], 12, 4000, sub_rate=0.0, block=12)
    synth.write_reads(os.path.join(d, "reads.fa"), reads)
    meta = {"recipe": "see make_golden.py:case_d", "genome_sha256": [hashlib.sha256(s).hexdigest() for s in genomes]}
    with tempfile.TemporaryDirectory() as tmp:
        names = []
        for g, s in enumerate(genomes):
            p = os.path.join(tmp, "gD%d.fa" % g)
            synth.write_fasta(p, ">gD%d" % g, s)
            names.append(p)
        with open(os.path.join(tmp, "list.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        for s in (200, 0):
            run_ref(tmp, ["-l", "list.txt", "-a", os.path.join(d, "reads.fa"), "-k", 31, "-h", 12, "-t", 1, "-s", s,
                          "-o", "hits.txt", "-d", "dump.gz"])
            shutil.copy(os.path.join(tmp, "hits.txt"), os.path.join(d, "hits_s%d.txt" % s))
        dump_to_npz(os.path.join(tmp, "dump.gz"), os.path.join(d, "dump.npz"))
        run_ref(tmp, ["-l", "list.txt", "-a", os.path.join(d, "reads.fa"), "-k", 31, "-h", 12, "-t", 1, "-e",
                      "-o", "exact.txt"])
        # exact lines end with the genome file path: keep the base name only
        with open(os.path.join(tmp, "exact.txt")) as f, open(os.path.join(d, "exact.txt"), "w") as g:
            for line in f:
                parts = line.rstrip("\n").split("\t")
                if len(parts) > 1:
                    parts[-1] = os.path.basename(parts[-1])
                g.write("\t".join(parts) + "\n")
    with open(os.path.join(d, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    if not REF.available:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    case_a()
    case_a_whole_files()
    case_b()
    case_c()
    case_d()
    print("golden vectors written under", HERE)
