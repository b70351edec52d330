"""BASELINE.json config 1 in full, end to end, against the UNMODIFIED reference binary run on
the same box: 100 synthetic 5 Mbp genomes (-k 31 -h 17), 1,000 synthetic 10 kbp reads.

Both programs read the same FASTA files through their own CLIs; compared are the decompressed
dump payloads (rows, genome_size, Bloom bytes, sketch_size; header byte 32 is uninitialised in
the reference, quirk G7) and the hit-line files.  At -h 17 every 5 Mbp sketch is saturated, so
genome_size is 0 and every list is empty at the default -s (quirk G2): the run uses -s 0, where
all intersections are 0 and the lists are decided purely by the heap's tie order.

A reduced exact-mode (-e) run on the first genomes follows.  Sizes can be lowered with
MIEKKI_C1_GENOMES / MIEKKI_C1_READS for a quick check.
"""
import os
import subprocess
from collections import Counter

import numpy as np
import pytest

from miekki_b200 import synth
from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

NG = int(os.environ.get("MIEKKI_C1_GENOMES", "100"))
NR = int(os.environ.get("MIEKKI_C1_READS", "1000"))
L = 5_000_000
CLI = os.path.join(H.ROOT, "miekki_b200", "cli", "miekki")
REF = orc.RefBinary()


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    if not REF.available:
        pytest.fail("oracle/_ref/Miekki is missing: build it with `make oracle` where /root/reference exists")
    d = tmp_path_factory.mktemp("c1")
    genomes = []
    names = []
    for g in range(NG):
        s = synth.genome(g, L)
        p = d / ("genome%d.fa" % g)
        synth.write_fasta(str(p), ">genome%d" % g, s)
        names.append(str(p))
        if g < 8:
            genomes.append(s)
    (d / "genomes.txt").write_text("\n".join(names) + "\n")
    # reads: sources drawn over all genomes; regenerate a source genome on demand
    rng = np.random.default_rng(2_000_000)
    src = rng.integers(NG, size=NR)
    pos = rng.integers(L - 10_000, size=NR)
    cache = {}
    with open(d / "reads.fa", "wb") as f:
        for r in range(NR):
            g = int(src[r])
            if g not in cache:
                if len(cache) > 4:
                    cache.clear()
                cache[g] = synth.genome(g, L)
            f.write(b">read%d_g%d_p%d\n" % (r, g, int(pos[r])) + cache[g][int(pos[r]):int(pos[r]) + 10_000] + b"\n")
    return d


def run(binary, args, cwd):
    r = subprocess.run([binary] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout


def test_config1_dump_and_hit_lines_equal_reference(data):
    d = data
    nproc = os.cpu_count() or 8
    # reference: -t 1 so that ids are list order and Bloom byte values are deterministic
    out_ref = run(REF.path, ["-l", "genomes.txt", "-a", "reads.fa", "-k", 31, "-h", 17, "-t", 1, "-s", 0,
                             "-o", "ref_hits.txt", "-d", "ref.gz"], d)
    out_gpu = run(CLI, ["-l", "genomes.txt", "-a", "reads.fa", "-k", 31, "-h", 17, "-t", nproc, "-s", 0,
                        "-o", "gpu_hits.txt", "-d", "gpu.gz"], d)
    t_ref, t_gpu = REF.elapsed(out_ref), REF.elapsed(out_gpu)
    print("config 1 wall (build+dump, query): reference -t 1 %s s, miekki_b200 %s s" % (t_ref, t_gpu))
    a = (d / "ref_hits.txt").read_text()
    b = (d / "gpu_hits.txt").read_text()
    assert a.count("\n") == NR
    assert a == b
    ra, rb = orc.parse_dump(str(d / "ref.gz")), orc.parse_dump(str(d / "gpu.gz"))
    assert (ra.k, ra.h, ra.nbm, ra.nbmant, ra.n, ra.b, ra.bloom_bits, ra.threshold, ra.compressed) == \
           (rb.k, rb.h, rb.nbm, rb.nbmant, rb.n, rb.b, rb.bloom_bits, rb.threshold, rb.compressed)
    assert np.array_equal(ra.rows, rb.rows)
    assert np.array_equal(ra.genome_size, rb.genome_size) and not ra.genome_size.any()   # quirk G2
    assert np.array_equal(ra.sketch_size, rb.sketch_size)
    assert np.array_equal(ra.bloom, rb.bloom)
    # and the reference answers the same from OUR dump (-i), any thread count
    run(REF.path, ["-i", "gpu.gz", "-a", "reads.fa", "-t", nproc, "-o", "ref_from_gpu_dump.txt"], d)
    assert Counter((d / "ref_from_gpu_dump.txt").read_text().split("\n")) == Counter(a.split("\n"))


def test_config1_exact_mode_subset(data):
    d = data
    n = min(6, NG)
    names = (d / "genomes.txt").read_text().split("\n")[:n]
    (d / "few.txt").write_text("\n".join(names) + "\n")
    # reads cut from those genomes, -h 20 so that genome_size != 0 and candidates pass -s
    rng = np.random.default_rng(5)
    with open(d / "few_reads.fa", "wb") as f:
        for r in range(60):
            g = int(rng.integers(n))
            s = H.genome_like_reference(names[g])
            p = int(rng.integers(L - 2000))
            f.write(b">x%d_g%d\n" % (r, g) + synth.substitute(np.frombuffer(s[p:p + 2000], np.uint8), 0.02, rng).tobytes() + b"\n")
    nproc = os.cpu_count() or 8
    # -t 1: with more threads the reference's genome ids (and so tie order) are nondeterministic
    run(REF.path, ["-l", "few.txt", "-a", "few_reads.fa", "-h", 20, "-e", "-t", 1, "-o", "ref_exact.txt"], d)
    run(CLI, ["-l", "few.txt", "-a", "few_reads.fa", "-h", 20, "-e", "-t", nproc, "-o", "gpu_exact.txt"], d)
    a = [l for l in (d / "ref_exact.txt").read_text().split("\n") if l]
    b = [l for l in (d / "gpu_exact.txt").read_text().split("\n") if l]
    assert len(a) >= 50
    assert Counter(a) == Counter(b)
