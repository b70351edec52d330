"""Parity at the shapes of BASELINE configs 3 and 5, sampled:

* C3: one GPU's share of 100,000 genomes at -h 17 (12,500 columns, 2^17 rows), 10 kbp reads with
  5 % substitutions -- the read sketch then needs its largest shared-memory table (16,384 slots)
  and a read touches ~9,000 rows.  Counts of sampled reads == numpy on the EXPORTED matrix with
  the oracle's Bloom-masked sketch; hit lists == the oracle's filter at -s 0, where every
  intersection is 0 (saturated sketches, quirk G2) and the list is decided by the heap's tie order.
  Also reads just below / above the shared-memory sketch limit (sparse and dense sketch paths).
* C5: `-e` through both command lines (the unmodified reference binary and ours) on 5 Mbp genomes
  at -h 20: exact-mode lines as multisets.

MIEKKI_C3_GENOMES / MIEKKI_C5_GENOMES / MIEKKI_C5_READS lower the sizes for a quick check.
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

SEED = 0x5EED_B200
L = 5_000_000
N3 = int(os.environ.get("MIEKKI_C3_GENOMES", "12500"))
NG5 = int(os.environ.get("MIEKKI_C5_GENOMES", "24"))
NR5 = int(os.environ.get("MIEKKI_C5_READS", "300"))
CLI = os.path.join(H.ROOT, "miekki_b200", "cli", "miekki")
REF = orc.RefBinary()


@pytest.mark.parametrize("tiled", ["0", "1"])
def test_c3_shape_counts_and_all_ties_hit_lists(monkeypatch, tiled):
    """tiled = 1 forces the tiled scan (scan_tiled.cu) for the reads it can take (lists of at most
    16,380 entries), 0 the ring kernel; at this shape the tiled kernel is what a full batch runs."""
    import miekki_b200 as mk
    monkeypatch.setenv("MIEKKI_SCAN_TILED", tiled)
    K, Hb = 31, 17
    ix = mk.Miekki(k=K, h=Hb, threshold=200)
    ix.reserve(N3)
    for g0 in range(0, N3, 128):
        b = ix.synth(SEED, g0, min(128, N3 - g0), L)
        ix.insert_batch(b)
        b.free()
    e = ix.export()
    rows, bloom, ss, gsz = e["rows"], e["bloom"], e["sketch_size"], e["genome_size"]
    assert (ss == 1 << Hb).all() and not gsz.any()          # saturated sketches: genome_size 0 (quirk G2)
    long_reads, src, _ = synth.cb_reads_block(SEED, N3, L, 20, 10_000, 0.05, block=7)
    reads = [r.tobytes() for r in long_reads]
    # the sketch paths' boundary: 12,288 k-mers fit the shared-memory table, one more does not
    for n in (1000, 12_288 + K, 12_289 + K, 40_000):
        reads.append(synth.cb_bases(SEED, 3, 1234, n).tobytes())
    # the 40 kbp read (39,969 list entries) is beyond the tiled kernel's counters: its own batch
    counts, surv = ix.query_counts(reads[:-1])
    hits = ix.query(reads[:-1], 10, 10, 0.0)                 # -s 0
    c2, s2 = ix.query_counts(reads[-1:])
    counts, surv = np.vstack([counts, c2]), np.concatenate([surv, s2])
    hits = hits + ix.query(reads[-1:], 10, 10, 0.0)
    L_ = orc.lib()
    for i, s in enumerate(reads):
        fp, anc, _ = orc.sketch(s, K, Hb)
        act = np.flatnonzero(fp != 255)
        keep = np.array([b for b in act if L_.mko_bloom_check(bloom.ctypes.data, 33, int(anc[b]))], np.int64)
        assert surv[i] == len(keep), i
        want = (rows[keep] == fp[keep][:, None]).sum(axis=0, dtype=np.uint32)
        assert np.array_equal(counts[i], want), i
        hits_o = np.zeros(10, orc.HIT_DTYPE)
        n = L_.mko_filter(want.ctypes.data, N3, np.ascontiguousarray(ss).ctypes.data,
                          np.ascontiguousarray(gsz).ctypes.data, 10, 10, 0.0, hits_o.ctypes.data)
        assert len(hits[i]) == n, (i, len(s), len(hits[i]), n, int((want >= 10).sum()), hits[i], hits_o[:n])
        assert hits[i].tobytes() == hits_o[:n].tobytes(), i    # ids, matches and both doubles, bit for bit
        if i < 20:                                             # true matches on top of the 8-bit collisions
            assert counts[i][src[i]] > np.median(counts[i])
    assert surv[:20].mean() > 8000                             # A(q) ~ 9,000
    ix.close()


@pytest.fixture(scope="module")
def c5(tmp_path_factory):
    if not REF.available:
        pytest.fail("oracle/_ref/Miekki is missing: build it with `make oracle` where /root/reference exists")
    d = tmp_path_factory.mktemp("c5")
    names = []
    for g in range(NG5):
        p = d / ("genome%d.fa" % g)
        synth.write_fasta(str(p), ">genome%d" % g, synth.cb_bases(SEED, g, 0, L).tobytes())
        names.append(str(p))
    (d / "genomes.txt").write_text("\n".join(names) + "\n")
    rng = np.random.default_rng(55)
    clean, gs, ps = synth.cb_reads(SEED, NG5, L, NR5, 1000, 0.0, block=5)
    with open(d / "reads.fa", "wb") as f:
        for r in range(NR5):
            seq = clean[r] if r % 2 == 0 else synth.substitute(clean[r], 0.02, rng)
            f.write(b">read%d_g%d_p%d\n" % (r, int(gs[r]), int(ps[r])) + seq.tobytes() + b"\n")
    return d


def test_c5_exact_mode_lines_equal_reference(c5):
    d = c5
    nproc = os.cpu_count() or 8

    def run(binary, args):
        r = subprocess.run([binary] + [str(a) for a in args], cwd=d, capture_output=True, text=True, timeout=3000)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        return r.stdout
    # -t 1: with more threads the reference's genome ids (and so its tie order) are nondeterministic
    run(REF.path, ["-l", "genomes.txt", "-a", "reads.fa", "-h", 20, "-e", "-t", 1, "-o", "ref_exact.txt"])
    run(CLI, ["-l", "genomes.txt", "-a", "reads.fa", "-h", 20, "-e", "-t", nproc, "-o", "gpu_exact.txt"])
    a = [l for l in (d / "ref_exact.txt").read_text().split("\n") if l]
    b = [l for l in (d / "gpu_exact.txt").read_text().split("\n") if l]
    assert len(a) >= NR5 * 9 // 10                    # nearly every read finds at least its source genome
    assert Counter(a) == Counter(b)
    # error-free reads: all 970 k-mers of the read are in the source genome
    src_lines = [l for l in b if l.split("\t")[2] == "970"]
    assert len(src_lines) >= NR5 // 2 - 5
    # the approximate lists of the same run agree as well (file order, -t 1 ids)
    run(REF.path, ["-l", "genomes.txt", "-a", "reads.fa", "-h", 20, "-t", 1, "-o", "ref_hits.txt"])
    run(CLI, ["-l", "genomes.txt", "-a", "reads.fa", "-h", 20, "-t", nproc, "-o", "gpu_hits.txt"])
    assert (d / "ref_hits.txt").read_text() == (d / "gpu_hits.txt").read_text()
