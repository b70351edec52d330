"""Shared test helpers: reference-style file parsing and golden loaders."""
from __future__ import annotations

import gzip
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def read_text(path: str) -> str:
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as f:
            return f.read().decode("latin-1")
    with open(path, "rb") as f:
        return f.read().decode("latin-1")


def genome_like_reference(path: str) -> bytes:
    """Miekki.cpp:559-567: every line not starting with '>' is appended."""
    return "".join(l for l in read_text(path).split("\n") if l[:1] != ">").encode("latin-1")


def reads_like_reference(path: str, k: int):
    """Miekki.cpp:458-474: strict 2-line records; reads shorter than k skipped."""
    lines = read_text(path).split("\n")
    out = []
    for i in range(0, len(lines) - 1, 2):
        head, seq = lines[i], lines[i + 1]
        if len(seq) >= k:
            out.append((head, seq.encode("latin-1")))
    return out


def load_list(case_dir: str):
    with open(os.path.join(case_dir, "list.txt")) as f:
        return [l for l in f.read().split("\n") if len(l) > 3]


def load_dump_npz(path: str):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def bloom_nonzero(bloom: np.ndarray):
    nz = np.flatnonzero(bloom)
    return nz.astype(np.uint32), bloom[nz]


def golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def fmt_g(x: float) -> str:
    """C++ ostream default formatting of a double (6 significant digits)."""
    return "%g" % x


def exact_line(real_jax, jac_est, inter, inter_est, header, fname) -> str:
    """Miekki.cpp:853"""
    return "\t".join([fmt_g(real_jax), fmt_g(jac_est), fmt_g(inter), fmt_g(inter_est), header, fname])


def case_d(tmp_dir: str | None = None):
    """Golden case D (70 genomes, regenerated from seeds and checked against the recorded
    sha256).  -> (case dir, genomes); with tmp_dir also writes gD<i>.fa + list.txt there."""
    import hashlib
    import json
    d = os.path.join(GOLDEN, "caseD")
    meta = json.load(open(os.path.join(d, "meta.json")))
    genomes = golden_module().case_d_genomes()
    assert [hashlib.sha256(s).hexdigest() for s in genomes] == meta["genome_sha256"], \
        "numpy stream changed: regenerate tests/golden with make_golden.py"
    if tmp_dir is not None:
        names = []
        for g, s in enumerate(genomes):
            name = "gD%d.fa" % g
            with open(os.path.join(tmp_dir, name), "wb") as f:
                f.write(b">gD%d\n" % g + s + b"\n")
            names.append(name)
        with open(os.path.join(tmp_dir, "list.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
    return d, genomes


def merge_scenario(cli: str, tmp_path, devices: str):
    """SURVEY.md 8 row f4 (Miekki::merge_indexes, Miekki.cpp:901-910, Bloom merge included): the
    genomes of golden case A are indexed as two separate dumps (ids 0-5 and 6-13: neither a
    multiple of 32, so the second lands inside a 32-genome plane group), `-i a -i b -d merged`
    joins them, and the merged dump must equal the reference binary's `-t 1` dump of the whole
    list (tests/golden/caseA/dump.npz) -- rows, statistics and every Bloom byte -- answer the
    same hit lines, keep exact mode working through the names side-car, and load in the
    reference binary.  `cli` is the miekki executable (real library or CPU stub)."""
    import subprocess
    from collections import Counter
    from oracle import oracle as orc
    d = os.path.join(GOLDEN, "caseA")

    def run(args, ok=True):
        r = subprocess.run([cli] + [str(a) for a in args], cwd=d, capture_output=True, text=True, timeout=900)
        if ok:
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return r

    names = load_list(d)
    parts = {"a": names[:6], "b": names[6:]}
    for tag, part in parts.items():
        (tmp_path / ("list_%s.txt" % tag)).write_text("\n".join(part) + "\n")
        run(["-l", tmp_path / ("list_%s.txt" % tag), "-k", 31, "-h", 12, "-t", 3, "-d", tmp_path / (tag + ".gz"),
             "-o", tmp_path / "unused.txt", "--devices", devices])
    out, merged = tmp_path / "hits.txt", tmp_path / "merged.gz"
    r = run(["-i", tmp_path / "a.gz", "-i", tmp_path / "b.gz", "-d", merged, "-a", "reads.fa", "-o", out,
             "--devices", devices])
    assert "Load sucessful" in r.stdout
    assert out.read_text() == open(os.path.join(d, "hits_s200.txt")).read()
    got = orc.parse_dump(str(merged))
    z = load_dump_npz(os.path.join(d, "dump.npz"))
    assert (got.k, got.h, got.nbm, got.nbmant, got.n, got.b, got.threshold) == (31, 12, 8, 5, 14, 33, 200)
    assert np.array_equal(got.rows, z["rows"])
    assert np.array_equal(got.genome_size, z["genome_size"]) and np.array_equal(got.sketch_size, z["sketch_size"])
    idx, val = bloom_nonzero(got.bloom)
    assert np.array_equal(idx, z["bloom_idx"]) and np.array_equal(val, z["bloom_val"])
    # both parts carried their names: exact mode works from the merged dump
    assert os.path.exists(str(merged) + ".names")
    oute = tmp_path / "exact.txt"
    run(["-i", merged, "-a", "reads.fa", "-e", "-o", oute, "--devices", devices])
    want = [l for l in open(os.path.join(d, "exact.txt")).read().split("\n") if l]
    assert Counter(l for l in oute.read_text().split("\n") if l) == Counter(want)
    # the reference binary answers the same from the merged file
    ref = orc.RefBinary()
    if ref.available:
        outr = tmp_path / "ref_hits.txt"
        ref.run(["-i", str(merged), "-a", os.path.join(d, "reads.fa"), "-t", "1", "-o", str(outr)], cwd=d)
        assert outr.read_text() == open(os.path.join(d, "hits_s200.txt")).read()
    # merging in the other order gives other ids: not the same index (sanity of the test itself)
    run(["-i", tmp_path / "b.gz", "-i", tmp_path / "a.gz", "-d", tmp_path / "ba.gz", "-o", tmp_path / "unused.txt",
         "--devices", devices])
    assert not np.array_equal(orc.parse_dump(str(tmp_path / "ba.gz")).rows, z["rows"])
    # indexes built with different parameters do not merge
    run(["-l", tmp_path / "list_b.txt", "-k", 31, "-h", 10, "-d", tmp_path / "b10.gz", "-o", tmp_path / "unused.txt"])
    r = run(["-i", tmp_path / "a.gz", "-i", tmp_path / "b10.gz", "-o", tmp_path / "unused.txt"], ok=False)
    assert r.returncode == 1 and "cannot merge" in r.stderr
    # a side-car that belongs to another dump is not trusted
    import shutil
    shutil.copy(str(tmp_path / "b.gz") + ".names", str(merged) + ".names")
    r = run(["-i", merged, "-a", "reads.fa", "-e", "-o", oute])
    assert "does not belong to this dump" in r.stderr and "exact mode needs the genome list" in r.stderr
