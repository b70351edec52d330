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
