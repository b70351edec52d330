"""ctypes front-end to the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Two checkers live here:

* ``Oracle`` -- the C restatement in ``miekki_oracle.c`` (kind "port"),
* ``RefBinary`` -- the unmodified reference compiled by ``oracle/Makefile ref``
  into ``oracle/_ref/Miekki`` (kind "reference"), driven through its CLI, plus
  parsers for its three observable surfaces: the ``-d`` dump
  (SURVEY.md Appendix C), the ``-a`` hit lines and the ``-e`` exact lines.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product package ``miekki_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
import zlib
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmiekki_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "Miekki")


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when the reference tree is present, the
    reference binary.  Building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", HERE])
    if ref and os.path.isdir(os.environ.get("MIEKKI_REF_SRC", "/root/reference")):
        subprocess.check_call(
            ["make", "-s", "-C", HERE, "ref", "REF=" + os.environ.get("MIEKKI_REF_SRC", "/root/reference")]
        )


class Hit(C.Structure):
    _fields_ = [("genome", C.c_uint32), ("matches", C.c_uint32),
                ("jaccard", C.c_double), ("intersection", C.c_double)]


HIT_DTYPE = np.dtype([("genome", "<u4"), ("matches", "<u4"),
                      ("jaccard", "<f8"), ("intersection", "<f8")])

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.mko_revhash64.restype = C.c_uint64
        L.mko_revhash64.argtypes = [C.c_uint64]
        L.mko_unrevhash64.restype = C.c_uint64
        L.mko_unrevhash64.argtypes = [C.c_uint64]
        L.mko_universal_hash.restype = C.c_uint64
        L.mko_universal_hash.argtypes = [C.c_uint64, C.c_uint32]
        L.mko_mantis.restype = C.c_uint8
        L.mko_mantis.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.mko_str2numstrand.restype = C.c_uint64
        L.mko_str2numstrand.argtypes = [C.c_char_p, C.c_uint64]
        L.mko_str2num.restype = C.c_uint64
        L.mko_str2num.argtypes = [C.c_char_p, C.c_uint32]
        L.mko_sketch.restype = C.c_uint32
        L.mko_sketch.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_void_p, C.c_void_p]
        L.mko_bloom_check.restype = C.c_int
        L.mko_bloom_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.mko_bloom_window.restype = C.c_uint64
        L.mko_bloom_window.argtypes = [C.c_uint32, C.c_uint32]
        L.mko_index_new.restype = C.c_void_p
        L.mko_index_new.argtypes = [C.c_uint32] * 6 + [C.c_uint64]
        L.mko_index_free.argtypes = [C.c_void_p]
        for name, rt in [("mko_index_n", C.c_uint32), ("mko_index_cap", C.c_uint32),
                         ("mko_index_rows", u8p), ("mko_index_sketch_size", u32p),
                         ("mko_index_genome_size", u64p), ("mko_index_bloom", u8p),
                         ("mko_index_bloom_bytes", C.c_uint64)]:
            f = getattr(L, name)
            f.restype = rt
            f.argtypes = [C.c_void_p]
        L.mko_index_set_n.argtypes = [C.c_void_p, C.c_uint32]
        L.mko_index_insert.restype = C.c_int64
        L.mko_index_insert.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.mko_query_counts.restype = C.c_uint32
        L.mko_query_counts.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.mko_filter.restype = C.c_uint32
        L.mko_filter.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                 C.c_uint32, C.c_double, C.c_void_p]
        L.mko_filter_chain.restype = C.c_uint32
        L.mko_filter_chain.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_uint32, C.c_uint32, C.c_double, C.c_void_p, C.c_uint32,
                                       C.c_int]
        L.mko_exact_genome_set.restype = C.c_void_p
        L.mko_exact_genome_set.argtypes = [C.POINTER(C.c_char_p), u64p, C.c_uint32, C.c_uint32, u64p]
        L.mko_free.argtypes = [C.c_void_p]
        L.mko_exact_read.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_uint32,
                                     u64p, u64p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def sketch(seq: bytes, k: int, h: int, nbm: int = 8, nbmant: int = 5):
    """minhash_sketch_partition -> (fp[B] u8, anc[B] u64, active)."""
    B = 1 << h
    fp = np.empty(B, np.uint8)
    anc = np.empty(B, np.uint64)
    active = lib().mko_sketch(seq, len(seq), k, h, nbm, nbmant, _ptr(fp), _ptr(anc))
    return fp, anc, int(active)


def bloom_window(k: int, b: int) -> int:
    return int(lib().mko_bloom_window(k, b))


class Oracle:
    """Index + query restatement (Miekki.cpp:277-397)."""

    def __init__(self, k=31, h=17, b=33, cap=16, nbm=8, nbmant=5, full_bloom=False):
        self.k, self.h, self.b, self.nbm, self.nbmant = k, h, b, nbm, nbmant
        self.B = 1 << h
        bloom_bytes = 0 if full_bloom else bloom_window(k, b)
        self._ix = lib().mko_index_new(k, h, nbm, nbmant, b, cap, bloom_bytes)
        self.cap = cap

    def __del__(self):
        if getattr(self, "_ix", None) and _lib is not None:
            _lib.mko_index_free(self._ix)
            self._ix = None

    # -- build ---------------------------------------------------------------
    def insert(self, seq: bytes) -> int:
        g = lib().mko_index_insert(self._ix, seq, len(seq))
        if g < 0:
            raise RuntimeError("oracle index full")
        return int(g)

    @property
    def n(self) -> int:
        return int(lib().mko_index_n(self._ix))

    @property
    def rows(self) -> np.ndarray:
        """bucket-major view [B, n]."""
        p = lib().mko_index_rows(self._ix)
        a = np.ctypeslib.as_array(p, shape=(self.B, self.cap))
        return a[:, : self.n]

    @property
    def sketch_size(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().mko_index_sketch_size(self._ix), shape=(self.cap,))[: self.n]

    @property
    def genome_size(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().mko_index_genome_size(self._ix), shape=(self.cap,))[: self.n]

    @property
    def bloom(self) -> np.ndarray:
        nb = int(lib().mko_index_bloom_bytes(self._ix))
        return np.ctypeslib.as_array(lib().mko_index_bloom(self._ix), shape=(nb,))

    def load(self, rows: np.ndarray, genome_size, bloom, sketch_size) -> None:
        """Adopt an existing index (e.g. parsed from a reference dump)."""
        n = rows.shape[1]
        assert rows.shape[0] == self.B and n <= self.cap
        p = lib().mko_index_rows(self._ix)
        full = np.ctypeslib.as_array(p, shape=(self.B, self.cap))
        full[:, :n] = rows
        lib().mko_index_set_n(self._ix, n)
        self.sketch_size[:] = sketch_size
        self.genome_size[:] = genome_size
        bl = self.bloom
        m = min(len(bl), len(bloom))
        bl[:m] = bloom[:m]

    # -- query ---------------------------------------------------------------
    def counts(self, seq: bytes, want_fp=False):
        c = np.zeros(self.n, np.uint32)
        fp = np.empty(self.B, np.uint8) if want_fp else None
        a = lib().mko_query_counts(self._ix, seq, len(seq), _ptr(c), _ptr(fp) if want_fp else None)
        return (c, int(a), fp) if want_fp else (c, int(a))

    def filter(self, counts: np.ndarray, nresults=10, min_score=10, min_intersection=100.0):
        hits = np.zeros(nresults, HIT_DTYPE)
        counts = np.ascontiguousarray(counts, np.uint32)
        ss = np.ascontiguousarray(self.sketch_size)
        gs = np.ascontiguousarray(self.genome_size)
        n = lib().mko_filter(_ptr(counts), len(counts), _ptr(ss), _ptr(gs), nresults, min_score,
                             float(min_intersection), _ptr(hits))
        return hits[:n]

    def query(self, seq: bytes, nresults=10, min_score=10, min_intersection=100.0):
        c, _ = self.counts(seq)
        return self.filter(c, nresults, min_score, min_intersection)


def filter_chain(counts, first_id, sketch_size, genome_size, nresults, min_score, min_intersection,
                 heap: np.ndarray, length: int, finalize: bool) -> int:
    """Shard-chained filter: ``heap`` (HIT_DTYPE[nresults]) carries the state."""
    counts = np.ascontiguousarray(counts, np.uint32)
    ss = np.ascontiguousarray(sketch_size, np.uint32)
    gs = np.ascontiguousarray(genome_size, np.uint64)
    return int(lib().mko_filter_chain(_ptr(counts), len(counts), first_id, _ptr(ss), _ptr(gs), nresults,
                                      min_score, float(min_intersection), _ptr(heap), length,
                                      1 if finalize else 0))


def format_hit_line(header: str, hits) -> str:
    """Miekki.cpp:438-445 (std::to_string(double) == '%f')."""
    out = [header, ":"]
    for hh in hits:
        out.append("%u\t%u\t%u\t%f;" % (int(hh["genome"]), int(hh["matches"]),
                                          int(np.uint32(int(hh["intersection"]) & 0xFFFFFFFF)),
                                          float(hh["jaccard"])))
    out.append("\n")
    return "".join(out)


# ---- exact mode --------------------------------------------------------------

def records_like_reference(fasta_text: str, k: int) -> list[bytes]:
    """Miekki.cpp:801-822: how ground_truth_batch splits a genome file into the
    strings whose k-mers go into set B.  A record shorter than k is NOT cleared
    at the next header and runs into the following record (quirk G16); the
    reference loops ``while(!eof) getline`` so a trailing newline yields one
    extra empty line, which is harmless."""
    recs, ref = [], ""
    for line in fasta_text.split("\n"):
        if line[:1] == ">":
            if len(ref) >= k:
                recs.append(ref.encode())
                ref = ""
        else:
            ref += line
    if len(ref) >= k:
        recs.append(ref.encode())
    return recs


def exact(records: list[bytes], reads: list[bytes], k: int):
    """-> (|B|, [(nb_inter, nb_union)] per read)  (Miekki.cpp:792-842)."""
    L = lib()
    arr = (C.c_char_p * max(1, len(records)))(*records)
    lens = np.array([len(r) for r in records] or [0], np.uint64)
    nB = C.c_uint64(0)
    setB = L.mko_exact_genome_set(arr, lens.ctypes.data_as(C.POINTER(C.c_uint64)), len(records), k,
                                  C.byref(nB))
    out = []
    try:
        for r in reads:
            a, u = C.c_uint64(0), C.c_uint64(0)
            L.mko_exact_read(setB, nB.value, r, len(r), k, C.byref(a), C.byref(u))
            out.append((a.value, u.value))
    finally:
        L.mko_free(setB)
    return nB.value, out


# ---- the reference binary ----------------------------------------------------

@dataclass
class Dump:
    k: int
    h: int
    nbm: int
    nbmant: int
    n: int
    b: int
    bloom_bits: int
    threshold: int
    compressed: int
    rows: np.ndarray          # [B, n] u8
    genome_size: np.ndarray   # [n] u64
    bloom: np.ndarray         # [2^b/8] u8 (may be empty)
    sketch_size: np.ndarray   # [n] u32


def read_payload(path: str) -> bytes:
    """gzip / zlib / plain auto-detect like zstr.hpp:154-167."""
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        raw = f.read()
    if len(head) == 2 and head[0] == 0x78 and head[1] in (0x01, 0x9C, 0xDA):
        return zlib.decompress(raw)
    return raw


def parse_dump(path: str) -> Dump:
    """SURVEY.md Appendix C / Miekki.cpp:649-678."""
    p = read_payload(path)
    k, h, nbm, nbmant, n, b = np.frombuffer(p, "<u4", 6, 0)
    bloom_bits = int(np.frombuffer(p, "<u8", 1, 24)[0])
    threshold = int(np.frombuffer(p, "<u4", 1, 34)[0])
    compressed = p[38]
    off = 39
    B = 1 << int(h)
    rows = np.frombuffer(p, np.uint8, B * int(n), off).reshape(B, int(n))
    off += B * int(n)
    gs = np.frombuffer(p, "<u8", int(n), off)
    off += 8 * int(n)
    bl = np.frombuffer(p, np.uint8, bloom_bits // 8, off)
    off += bloom_bits // 8
    ss = np.frombuffer(p, "<u4", int(n), off)
    off += 4 * int(n)
    assert off == len(p), (off, len(p))
    return Dump(int(k), int(h), int(nbm), int(nbmant), int(n), int(b), bloom_bits, threshold,
                int(compressed), rows, gs, bl, ss)


class RefBinary:
    """Runs oracle/_ref/Miekki (the unmodified reference)."""

    def __init__(self, path: str = REF_BIN):
        self.path = path

    @property
    def available(self) -> bool:
        return os.path.exists(self.path) and os.access(self.path, os.X_OK)

    def run(self, args: list[str], cwd: str | None = None, timeout: float | None = None) -> str:
        r = subprocess.run([self.path] + [str(a) for a in args], cwd=cwd, capture_output=True,
                           text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError("reference binary failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        return r.stdout

    @staticmethod
    def elapsed(stdout: str) -> list[float]:
        """The two 'elapsed time: Xs' lines (main.cpp:211,234): build, query."""
        import re
        # progress ticks ("-", no newline) may precede the text on the same line
        return [float(x) for x in re.findall(r"elapsed time: ([0-9.eE+-]+)s", stdout)]
