"""ctypes binding of the C ABI in include/miekki_b200.h.

This is what tests/ and bench.py drive; the product's host side is the C++ CLI in
miekki_b200/cli/ (the reference is compiled code).  The class mirrors the subset of
``class Miekki`` (Miekki.h:33-137) that main.cpp uses: ``insert_sequences``,
``query_sequences`` + ``filter_results`` (fused here as ``query``), ``dump_disk`` /
the loader constructor (``export`` / ``import_``).

There is no CPU fallback: importing works anywhere (so that CPU-only tests can check
that the library loads and exports its symbols), but ``Miekki(...)`` raises when no
B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MIEKKI_B200_LIB points at another build of the same ABI (A/B measurements)
LIB_PATH = os.environ.get("MIEKKI_B200_LIB") or os.path.join(HERE, "libmiekki_b200.so")

HIT_DTYPE = np.dtype([("genome", "<u4"), ("matches", "<u4"),
                      ("jaccard", "<f8"), ("intersection", "<f8")])


class Stats(C.Structure):
    _fields_ = [("sketch_ms", C.c_double), ("read_sketch_ms", C.c_double), ("scan_ms", C.c_double),
                ("topk_ms", C.c_double), ("exact_ms", C.c_double), ("scan_launches", C.c_uint64),
                ("scan_row_bytes", C.c_uint64), ("scan_rows", C.c_uint64),
                ("bases_sketched", C.c_uint64), ("bases_queried", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MiekkiError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); also the list of symbols include/miekki_b200.h declares
_vp, _u32, _u64, _i, _d = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "mk_abi_version": (_i, []),
    "mk_create": (_i, [_u32, _u32, _u32, _u32, _u32, _u32, _i, _pp]),
    "mk_destroy": (None, [_vp]),
    "mk_last_error": (C.c_char_p, [_vp]),
    "mk_set_stream": (_i, [_vp, _vp]),
    "mk_set_scan_spare_sms": (_i, [_vp, _i]),
    "mk_set_shard": (_i, [_vp, _u32]),
    "mk_get_params": (_i, [_vp] + [C.POINTER(_u32)] * 6),
    "mk_batch_upload": (_i, [_vp, _vp, _vp, _u32, _pp]),
    "mk_batch_upload_flat": (_i, [_vp, _vp, _vp, _vp, _u32, _pp]),
    "mk_batch_synth": (_i, [_vp, _u64, _u32, _u32, _u64, _pp]),
    "mk_batch_download": (_i, [_vp, _vp, _u32, _vp, _u64]),
    "mk_batch_size": (_u32, [_vp]),
    "mk_batch_bases": (_u64, [_vp]),
    "mk_batch_free": (None, [_vp, _vp]),
    "mk_index_reserve": (_i, [_vp, _u32]),
    "mk_index_add": (_i, [_vp, _vp, _vp, _u32]),
    "mk_index_add_batch": (_i, [_vp, _vp]),
    "mk_index_size": (_i, [_vp, C.POINTER(_u32)]),
    "mk_index_stats": (_i, [_vp, _u32, _u32, _vp, _vp]),
    "mk_index_export": (_i, [_vp, _vp, _vp, _vp, _u64, _vp]),
    "mk_index_import": (_i, [_vp, _u32, _vp, _u64, _vp, _vp, _u64, _vp]),
    "mk_index_merge": (_i, [_vp, _vp]),
    "mk_bloom_reach": (_u64, [_u32, _u32]),
    "mk_bloom_window": (_u64, [_vp]),
    "mk_bloom_get": (_i, [_vp, _vp, _u64]),
    "mk_bloom_merge": (_i, [_vp, _vp, _u64]),
    "mk_bloom_set": (_i, [_vp, _vp, _u64]),
    "mk_scan": (_i, [_vp, _vp]),
    "mk_topk": (_i, [_vp, _u32, _u32, _d, _vp, _vp, _i, _i]),
    "mk_scan_async": (_i, [_vp, _vp, C.POINTER(_i)]),
    "mk_sketch_async": (_i, [_vp, _vp]),
    "mk_topk_slot": (_i, [_vp, _i, _u32, _u32, _d, _vp, _vp, _i, _i]),
    "mk_topk_slot_range": (_i, [_vp, _i, _u32, _u32, _u32, _u32, _d, _vp, _vp, _i, _i]),
    "mk_query": (_i, [_vp, _vp, _vp, _u32, _u32, _u32, _d, _vp, _vp]),
    "mk_query_batch": (_i, [_vp, _vp, _u32, _u32, _d, _vp, _vp]),
    "mk_query_chain": (_i, [_vp, _vp, _u32, _u32, _d, _vp, _vp, _i]),
    "mk_query_counts": (_i, [_vp, _vp, _vp, _u32, _vp, _vp]),
    "mk_sketch": (_i, [_vp, C.c_char_p, _u64, _vp, _vp, C.POINTER(_u32)]),
    "mk_exact": (_i, [_vp, _vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp, C.POINTER(_u64)]),
    "mk_exact_many": (_i, [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mk_index_export_rows": (_i, [_vp, _u64, _u64, _vp, _u64]),
    "mk_index_import_begin": (_i, [_vp, _u32]),
    "mk_index_import_rows": (_i, [_vp, _u64, _u64, _vp, _u64]),
    "mk_index_import_end": (_i, [_vp, _vp, _vp, _u64, _vp]),
    "mk_exact_batch": (_i, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_u64)]),
    "mk_stats_get": (_i, [_vp, C.POINTER(Stats)]),
    "mk_stats_reset": (_i, [_vp]),
    "mk_sync": (_i, [_vp]),
}


def lib() -> C.CDLL:
    """Loads libmiekki_b200.so (built by ``make`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MiekkiError("%s is missing: run `make` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _seq_arrays(seqs):
    """list[bytes] -> (char*[] , u64 lens[]) kept alive by the returned tuple."""
    n = len(seqs)
    arr = (C.c_char_p * max(1, n))(*seqs)
    lens = np.array([len(s) for s in seqs] or [0], np.uint64)
    return arr, lens


class Batch:
    """Sequences resident in HBM (mk_batch)."""

    def __init__(self, owner: "Miekki", handle):
        self._owner, self._h = owner, handle

    def __len__(self):
        return int(lib().mk_batch_size(self._h))

    @property
    def bases(self) -> int:
        return int(lib().mk_batch_bases(self._h))

    def download(self, i: int, length: int) -> bytes:
        buf = C.create_string_buffer(length)
        self._owner._ck(lib().mk_batch_download(self._owner._ctx, self._h, i, buf, length))
        return buf.raw

    def free(self):
        if self._h:
            lib().mk_batch_free(self._owner._ctx, self._h)
            self._h = None

    def __del__(self):
        try:
            if self._owner._ctx:
                self.free()
        except Exception:
            pass


class Miekki:
    """One GPU's shard of the index.  Mirrors Miekki::Miekki(k, h, ...) (Miekki.h:66)."""

    def __init__(self, k=31, h=17, b=33, threshold=200, device=0, bits_per_min=8, bits_mantis=5):
        self._ctx = None
        ctx = C.c_void_p()
        rc = lib().mk_create(k, h, bits_per_min, bits_mantis, b, int(threshold), device, C.byref(ctx))
        if rc != 0:
            raise MiekkiError("mk_create failed (%d): %s" % (rc, lib().mk_last_error(None).decode()))
        self._ctx = ctx
        self.k, self.h, self.b, self.threshold = k, h, b, int(threshold)
        self.B = 1 << h

    def close(self):
        if self._ctx:
            lib().mk_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise MiekkiError("miekki_b200 error %d: %s" % (rc, lib().mk_last_error(self._ctx).decode()))

    # ---- plumbing ------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        self._ck(lib().mk_set_stream(self._ctx, cuda_stream))

    def set_scan_spare_sms(self, n: int):
        self._ck(lib().mk_set_scan_spare_sms(self._ctx, n))

    def set_shard(self, first_id: int):
        self._ck(lib().mk_set_shard(self._ctx, first_id))

    def upload(self, seqs) -> Batch:
        arr, lens = _seq_arrays(seqs)
        h = C.c_void_p()
        self._ck(lib().mk_batch_upload(self._ctx, arr, _ptr(lens), len(seqs), C.byref(h)))
        return Batch(self, h)

    def upload_flat(self, data: np.ndarray, offsets, lens) -> Batch:
        """data: uint8 array (ideally pinned); read i = data[offsets[i] : offsets[i] + lens[i]]."""
        return self.upload_flat_ptr(data.ctypes.data, offsets, lens)

    def upload_flat_ptr(self, ptr: int, offsets, lens) -> Batch:
        h = C.c_void_p()
        offsets = np.ascontiguousarray(offsets, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint64)
        self._ck(lib().mk_batch_upload_flat(self._ctx, C.c_void_p(ptr), _ptr(offsets), _ptr(lens),
                                            len(lens), C.byref(h)))
        return Batch(self, h)

    def synth(self, seed: int, first_g: int, n: int, length: int) -> Batch:
        h = C.c_void_p()
        self._ck(lib().mk_batch_synth(self._ctx, seed, first_g, n, length, C.byref(h)))
        return Batch(self, h)

    # ---- build (insert_sequences, Miekki.cpp:277) ------------------------------
    def reserve(self, n: int):
        self._ck(lib().mk_index_reserve(self._ctx, n))

    def insert_sequences(self, seqs):
        arr, lens = _seq_arrays(seqs)
        self._ck(lib().mk_index_add(self._ctx, arr, _ptr(lens), len(seqs)))

    def insert_batch(self, batch: Batch):
        self._ck(lib().mk_index_add_batch(self._ctx, batch._h))

    @property
    def n(self) -> int:
        v = C.c_uint32()
        self._ck(lib().mk_index_size(self._ctx, C.byref(v)))
        return v.value

    def stats_arrays(self):
        n = self.n
        ss, gs = np.zeros(n, np.uint32), np.zeros(n, np.uint64)
        self._ck(lib().mk_index_stats(self._ctx, 0, n, _ptr(ss), _ptr(gs)))
        return ss, gs

    def bloom_window(self) -> int:
        return int(lib().mk_bloom_window(self._ctx))

    def bloom_get(self) -> np.ndarray:
        w = self.bloom_window()
        out = np.zeros(w, np.uint8)
        self._ck(lib().mk_bloom_get(self._ctx, _ptr(out), w))
        return out

    def bloom_merge(self, src: np.ndarray):
        src = np.ascontiguousarray(src, np.uint8)
        self._ck(lib().mk_bloom_merge(self._ctx, _ptr(src), len(src)))

    def bloom_set(self, src: np.ndarray):
        src = np.ascontiguousarray(src, np.uint8)
        self._ck(lib().mk_bloom_set(self._ctx, _ptr(src), len(src)))

    def bloom_get_ptr(self, ptr: int, n: int):
        self._ck(lib().mk_bloom_get(self._ctx, C.c_void_p(ptr), n))

    def bloom_set_ptr(self, ptr: int, n: int):
        self._ck(lib().mk_bloom_set(self._ctx, C.c_void_p(ptr), n))

    def export(self, rows=True, bloom_bytes: int | None = None):
        """-> dict(rows[B,n], genome_size, bloom, sketch_size): dump_disk payload."""
        n = self.n
        r = np.empty((self.B, n), np.uint8) if rows else None
        gs, ss = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        nb = self.bloom_window() if bloom_bytes is None else bloom_bytes
        bl = np.zeros(nb, np.uint8)
        self._ck(lib().mk_index_export(self._ctx, _ptr(r), _ptr(gs), _ptr(bl), nb, _ptr(ss)))
        return {"rows": r, "genome_size": gs, "bloom": bl, "sketch_size": ss}

    def import_(self, rows: np.ndarray, genome_size, bloom, sketch_size):
        rows = np.ascontiguousarray(rows, np.uint8)
        n = rows.shape[1]
        gs = np.ascontiguousarray(genome_size, np.uint64)
        ss = np.ascontiguousarray(sketch_size, np.uint32)
        bl = np.ascontiguousarray(bloom, np.uint8)
        self._ck(lib().mk_index_import(self._ctx, n, _ptr(rows), rows.strides[0], _ptr(gs), _ptr(bl),
                                       len(bl), _ptr(ss)))

    def merge(self, other: "Miekki"):
        """merge_indexes (Miekki.cpp:901): `other`'s genomes follow ours; Bloom tables folded."""
        self._ck(lib().mk_index_merge(self._ctx, other._ctx))

    def export_rows(self, row0: int, nrows: int, out: np.ndarray | None = None, col0: int = 0) -> np.ndarray:
        """Rows [row0, row0+nrows) of the dump payload.  With `out` (uint8 [nrows, width]) this
        shard's columns land at out[:, col0 : col0 + n] (several shards fill one slab)."""
        n = self.n
        if out is None:
            out = np.empty((nrows, n), np.uint8)
        assert out.dtype == np.uint8 and out.shape[0] >= nrows and out.strides[1] == 1
        self._ck(lib().mk_index_export_rows(self._ctx, row0, nrows, C.c_void_p(out.ctypes.data + col0),
                                            out.strides[0]))
        return out

    def import_begin(self, n: int):
        self._ck(lib().mk_index_import_begin(self._ctx, n))

    def import_rows(self, row0: int, rows: np.ndarray, col0: int = 0):
        """rows: uint8 [nrows, width]; this shard takes columns [col0, col0 + n)."""
        assert rows.dtype == np.uint8 and rows.strides[1] == 1
        self._ck(lib().mk_index_import_rows(self._ctx, row0, rows.shape[0], C.c_void_p(rows.ctypes.data + col0),
                                            rows.strides[0]))

    def import_end(self, genome_size, bloom, sketch_size):
        gs = np.ascontiguousarray(genome_size, np.uint64)
        ss = np.ascontiguousarray(sketch_size, np.uint32)
        bl = np.ascontiguousarray(bloom, np.uint8)
        self._ck(lib().mk_index_import_end(self._ctx, _ptr(gs), _ptr(bl), len(bl), _ptr(ss)))

    # ---- query (query_sequences + filter_results, Miekki.cpp:344,409) ----------
    def query(self, seqs, nresults=10, min_score=10, min_intersection=None):
        """-> list of HIT_DTYPE arrays, one per read."""
        if min_intersection is None:
            min_intersection = 0.5 * self.threshold          # Miekki.cpp:437
        n = len(seqs)
        arr, lens = _seq_arrays(seqs)
        hits = np.zeros((max(1, n), nresults), HIT_DTYPE)
        nh = np.zeros(max(1, n), np.uint32)
        self._ck(lib().mk_query(self._ctx, arr, _ptr(lens), n, nresults, min_score,
                                float(min_intersection), _ptr(hits), _ptr(nh)))
        return [hits[i, : nh[i]] for i in range(n)]

    def query_batch(self, batch: Batch, nresults=10, min_score=10, min_intersection=None, fetch=True):
        if min_intersection is None:
            min_intersection = 0.5 * self.threshold
        n = len(batch)
        if not fetch:
            self._ck(lib().mk_query_batch(self._ctx, batch._h, nresults, min_score,
                                          float(min_intersection), None, None))
            return None
        hits = np.zeros((max(1, n), nresults), HIT_DTYPE)
        nh = np.zeros(max(1, n), np.uint32)
        self._ck(lib().mk_query_batch(self._ctx, batch._h, nresults, min_score,
                                      float(min_intersection), _ptr(hits), _ptr(nh)))
        return hits, nh

    def query_chain(self, batch: Batch, heap: np.ndarray, lens: np.ndarray, nresults=10, min_score=10,
                    min_intersection=None, finalize=False):
        if min_intersection is None:
            min_intersection = 0.5 * self.threshold
        assert heap.dtype == HIT_DTYPE and heap.flags.c_contiguous and lens.dtype == np.uint32
        self._ck(lib().mk_query_chain(self._ctx, batch._h, nresults, min_score, float(min_intersection),
                                      _ptr(heap), _ptr(lens), 1 if finalize else 0))

    def scan(self, batch: Batch):
        """Sketch the reads and scan this shard; the counts stay in HBM for topk()."""
        self._ck(lib().mk_scan(self._ctx, batch._h))

    def topk_ptr(self, heap_ptr: int, len_ptr: int, nresults=10, min_score=10, min_intersection=None,
                 chain_in=False, finalize=True):
        """Bounded-heap step on the stored counts; pointers may be host or device memory."""
        if min_intersection is None:
            min_intersection = 0.5 * self.threshold
        self._ck(lib().mk_topk(self._ctx, nresults, min_score, float(min_intersection),
                               C.c_void_p(heap_ptr), C.c_void_p(len_ptr), 1 if chain_in else 0,
                               1 if finalize else 0))

    def scan_async(self, batch: Batch) -> int:
        """Enqueue sketch + scan and return at once; -> slot (0/1) for topk_ptr(slot=...)."""
        slot = C.c_int(0)
        self._ck(lib().mk_scan_async(self._ctx, batch._h, C.byref(slot)))
        return slot.value

    def sketch_async(self, batch: Batch) -> None:
        """Enqueue the read sketch of the batch the next scan_async will be given."""
        self._ck(lib().mk_sketch_async(self._ctx, batch._h))

    def topk_slot_ptr(self, slot: int, heap_ptr: int, len_ptr: int, nresults=10, min_score=10,
                      min_intersection=None, chain_in=False, finalize=True, first=0, count=None):
        """Heap step on the counts of `slot`; with first / count, on that range of the batch's
        reads only (the pointers then address read `first`)."""
        if min_intersection is None:
            min_intersection = 0.5 * self.threshold
        if first == 0 and count is None:
            self._ck(lib().mk_topk_slot(self._ctx, slot, nresults, min_score, float(min_intersection),
                                        C.c_void_p(heap_ptr), C.c_void_p(len_ptr), 1 if chain_in else 0,
                                        1 if finalize else 0))
        else:
            self._ck(lib().mk_topk_slot_range(self._ctx, slot, first, 0xFFFFFFFF if count is None else count,
                                              nresults, min_score, float(min_intersection),
                                              C.c_void_p(heap_ptr), C.c_void_p(len_ptr),
                                              1 if chain_in else 0, 1 if finalize else 0))

    def topk(self, heap: np.ndarray, lens: np.ndarray, **kw):
        assert heap.dtype == HIT_DTYPE and heap.flags.c_contiguous and lens.dtype == np.uint32
        self.topk_ptr(heap.ctypes.data, lens.ctypes.data, nresults=heap.shape[1], **kw)

    def query_counts(self, seqs):
        """-> (counts[n_reads, n_genomes] u32, surviving[n_reads] u32)."""
        n = len(seqs)
        arr, lens = _seq_arrays(seqs)
        counts = np.zeros((max(1, n), max(1, self.n)), np.uint32)
        surv = np.zeros(max(1, n), np.uint32)
        self._ck(lib().mk_query_counts(self._ctx, arr, _ptr(lens), n, _ptr(counts), _ptr(surv)))
        return counts[:n, : self.n], surv[:n]

    def sketch(self, seq: bytes):
        """minhash_sketch_partition (Miekki.cpp:150) -> (fp[B], anc[B], active)."""
        fp = np.empty(self.B, np.uint8)
        anc = np.empty(self.B, np.uint64)
        act = C.c_uint32()
        self._ck(lib().mk_sketch(self._ctx, seq, len(seq), _ptr(fp), _ptr(anc), C.byref(act)))
        return fp, anc, act.value

    def exact(self, records, reads):
        """ground_truth_batch (Miekki.cpp:792) -> (|B|, nb_inter[], nb_union[])."""
        ra, rl = _seq_arrays(records)
        qa, ql = _seq_arrays(reads)
        inter = np.zeros(max(1, len(reads)), np.uint64)
        uni = np.zeros(max(1, len(reads)), np.uint64)
        nb = C.c_uint64()
        self._ck(lib().mk_exact(self._ctx, ra, _ptr(rl), len(records), qa, _ptr(ql), len(reads),
                                _ptr(inter), _ptr(uni), C.byref(nb)))
        return nb.value, inter[: len(reads)], uni[: len(reads)]

    def exact_many(self, genomes):
        """genomes: [(records, reads)] -> [(|B|, nb_inter[], nb_union[])] in one call (mk_exact_many)."""
        recs = [r for g in genomes for r in g[0]]
        reads = [r for g in genomes for r in g[1]]
        ra, rl = _seq_arrays(recs)
        qa, ql = _seq_arrays(reads)
        rc = np.array([len(g[0]) for g in genomes] or [0], np.uint32)
        qc = np.array([len(g[1]) for g in genomes] or [0], np.uint32)
        inter = np.zeros(max(1, len(reads)), np.uint64)
        uni = np.zeros(max(1, len(reads)), np.uint64)
        nb = np.zeros(max(1, len(genomes)), np.uint64)
        self._ck(lib().mk_exact_many(self._ctx, len(genomes), ra, _ptr(rl), _ptr(rc), qa, _ptr(ql), _ptr(qc),
                                     _ptr(inter), _ptr(uni), _ptr(nb)))
        out, q0 = [], 0
        for g, (_, rd) in enumerate(genomes):
            out.append((int(nb[g]), inter[q0:q0 + len(rd)], uni[q0:q0 + len(rd)]))
            q0 += len(rd)
        return out

    def exact_batch(self, records: Batch, reads: Batch):
        """Same with both sides already in HBM (mk_exact_batch)."""
        n = len(reads)
        inter = np.zeros(max(1, n), np.uint64)
        uni = np.zeros(max(1, n), np.uint64)
        nb = C.c_uint64()
        self._ck(lib().mk_exact_batch(self._ctx, records._h, reads._h, _ptr(inter), _ptr(uni), C.byref(nb)))
        return nb.value, inter[:n], uni[:n]

    # ---- measurement -----------------------------------------------------------
    def stats(self) -> dict:
        s = Stats()
        self._ck(lib().mk_stats_get(self._ctx, C.byref(s)))
        return s.as_dict()

    def stats_reset(self):
        self._ck(lib().mk_stats_reset(self._ctx))

    def sync(self):
        self._ck(lib().mk_sync(self._ctx))
