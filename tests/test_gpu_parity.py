"""Parity of the CUDA path (through the C ABI) against the golden vectors of the reference
binary and against the CPU oracle on seeded inputs.  Bit-exact for sketches, rows, Bloom
bytes, counts and hit lists; jaccard / intersection are compared through the reference's
own text format (``%f``) and, numerically, at 1e-6 relative (north_star tolerance).
"""
import hashlib
import json
import os
from collections import Counter

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6   # north_star: estimated intersection and Jaccard within 1e-6 relative


@pytest.fixture(scope="module")
def mk():
    import miekki_b200
    return miekki_b200


def gpu_hit_lines(ix, reads, threshold, chunk=17):
    out = []
    for i in range(0, len(reads), chunk):
        part = reads[i:i + chunk]
        hits = ix.query([s for _, s in part], 10, 10, 0.5 * np.uint32(threshold))
        for (head, _), hh in zip(part, hits):
            out.append(orc.format_hit_line(head, hh))
    return "".join(out)


def assert_index_equals_dump(ix, z):
    e = ix.export(bloom_bytes=int(z["bloom_bits"]) // 8 if int(z["bloom_bits"]) <= (1 << 31) else None)
    assert np.array_equal(e["rows"], z["rows"])
    assert np.array_equal(e["sketch_size"], z["sketch_size"])
    assert np.array_equal(e["genome_size"], z["genome_size"])
    idx, val = H.bloom_nonzero(e["bloom"])
    assert np.array_equal(idx, z["bloom_idx"])
    assert np.array_equal(val, z["bloom_val"])


# ---- sketch kernel vs oracle ------------------------------------------------------------

def rand_seq(rng, n, special=False):
    s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if special and n > 50:
        for _ in range(3):
            p = int(rng.integers(0, n - 10))
            s[p:p + int(rng.integers(1, 10))] = ord("N")
        p = int(rng.integers(0, n - 20))
        s[p:p + 15] = np.frombuffer(bytes(s[p:p + 15]).lower(), np.uint8)
        s[int(rng.integers(n))] = ord("R")
    return s.tobytes()


@pytest.mark.parametrize("k,h", [(31, 12), (31, 17), (21, 10), (15, 8), (31, 20), (5, 3)])
def test_sketch_matches_oracle(mk, k, h):
    rng = np.random.default_rng(k * 100 + h)
    ix = mk.Miekki(k=k, h=h)
    lens = [0, 1, k - 2, k - 1, k, k + 1, k + 2, 47, 48, 49, 100, 1000, 4097, 20000, 70001]
    for i, n in enumerate(lens):
        for special in (False, True):
            s = rand_seq(rng, n, special)
            fp, anc, act = ix.sketch(s)
            ofp, oanc, oact = orc.sketch(s, k, h)
            assert act == oact, (n, special)
            assert np.array_equal(fp, ofp), (n, special)
            assert np.array_equal(anc, oanc), (n, special)
    # invalid characters inside the first k-1 positions (prefix encoder, utils.cpp:252)
    s = bytearray(rand_seq(rng, 500))
    s[3] = ord("N")
    s[7] = ord("c")
    for variant in (bytes(s), bytes(s).lower(), bytes(s[:k + 5])):
        fp, anc, act = ix.sketch(variant)
        ofp, oanc, oact = orc.sketch(variant, k, h)
        assert act == oact and np.array_equal(fp, ofp) and np.array_equal(anc, oanc)
    ix.close()


# ---- golden case A: build, dump, hit lines ------------------------------------------------

@pytest.fixture(scope="module")
def case_a(mk):
    d = os.path.join(H.GOLDEN, "caseA")
    genomes = [H.genome_like_reference(os.path.join(d, f)) for f in H.load_list(d)]
    ix = mk.Miekki(k=31, h=12, b=33, threshold=200)
    # uneven insert calls: ids must still be call order
    ix.insert_sequences(genomes[:1])
    ix.insert_sequences(genomes[1:9])
    ix.insert_sequences(genomes[9:])
    yield d, ix, genomes
    ix.close()


def test_case_a_index_equals_reference_dump(case_a):
    d, ix, _ = case_a
    assert ix.n == 14
    assert_index_equals_dump(ix, H.load_dump_npz(os.path.join(d, "dump.npz")))


@pytest.mark.parametrize("s", [200, 0, 5000])
def test_case_a_hit_lines_equal_reference(case_a, s):
    d, ix, _ = case_a
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    want = open(os.path.join(d, "hits_s%d.txt" % s)).read()
    assert gpu_hit_lines(ix, reads, s) == want


def test_case_a_counts_and_floats_equal_oracle(case_a):
    d, ix, genomes = case_a
    o = orc.Oracle(k=31, h=12, b=33, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    reads = [s for _, s in H.reads_like_reference(os.path.join(d, "reads.fa"), 31)]
    counts, surv = ix.query_counts(reads)
    hits = ix.query(reads, 10, 10, 100.0)
    for i, s in enumerate(reads):
        oc, oa = o.counts(s)
        assert np.array_equal(counts[i], oc), i
        assert surv[i] == oa, i
        oh = o.filter(oc, 10, 10, 100.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"])
        assert np.array_equal(hits[i]["matches"], oh["matches"])
        np.testing.assert_allclose(hits[i]["jaccard"], oh["jaccard"], rtol=REL_TOL, atol=0)
        np.testing.assert_allclose(hits[i]["intersection"], oh["intersection"], rtol=REL_TOL, atol=0)


def test_case_a_export_import_roundtrip(mk, case_a):
    d, ix, _ = case_a
    e = ix.export()
    ix2 = mk.Miekki(k=31, h=12, b=33, threshold=200)
    ix2.import_(e["rows"], e["genome_size"], e["bloom"], e["sketch_size"])
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    assert gpu_hit_lines(ix2, reads, 200) == open(os.path.join(d, "hits_s200.txt")).read()
    ix2.close()


def test_case_a_streamed_export_import(mk, case_a):
    """mk_index_export_rows / mk_index_import_{begin,rows,end}: the matrix in slabs, two shards
    filling their columns of one host slab, equals the one-shot calls."""
    d, ix, _ = case_a
    whole = ix.export()
    B, n = whole["rows"].shape
    got = np.concatenate([ix.export_rows(r0, min(1000, B - r0)) for r0 in range(0, B, 1000)])
    assert np.array_equal(got, whole["rows"])
    assert np.array_equal(got, H.load_dump_npz(os.path.join(d, "dump.npz"))["rows"])
    split = 5
    a, b = mk.Miekki(k=31, h=12), mk.Miekki(k=31, h=12)
    a.import_begin(split)
    b.import_begin(n - split)
    for r0 in range(0, B, 700):
        slab = np.ascontiguousarray(whole["rows"][r0:r0 + 700])
        a.import_rows(r0, slab, 0)
        b.import_rows(r0, slab, split)
    a.import_end(whole["genome_size"][:split], whole["bloom"], whole["sketch_size"][:split])
    b.import_end(whole["genome_size"][split:], whole["bloom"], whole["sketch_size"][split:])
    b.set_shard(split)
    out = np.zeros((B, n), np.uint8)
    for r0 in range(0, B, 512):
        a.export_rows(r0, 512, out[r0:r0 + 512], 0)
        b.export_rows(r0, 512, out[r0:r0 + 512], split)
    assert np.array_equal(out, whole["rows"])
    ea, eb = a.export(rows=False), b.export(rows=False)
    assert np.array_equal(np.concatenate([ea["genome_size"], eb["genome_size"]]), whole["genome_size"])
    assert np.array_equal(np.concatenate([ea["sketch_size"], eb["sketch_size"]]), whole["sketch_size"])
    assert np.array_equal(ea["bloom"], whole["bloom"]) and np.array_equal(eb["bloom"], whole["bloom"])
    with pytest.raises(mk.MiekkiError):
        mk.Miekki(k=31, h=12).import_rows(0, whole["rows"][:4].copy())    # no import_begin
    a.close()
    b.close()


def test_case_a_sharded_chain_equals_unsharded(mk, case_a):
    """Two genome shards (ids 0-5 and 6-13), Bloom merged "lowest rank wins", heap chained
    in ascending id order: identical lines to the single-index run."""
    d, ix, genomes = case_a
    a = mk.Miekki(k=31, h=12, b=33, threshold=200)
    b = mk.Miekki(k=31, h=12, b=33, threshold=200)
    a.insert_sequences(genomes[:6])
    b.insert_sequences(genomes[6:])
    b.set_shard(6)
    bl_a, bl_b = a.bloom_get(), b.bloom_get()
    a.bloom_merge(bl_b)          # a keeps its own bytes, takes b's where it had none
    merged = a.bloom_get()
    b_new = mk.Miekki(k=31, h=12, b=33, threshold=200)
    eb = b.export()
    b_new.import_(eb["rows"], eb["genome_size"], merged, eb["sketch_size"])
    b_new.set_shard(6)
    assert np.array_equal(merged, ix.bloom_get())      # byte-exact with the single build
    assert np.array_equal(np.where(bl_a != 0, bl_a, bl_b), merged)
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    for s in (200, 0):
        seqs = [x for _, x in reads]
        heap = np.zeros((len(seqs), 10), mk.HIT_DTYPE)
        lens = np.zeros(len(seqs), np.uint32)
        ba, bb = a.upload(seqs), b_new.upload(seqs)
        a.query_chain(ba, heap, lens, 10, 10, 0.5 * s, finalize=False)
        b_new.query_chain(bb, heap, lens, 10, 10, 0.5 * s, finalize=True)
        got = "".join(orc.format_hit_line(h, heap[i, :lens[i]]) for i, (h, _) in enumerate(reads))
        assert got == open(os.path.join(d, "hits_s%d.txt" % s)).read()
        ba.free(); bb.free()
    for x in (a, b, b_new):
        x.close()


def test_case_a_concurrent_callers(mk, case_a):
    """The C ABI serialises calls per context, so host threads may share one (the reference's
    workflows call into a shared Miekki object from every OpenMP thread)."""
    import threading
    d, ix, _ = case_a
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    want = open(os.path.join(d, "hits_s200.txt")).read().split("\n")
    out, errs = {}, []

    def work(t):
        try:
            part = reads[t::4]
            got = gpu_hit_lines(ix, part, 200, chunk=5).split("\n")
            out[t] = got[:-1]
        except Exception as e:          # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs
    for t in range(4):
        assert out[t] == want[t:len(reads):4]


def test_case_a_read_slices_bound_list_memory(mk, case_a, monkeypatch):
    """mk_query cuts a batch into slices of reads so that list memory stays bounded; with a tiny
    budget (several slices, slices of one read) the lines must not change."""
    d, ix, _ = case_a
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    want = open(os.path.join(d, "hits_s200.txt")).read()
    for budget in (1, 3000, 20000):
        monkeypatch.setenv("MIEKKI_LIST_BUDGET_ENTRIES", str(budget))
        assert gpu_hit_lines(ix, reads, 200, chunk=64) == want
    monkeypatch.delenv("MIEKKI_LIST_BUDGET_ENTRIES")


@pytest.mark.parametrize("tiled", ["0", "1"])
def test_case_a_sketch_ahead_of_the_scan(mk, case_a, monkeypatch, tiled):
    """mk_sketch_async: the lists of the next batch are built before mk_scan_async is called with it
    (both scan kernels); a different batch than the one sketched ahead is sketched as usual, and so
    is a batch sketched ahead of an index change."""
    monkeypatch.setenv("MIEKKI_SCAN_TILED", tiled)
    d, ix, _ = case_a
    reads = [s for _, s in H.reads_like_reference(os.path.join(d, "reads.fa"), 31)]
    parts = [reads[:20], reads[20:], reads[7:31]]
    batches = [ix.upload(p) for p in parts]
    want = [ix.query(p, 10, 10, 100.0) for p in parts]

    def step(slot, i):
        heap = np.zeros((len(parts[i]), 10), mk.HIT_DTYPE)
        lens = np.zeros(len(parts[i]), np.uint32)
        ix.topk_slot_ptr(slot, heap.ctypes.data, lens.ctypes.data, 10, 10, 100.0)
        for j in range(len(parts[i])):
            assert heap[j, :lens[j]].tobytes() == want[i][j].tobytes(), (i, j)

    ix.sketch_async(batches[0])
    s0 = ix.scan_async(batches[0])             # uses the lists sketched ahead
    ix.sketch_async(batches[1])                # beside the scan of batch 0
    step(s0, 0)
    s1 = ix.scan_async(batches[1])
    ix.sketch_async(batches[0])                # sketched ahead, but batch 2 comes next
    step(s1, 1)
    s2 = ix.scan_async(batches[2])
    step(s2, 2)
    launches = ix.stats()["kernel_launches"]
    ix.sketch_async(batches[1])
    assert ix.stats()["kernel_launches"] > launches
    s3 = ix.scan_async(batches[1])
    step(s3, 1)
    ix.sync()
    for b in batches:
        b.free()


def test_case_a_topk_slot_in_read_ranges(mk, case_a):
    """mk_topk_slot_range: the heap step over a batch in ragged ranges of reads (in any order,
    pointers addressing the range's first read) gives the lists of the whole-batch step; a range
    past the batch is clipped to nothing."""
    d, ix, _ = case_a
    reads = [s for _, s in H.reads_like_reference(os.path.join(d, "reads.fa"), 31)]
    n = len(reads)
    b = ix.upload(reads)
    for s in (200, 0):
        want = ix.query(reads, 10, 10, 0.5 * s)
        heap = np.zeros((n, 10), mk.HIT_DTYPE)
        lens = np.zeros(n, np.uint32)
        slot = ix.scan_async(b)
        for first, count in ((17, 9), (0, 17), (26, None), (n, 5), (n + 3, 1)):
            off = min(first, n - 1)
            ix.topk_slot_ptr(slot, heap[off:].ctypes.data, lens[off:].ctypes.data, 10, 10, 0.5 * s,
                             first=first, count=count)
        ix.sync()
        for i in range(n):
            assert heap[i, :lens[i]].tobytes() == want[i].tobytes(), (s, i)
    b.free()


def test_case_a_pipelined_scan_and_topk_slots(mk, case_a):
    """mk_scan_async / mk_topk_slot: two read batches in flight (scan of the second enqueued
    before the first one's heap step) give the same lists as the plain query."""
    d, ix, _ = case_a
    reads = [s for _, s in H.reads_like_reference(os.path.join(d, "reads.fa"), 31)]
    parts = [reads[:20], reads[20:], reads[5:33]]
    batches = [ix.upload(p) for p in parts]
    want = [ix.query(p, 10, 10, 100.0) for p in parts]
    heaps = [np.zeros((len(p), 10), mk.HIT_DTYPE) for p in parts]
    lens = [np.zeros(len(p), np.uint32) for p in parts]
    slots = [ix.scan_async(batches[0]), ix.scan_async(batches[1])]
    assert sorted(slots) == [0, 1]
    ix.topk_slot_ptr(slots[0], heaps[0].ctypes.data, lens[0].ctypes.data, 10, 10, 100.0)
    slots.append(ix.scan_async(batches[2]))          # reuses the tile of batch 0
    assert slots[2] == slots[0]
    ix.topk_slot_ptr(slots[1], heaps[1].ctypes.data, lens[1].ctypes.data, 10, 10, 100.0)
    ix.topk_slot_ptr(slots[2], heaps[2].ctypes.data, lens[2].ctypes.data, 10, 10, 100.0)
    ix.sync()
    for p, w, hp, ln in zip(parts, want, heaps, lens):
        for i in range(len(p)):
            assert hp[i, :ln[i]].tobytes() == w[i].tobytes()
    for b in batches:
        b.free()
    st = ix.stats()
    assert st["scan_rows"] > 0 and st["scan_row_bytes"] == st["scan_rows"] * ix.n


# ---- exact mode ---------------------------------------------------------------------------

@pytest.mark.parametrize("sort_variant", ["0", "1"])
@pytest.mark.parametrize("s,fname", [(200, "exact.txt"), (0, "exact_s0.txt")])
def test_case_a_exact_lines_equal_reference(case_a, s, fname, sort_variant, monkeypatch):
    """sort_variant = 1: set B as a radix-sorted array with binary search (MIEKKI_EXACT_SORT), the
    sort-merge wording of BASELINE.json's north_star; 0: the hash sets that ship by default."""
    monkeypatch.setenv("MIEKKI_EXACT_SORT", sort_variant)
    d, ix, _ = case_a
    k = 31
    names = H.load_list(d)
    per_genome = {}
    for head, seq in H.reads_like_reference(os.path.join(d, "reads.fa"), k):
        if seq[:1] not in (b"A", b"C", b"G", b"T", b"N"):            # Miekki.cpp:736
            continue
        for hit in ix.query([seq], 5, 10, float(s))[0]:              # :741
            per_genome.setdefault(int(hit["genome"]), []).append((head, seq, hit))
    lines = []
    for g, items in per_genome.items():
        recs = orc.records_like_reference(H.read_text(os.path.join(d, names[g])), k)
        nB, inter, uni = ix.exact(recs, [seq for _, seq, _ in items])
        # cross-check the raw integers with the oracle too
        onB, ores = orc.exact(recs, [seq for _, seq, _ in items], k)
        assert nB == onB
        assert [(int(a), int(u)) for a, u in zip(inter, uni)] == ores
        for (head, _, hit), a, u in zip(items, inter, uni):
            if a > 0:
                lines.append(H.exact_line(int(a) / int(u), hit["jaccard"], int(a), hit["intersection"],
                                          head, names[g]))
    want = [l for l in open(os.path.join(d, fname)).read().split("\n") if l]
    assert Counter(lines) == Counter(want)


@pytest.mark.parametrize("sort_variant", ["0", "1"])
def test_exact_edge_cases(mk, monkeypatch, sort_variant):
    monkeypatch.setenv("MIEKKI_EXACT_SORT", sort_variant)
    ix = mk.Miekki(k=31, h=10)
    rng = np.random.default_rng(3)
    g = rand_seq(rng, 5000, special=True)
    recs = [g[:2000], g[2000:2020], g[2020:], b"", b"ACGT"]
    reads = [g[100:400], g[1990:2030], b"ACGT", b"", rand_seq(rng, 300), g[:31], g.lower()[500:900]]
    nB, inter, uni = ix.exact(recs, reads)
    onB, ores = orc.exact(recs, reads, 31)
    assert nB == onB
    assert [(int(a), int(u)) for a, u in zip(inter, uni)] == ores
    nB, inter, uni = ix.exact([], [g[:100]])
    assert nB == 0 and int(inter[0]) == 0 and int(uni[0]) == 70
    # the same through batches that already sit in HBM (mk_exact_batch)
    gb, rb = ix.upload(recs), ix.upload(reads)
    nB, inter, uni = ix.exact_batch(gb, rb)
    assert nB == onB and [(int(a), int(u)) for a, u in zip(inter, uni)] == ores
    gb.free()
    rb.free()
    # several genome files in one call (mk_exact_many): per genome the same numbers
    g2 = rand_seq(rng, 7000, special=True)
    many = [(recs, reads), ([g2[:3000], g2[3000:]], [g2[10:500], g[100:400], b"ACGT"]), ([], [g[:100]]),
            ([g2], [])]
    got = ix.exact_many(many)
    for (rr, qq), (nB, inter, uni) in zip(many, got):
        oB, ores2 = orc.exact(rr, qq, 31)
        assert nB == oB and [(int(a), int(u)) for a, u in zip(inter, uni)] == ores2
    ix.close()


def test_case_d_seventy_genomes_equal_reference(mk):
    """Golden case D: 70 genomes (two build chunks, three 32-genome groups), full 10-hit lists
    picked among 14 candidates per family -- dump and hit lines of the reference binary."""
    d, genomes = H.case_d()
    ix = mk.Miekki(k=31, h=12, b=33, threshold=200)
    ix.insert_sequences(genomes[:3])                 # uneven calls: chunks start inside a group
    ix.insert_sequences(genomes[3:])
    assert_index_equals_dump(ix, H.load_dump_npz(os.path.join(d, "dump.npz")))
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    for s in (200, 0):
        assert gpu_hit_lines(ix, reads, s) == open(os.path.join(d, "hits_s%d.txt" % s)).read()
    ix.close()


# ---- other geometries ---------------------------------------------------------------------

def test_case_b_k21_h10_b34(mk):
    a = os.path.join(H.GOLDEN, "caseA")
    d = os.path.join(H.GOLDEN, "caseB")
    genomes = [H.genome_like_reference(os.path.join(a, f)) for f in H.load_list(a)]
    ix = mk.Miekki(k=21, h=10, b=34, threshold=50)
    ix.insert_sequences(genomes)
    assert_index_equals_dump(ix, H.load_dump_npz(os.path.join(d, "dump.npz")))
    reads = H.reads_like_reference(os.path.join(a, "reads.fa"), 21)
    assert gpu_hit_lines(ix, reads, 50) == open(os.path.join(d, "hits_s50.txt")).read()
    ix.close()


def test_case_c_saturated_sketch_all_ties(mk):
    d = os.path.join(H.GOLDEN, "caseC")
    meta = json.load(open(os.path.join(d, "meta.json")))
    genomes = H.golden_module().case_c_genomes(meta["genome_len"])
    assert [hashlib.sha256(s).hexdigest() for s in genomes] == meta["genome_sha256"]
    ix = mk.Miekki(k=meta["k"], h=meta["h"], b=meta["b"], threshold=200)
    ix.insert_sequences(genomes)
    e = ix.export()
    assert hashlib.sha256(e["rows"].tobytes()).hexdigest() == meta["rows_sha256"]
    assert list(map(int, e["genome_size"])) == meta["genome_size"]          # all 0: quirk G2
    assert list(map(int, e["sketch_size"])) == meta["sketch_size"]
    idx, val = H.bloom_nonzero(e["bloom"])
    assert len(idx) == meta["bloom_nonzero"]
    assert hashlib.sha256(idx.tobytes()).hexdigest() == meta["bloom_idx_sha256"]
    assert hashlib.sha256(val.tobytes()).hexdigest() == meta["bloom_val_sha256"]
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), meta["k"])
    for s in (200, 0):
        assert gpu_hit_lines(ix, reads, s) == open(os.path.join(d, "hits_s%d.txt" % s)).read()
    ix.close()


# ---- seeded random parity at sizes the oracle finishes in seconds ---------------------------

@pytest.mark.parametrize("narrow", ["32", "0"])
def test_random_index_and_long_reads_vs_oracle(mk, monkeypatch, narrow):
    monkeypatch.setenv("MIEKKI_SCAN_NARROW_GROUPS", narrow)
    _random_index_and_long_reads(mk)


def _random_index_and_long_reads(mk, short_only=False):
    """40 genomes (more than one dense chunk), reads of 60 bp .. 40 kbp: the long ones take the
    dense read path, the short ones the shared-memory path."""
    rng = np.random.default_rng(11)
    k, h = 31, 14
    base = [rand_seq(rng, int(rng.integers(30_000, 120_000)), special=(i % 5 == 0)) for i in range(8)]
    genomes = []
    for i in range(40):
        src = np.frombuffer(base[i % 8], np.uint8).copy()
        m = rng.random(len(src)) < 0.01 * (i // 8)
        src[m] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(m.sum()))]
        genomes.append(src.tobytes())
    ix = mk.Miekki(k=k, h=h, threshold=20)
    ix.insert_sequences(genomes)
    o = orc.Oracle(k=k, h=h, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    e = ix.export()
    assert np.array_equal(e["rows"], o.rows)
    assert np.array_equal(e["sketch_size"], o.sketch_size)
    assert np.array_equal(e["genome_size"], o.genome_size)
    assert np.array_equal(e["bloom"][: len(o.bloom)], o.bloom[: len(e["bloom"])])
    reads = []
    for n in ((60, 200, 1000, 5000, 12000, 3000, 800, 31, 32, 10, 7000, 1500, 100) if short_only else
              (60, 200, 1000, 5000, 12000, 12350, 20000, 40000, 31, 32, 10)):
        g = genomes[int(rng.integers(len(genomes)))]
        p = int(rng.integers(0, len(g) - n)) if len(g) > n else 0
        reads.append(g[p:p + n])
    counts, surv = ix.query_counts(reads)
    hits = ix.query(reads, 10, 10, 10.0)
    for i, s in enumerate(reads):
        if len(s) < k:
            assert surv[i] == 0 and counts[i].sum() == 0 and len(hits[i]) == 0
            continue
        oc, oa = o.counts(s)
        assert surv[i] == oa, (i, len(s))
        assert np.array_equal(counts[i], oc), (i, len(s))
        oh = o.filter(oc, 10, 10, 10.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"])
        assert np.array_equal(hits[i]["matches"], oh["matches"])
        np.testing.assert_allclose(hits[i]["intersection"], oh["intersection"], rtol=REL_TOL)
    # the same reads through the pipelined pair with the sketch enqueued ahead of the scan: a batch
    # with a read past the shared-memory sketch's reach is left to mk_scan_async (dense read path),
    # a batch of short reads (ragged, some shorter than k) is sketched ahead
    for part in (reads, [r for r in reads if len(r) <= 12000]):
        b = ix.upload(part)
        heap = np.zeros((len(part), 10), mk.HIT_DTYPE)
        lens = np.zeros(len(part), np.uint32)
        ix.sketch_async(b)
        slot = ix.scan_async(b)
        ix.topk_slot_ptr(slot, heap.ctypes.data, lens.ctypes.data, 10, 10, 10.0)
        ix.sync()
        want = ix.query(part, 10, 10, 10.0)
        for i in range(len(part)):
            assert heap[i, :lens[i]].tobytes() == want[i].tobytes(), (i, len(part[i]))
        b.free()
    ix.close()


@pytest.mark.parametrize("tile_bytes", ["1", "700", "2000"])
def test_query_in_many_count_tiles_with_long_reads(mk, monkeypatch, tile_bytes):
    """The query pipeline scans a batch in tiles of reads bounded by the count-tile buffer (1 GiB:
    ~10,000 reads at 12,500 genomes) while the top-k of the previous tile runs beside it.
    MIEKKI_COUNT_TILE_BYTES shrinks the buffer so that this small case crosses many tiles, both
    for a batch with long reads (dense sketch, lists built at once) and without (sketched tile by
    tile, one ahead of the scan)."""
    monkeypatch.setenv("MIEKKI_COUNT_TILE_BYTES", tile_bytes)
    _random_index_and_long_reads(mk)
    monkeypatch.setenv("MIEKKI_SCAN_NARROW_GROUPS", "0")
    _random_index_and_long_reads(mk, short_only=True)


@pytest.mark.parametrize("narrow", ["32", "0"])
def test_whole_genome_query_crosses_counter_chunks(mk, monkeypatch, narrow):
    """A query with more than 65,504 surviving buckets (a 1.2 Mbp sequence at -h 17) makes the scan
    spill its 16-plane carry-save counters and accumulate across chunks (F_ACCUM path).  narrow=0
    keeps the shared-memory ring kernel on this 6-genome shard, the default is the warp-per-read
    kernel (whose own spill needs > 65,504 rows per lane: next test)."""
    monkeypatch.setenv("MIEKKI_SCAN_NARROW_GROUPS", narrow)
    rng = np.random.default_rng(21)
    k, h = 31, 17
    base = np.frombuffer(rand_seq(rng, 1_200_000), np.uint8)
    genomes = []
    for i in range(5):
        g = base.copy()
        m = rng.random(len(g)) < 0.002 * i
        g[m] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(m.sum()))]
        genomes.append(g.tobytes())
    genomes.append(rand_seq(rng, 300_000))
    ix = mk.Miekki(k=k, h=h, threshold=0)
    ix.insert_sequences(genomes)
    o = orc.Oracle(k=k, h=h, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    q = [genomes[2], genomes[0][:700_000], genomes[5]]
    counts, surv = ix.query_counts(q)
    assert surv[0] > 65_504
    for i, s in enumerate(q):
        oc, oa = o.counts(s)
        assert surv[i] == oa
        assert np.array_equal(counts[i], oc), i
    hits = ix.query(q, 10, 10, 0.0)
    for i, s in enumerate(q):
        oh = o.query(s, 10, 10, 0.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"]) and np.array_equal(hits[i]["matches"], oh["matches"])
    ix.close()


@pytest.mark.parametrize("n_genomes", [600, 40])
def test_narrow_scan_spills_counters(mk, n_genomes):
    """scan_narrow_kernel: 600 genomes = 19 groups -> 32 lanes per row, one row slot, so a query
    with more than 65,504 surviving buckets overflows the lane's 16-plane counters and takes the
    flush-and-accumulate path; 40 genomes (2 lanes per row, 16 row slots) do not, same answers."""
    rng = np.random.default_rng(33)
    k, h = 31, 17
    big = rand_seq(rng, 900_000)
    genomes = [big] + [rand_seq(rng, 3_000) for _ in range(n_genomes - 2)] + [big[:400_000] + rand_seq(rng, 1000)]
    ix = mk.Miekki(k=k, h=h, threshold=0)
    ix.insert_sequences(genomes)
    o = orc.Oracle(k=k, h=h, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    q = [big, genomes[-1], genomes[5], big[100_000:100_500]]
    counts, surv = ix.query_counts(q)
    assert surv[0] > 65_504
    for i, s in enumerate(q):
        oc, oa = o.counts(s)
        assert surv[i] == oa
        assert np.array_equal(counts[i], oc), i
    ix.close()


def test_wide_index_tiles_and_padding(mk):
    """Index wider than one scan tile is not testable cheaply with real sketches; import a
    random matrix instead (N = 20,001: two genome tiles, ragged last 16-byte group) and
    compare counts and hits with the oracle's scoring of the same matrix."""
    rng = np.random.default_rng(5)
    k, h, N = 31, 8, 20_001
    B = 1 << h
    rows = rng.integers(0, 256, (B, N), dtype=np.uint8)
    ss = rng.integers(1, B + 1, N).astype(np.uint32)
    gs = rng.integers(0, 5_000_000, N).astype(np.uint64)
    o = orc.Oracle(k=k, h=h, cap=N)
    bloom = np.ones(len(o.bloom), np.uint8)                 # every k-mer passes
    o.load(rows, gs, bloom, ss)
    ix = mk.Miekki(k=k, h=h, threshold=0)
    ix.import_(rows, gs, bloom, ss)
    reads = [rand_seq(rng, n) for n in (300, 2000, 40)]
    counts, surv = ix.query_counts(reads)
    hits = ix.query(reads, 10, 1, 0.0)
    for i, s in enumerate(reads):
        oc, oa = o.counts(s)
        assert surv[i] == oa
        assert np.array_equal(counts[i], oc)
        oh = o.filter(oc, 10, 1, 0.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"])
        assert np.array_equal(hits[i]["matches"], oh["matches"])
    ix.close()


def test_synthetic_generator_matches_host(mk):
    from miekki_b200 import synth
    ix = mk.Miekki(k=31, h=10)
    b = ix.synth(1234, 7, 3, 5000)
    for i in range(3):
        assert b.download(i, 5000) == synth.cb_bases(1234, 7 + i, 0, 5000).tobytes()
    b.free()
    ix.close()


def test_sketch_arbitrary_bytes_fuzz(mk):
    """Any byte is legal input (NUL, high-bit, lowercase, IUPAC): the reference maps what it does
    not know to code 0 on both strands (utils.cpp:31-49,107-125).  200 random strings."""
    rng = np.random.default_rng(99)
    ix = mk.Miekki(k=21, h=9)
    alphabets = [np.arange(256, dtype=np.uint8), np.frombuffer(b"ACGTacgtNnRYKM-*\x00\xff", np.uint8),
                 np.frombuffer(b"ACGTN", np.uint8)]
    for it in range(200):
        al = alphabets[it % 3]
        n = int(rng.integers(0, 900))
        s = al[rng.integers(0, len(al), n)].tobytes()
        fp, anc, act = ix.sketch(s)
        ofp, oanc, oact = orc.sketch(s, 21, 9)
        assert act == oact and np.array_equal(fp, ofp) and np.array_equal(anc, oanc), (it, n)
    ix.close()


@pytest.mark.parametrize("k,h,b,nres", [(2, 1, 32, 1), (3, 5, 33, 3), (11, 13, 36, 64), (16, 9, 40, 10),
                                        (17, 18, 33, 10), (31, 22, 34, 5), (25, 7, 32, 12)])
def test_parameter_corners_vs_oracle(mk, k, h, b, nres):
    """Corners of the parameter space (k 2..31, h 1..22, b 32..40, nresults 1..64): index, counts
    and hit lists against the oracle on small seeded inputs."""
    rng = np.random.default_rng(k * 1000 + h)
    base = rand_seq(rng, 6000, special=True)
    genomes = [base, rand_seq(rng, 3000), base[:2000] + rand_seq(rng, 2500), rand_seq(rng, 40 + k),
               base[1000:5000]]
    ix = mk.Miekki(k=k, h=h, b=b, threshold=2)
    ix.insert_sequences(genomes)
    o = orc.Oracle(k=k, h=h, b=b, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    e = ix.export()
    assert np.array_equal(e["rows"], o.rows)
    assert np.array_equal(e["sketch_size"], o.sketch_size)
    assert np.array_equal(e["genome_size"], o.genome_size)
    m = min(len(e["bloom"]), len(o.bloom))
    assert np.array_equal(e["bloom"][:m], o.bloom[:m])
    reads = [base[100:700], base[3000:3300], genomes[1][5:900], rand_seq(rng, 300), genomes[3], base[:k + 1]]
    counts, surv = ix.query_counts(reads)
    hits = ix.query(reads, nres, 1, 1.0)
    for i, s in enumerate(reads):
        oc, oa = o.counts(s)
        assert surv[i] == oa, i
        assert np.array_equal(counts[i], oc), i
        oh = o.filter(oc, nres, 1, 1.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"]) and np.array_equal(hits[i]["matches"], oh["matches"])
        np.testing.assert_allclose(hits[i]["intersection"], oh["intersection"], rtol=REL_TOL, atol=0)
    ix.close()


def test_empty_and_degenerate_inputs(mk):
    """Empty index, empty batches, reads with no valid k-mer, low-complexity sequences."""
    k, h = 31, 10
    ix = mk.Miekki(k=k, h=h, threshold=0)
    rng = np.random.default_rng(8)
    r1 = rand_seq(rng, 500)
    # empty index: every read gets an empty list (and an output line in the CLI)
    assert [len(x) for x in ix.query([r1, b"ACGT", b""], 10, 0, 0.0)] == [0, 0, 0]
    c, surv = ix.query_counts([r1])
    assert c.shape == (1, 0) and surv[0] == 0          # nothing is in the Bloom filter yet
    assert ix.query([]) == []
    ix.insert_sequences([])                              # no-op
    assert ix.n == 0
    genomes = [rand_seq(rng, 20_000), b"A" * 5_000, b"N" * 3_000, (b"ACGT" * 2_000), b"acgt" * 1_000 + r1]
    ix.insert_sequences(genomes)
    o = orc.Oracle(k=k, h=h, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    e = ix.export()
    assert np.array_equal(e["rows"], o.rows)
    assert np.array_equal(e["sketch_size"], o.sketch_size) and np.array_equal(e["genome_size"], o.genome_size)
    reads = [b"A" * 200, b"N" * 200, b"ACGT" * 50, genomes[0][100:400], b"acgt" * 50, r1, b"", b"ACGTN" * 30,
             genomes[0][:31], genomes[0][:32]]
    counts, surv = ix.query_counts(reads)
    hits = ix.query(reads, 10, 1, 0.0)
    for i, s in enumerate(reads):
        oc, oa = (o.counts(s) if len(s) >= 1 else (np.zeros(len(genomes), np.uint32), 0))
        assert surv[i] == oa, i
        assert np.array_equal(counts[i], oc), i
        oh = o.filter(oc, 10, 1, 0.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"]) and np.array_equal(hits[i]["matches"], oh["matches"])
    ix.close()


def test_plain_key_format_matches_oracle(mk, monkeypatch):
    """Sequences of 2^32 bases or more cannot carry the k-mer tag in their sketch keys and take
    the plain key format; MIEKKI_KEY_TAGS=0 forces that path on ordinary inputs."""
    monkeypatch.setenv("MIEKKI_KEY_TAGS", "0")
    k, h = 27, 12
    rng = np.random.default_rng(77)
    genomes = [rand_seq(rng, 60_000, special=True) for _ in range(5)] + [rand_seq(rng, k)]
    ix = mk.Miekki(k=k, h=h, threshold=0)
    ix.insert_sequences(genomes)
    o = orc.Oracle(k=k, h=h, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    e = ix.export()
    assert np.array_equal(e["rows"], o.rows)
    assert np.array_equal(e["sketch_size"], o.sketch_size) and np.array_equal(e["genome_size"], o.genome_size)
    m = min(len(e["bloom"]), len(o.bloom))
    assert np.array_equal(e["bloom"][:m], o.bloom[:m])
    fp, anc, act = ix.sketch(genomes[0])
    ofp, oanc, oact = orc.sketch(genomes[0], k, h)
    assert act == oact and np.array_equal(fp, ofp) and np.array_equal(anc, oanc)
    ix.close()


def test_build_with_saturated_bloom_pages(mk):
    """Build fast path (sketch.cu: resolve_kernel): once 4 KB pages of the Bloom table hold no zero
    byte, buckets whose k-mer can only touch such pages skip the look-up.  k = 25, b = 32 has a
    32 KiB window: after 16 genomes its first pages are full, the last ones never are, so later
    inserts take both paths.  Everything the build leaves behind must still equal the oracle."""
    k, h, b = 25, 14, 32
    rng = np.random.default_rng(5)
    genomes = [rand_seq(rng, 150_000, special=(g % 5 == 0)) for g in range(48)]
    genomes.append(genomes[3][:k])                      # exactly k bases: no k-mer, sketch_size 0
    genomes.append(genomes[7][1000:90_000])
    ix = mk.Miekki(k=k, h=h, b=b, threshold=0)
    o = orc.Oracle(k=k, h=h, b=b, cap=len(genomes))
    for first in range(0, len(genomes), 16):
        ix.insert_sequences(genomes[first:first + 16])
        for s in genomes[first:first + 16]:
            o.insert(s)
        e = ix.export()
        m = min(len(e["bloom"]), len(o.bloom))
        assert np.array_equal(e["bloom"][:m], o.bloom[:m]), first
        if first == 0:
            pages = (np.asarray(e["bloom"][:32768]).reshape(8, 4096) != 0).all(axis=1)
            assert pages[:3].all() and not pages[7], pages      # both paths are live from here on
    assert np.array_equal(e["rows"], o.rows)
    assert np.array_equal(e["sketch_size"], o.sketch_size)
    assert np.array_equal(e["genome_size"], o.genome_size)
    assert e["sketch_size"][48] == 0
    reads = [genomes[40][500:2500], genomes[7][2000:4000], genomes[20][100_000:101_000]]
    counts, surv = ix.query_counts(reads)
    for i, s in enumerate(reads):
        oc, oa = o.counts(s)
        assert surv[i] == oa and np.array_equal(counts[i], oc), i
    ix.close()


def test_build_with_tiny_saturated_bloom_window(mk):
    """-k 21 -b 33 (a BASELINE sweep point) probes only 64 table bytes: they are all set after the
    first genomes, the padding behind them never is, and the build must still recognise the page
    as saturated (bytes no k-mer can reach do not count) without changing any result."""
    k, h, b = 21, 12, 33
    rng = np.random.default_rng(9)
    genomes = [rand_seq(rng, 50_000, special=(g % 3 == 0)) for g in range(12)]
    ix = mk.Miekki(k=k, h=h, b=b, threshold=0)
    o = orc.Oracle(k=k, h=h, b=b, cap=len(genomes))
    for first in range(0, len(genomes), 4):
        ix.insert_sequences(genomes[first:first + 4])
        for s in genomes[first:first + 4]:
            o.insert(s)
        e = ix.export()
        m = min(len(e["bloom"]), len(o.bloom))
        assert np.array_equal(e["bloom"][:m], o.bloom[:m]), first
        assert (e["bloom"][:64] != 0).all() and not e["bloom"][64:m].any()
    assert np.array_equal(e["rows"], o.rows)
    assert np.array_equal(e["sketch_size"], o.sketch_size) and np.array_equal(e["genome_size"], o.genome_size)
    counts, surv = ix.query_counts([genomes[3][100:3000]])
    oc, oa = o.counts(genomes[3][100:3000])
    assert surv[0] == oa and np.array_equal(counts[0], oc)
    ix.close()


def test_errors(mk):
    with pytest.raises(mk.MiekkiError):
        mk.Miekki(k=32)
    with pytest.raises(mk.MiekkiError):
        mk.Miekki(bits_per_min=16)          # -f 11: "not implemented" (quirk G12)
    with pytest.raises(mk.MiekkiError):
        mk.Miekki(b=31)
    ix = mk.Miekki(k=31, h=10)
    with pytest.raises(mk.MiekkiError):
        ix.insert_sequences([b"ACGT"])      # shorter than k
    assert ix.n == 0
    hits = ix.query([b"ACGTACGTACGTACGTACGTACGTACGTACGTACGT"])     # empty index: one empty hit list per read
    assert len(hits) == 1 and len(hits[0]) == 0
    ix.close()


def test_smem_optin_failure_is_reported(mk, monkeypatch):
    """The ring scan needs > 48 KB of dynamic shared memory, granted per device by
    cudaFuncSetAttribute (kernels.h: smem_optin).  When that fails the call must return an error
    with a message -- not launch, not fall back; afterwards the same context works again."""
    rng = np.random.default_rng(6)
    k, h, N = 31, 8, 1500                                   # > 1,024 genomes: ring kernel
    B = 1 << h
    rows = rng.integers(0, 256, (B, N), dtype=np.uint8)
    ss = rng.integers(1, B + 1, N).astype(np.uint32)
    gs = rng.integers(0, 5_000_000, N).astype(np.uint64)
    o = orc.Oracle(k=k, h=h, cap=N)
    bloom = np.ones(len(o.bloom), np.uint8)
    o.load(rows, gs, bloom, ss)
    ix = mk.Miekki(k=k, h=h, threshold=0)
    ix.import_(rows, gs, bloom, ss)
    read = rand_seq(rng, 900)
    monkeypatch.setenv("MIEKKI_TEST_FAIL_SMEM_OPTIN", "1")
    with pytest.raises(mk.MiekkiError, match="scan launch configuration failed"):
        ix.query_counts([read])
    monkeypatch.delenv("MIEKKI_TEST_FAIL_SMEM_OPTIN")
    counts, surv = ix.query_counts([read])
    oc, oa = o.counts(read)
    assert surv[0] == oa and np.array_equal(counts[0], oc)
    ix.close()


def test_empty_last_shard_still_sorts_the_heap(mk, case_a):
    """mk_query_chain with finalize on a shard that holds no genome: the chained heap must still
    go through sort_heap (Miekki.cpp:396), as mk_topk_slot does."""
    d, ix, genomes = case_a
    a = mk.Miekki(k=31, h=12, b=33, threshold=200)
    a.insert_sequences(genomes)
    empty = mk.Miekki(k=31, h=12, b=33, threshold=200)
    empty.set_shard(len(genomes))
    empty.bloom_set(a.bloom_get())
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    seqs = [x for _, x in reads]
    for s in (200, 0):
        heap = np.zeros((len(seqs), 10), mk.HIT_DTYPE)
        lens = np.zeros(len(seqs), np.uint32)
        ba, be = a.upload(seqs), empty.upload(seqs)
        a.query_chain(ba, heap, lens, 10, 10, 0.5 * s, finalize=False)
        empty.query_chain(be, heap, lens, 10, 10, 0.5 * s, finalize=True)
        got = "".join(orc.format_hit_line(hd, heap[i, :lens[i]]) for i, (hd, _) in enumerate(reads))
        assert got == open(os.path.join(d, "hits_s%d.txt" % s)).read()
        ba.free(); be.free()
    a.close()
    empty.close()


@pytest.mark.parametrize("n_a,n_b", [(1, 1), (31, 2), (32, 40), (33, 100), (70, 31), (6, 8), (0, 5), (5, 0)])
def test_index_merge_appends_columns_and_folds_bloom(mk, n_a, n_b):
    """mk_index_merge (Miekki::merge_indexes, Miekki.cpp:901-910) on random fingerprint matrices:
    the merged rows are [A | B] for every alignment of B inside the 32-genome plane groups,
    statistics are concatenated, a Bloom byte is A's where A set it, else B's; B is unchanged;
    and the merged index scores reads like the oracle loaded with the same payload."""
    rng = np.random.default_rng(100 * n_a + n_b)
    k, h = 31, 9
    B = 1 << h
    o = orc.Oracle(k=k, h=h, cap=max(1, n_a + n_b))
    W = len(o.bloom)

    def part(n):
        rows = rng.integers(0, 256, (B, n), dtype=np.uint8)
        ss = rng.integers(1, B + 1, n).astype(np.uint32)
        gs = rng.integers(0, 5_000_000, n).astype(np.uint64)
        bloom = np.where(rng.random(W) < 0.5, 1 << rng.integers(0, 8, W), 0).astype(np.uint8)
        return rows, gs, bloom, ss

    pa, pb = part(n_a), part(n_b)
    a, b = mk.Miekki(k=k, h=h, threshold=0), mk.Miekki(k=k, h=h, threshold=0)
    a.import_(*pa)
    b.import_(*pb)
    a.merge(b)
    assert a.n == n_a + n_b and b.n == n_b
    e = a.export()
    rows = np.hstack([pa[0], pb[0]])
    gs, ss = np.concatenate([pa[1], pb[1]]), np.concatenate([pa[3], pb[3]])
    bloom = np.where(pa[2] != 0, pa[2], pb[2])
    assert np.array_equal(e["rows"], rows)
    assert np.array_equal(e["genome_size"], gs) and np.array_equal(e["sketch_size"], ss)
    assert np.array_equal(e["bloom"][:W], bloom)
    eb = b.export()
    assert np.array_equal(eb["rows"], pb[0]) and np.array_equal(eb["bloom"][:W], pb[2])
    o.load(rows, gs, bloom, ss)
    reads = [rand_seq(rng, n) for n in (400, 1500)]
    counts, surv = a.query_counts(reads)
    hits = a.query(reads, 10, 1, 0.0)
    for i, s in enumerate(reads):
        oc, oa = o.counts(s)
        assert surv[i] == oa and np.array_equal(counts[i], oc)
        oh = o.filter(oc, 10, 1, 0.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"]) and np.array_equal(hits[i]["matches"], oh["matches"])
    with pytest.raises(mk.MiekkiError):
        a.merge(a)
    other_h = mk.Miekki(k=k, h=h + 1, threshold=0)
    with pytest.raises(mk.MiekkiError):
        a.merge(other_h)
    for x in (a, b, other_h):
        x.close()


@pytest.mark.parametrize("N,h", [(2100, 10), (40, 9), (1025, 5), (3300, 12)])
def test_tiled_scan_equals_oracle_on_random_index(mk, monkeypatch, N, h):
    """The tiled scan (scan_tiled.cu: TMA-staged rows shared by a tile of reads, lists sorted by
    bucket) forced on with MIEKKI_SCAN_TILED=1, against the oracle's scoring of the same random
    fingerprint matrix: several genome tiles with ragged last groups, more reads than one read
    tile, lists from empty to thousands of entries, and the same answers as the ring kernel."""
    rng = np.random.default_rng(1000 * h + N)
    k = 31
    B = 1 << h
    rows = rng.integers(0, 256, (B, N), dtype=np.uint8)
    ss = rng.integers(1, B + 1, N).astype(np.uint32)
    gs = rng.integers(0, 5_000_000, N).astype(np.uint64)
    o = orc.Oracle(k=k, h=h, cap=N)
    bloom = np.ones(len(o.bloom), np.uint8)                 # every k-mer passes
    o.load(rows, gs, bloom, ss)
    ix = mk.Miekki(k=k, h=h, threshold=0)
    ix.import_(rows, gs, bloom, ss)
    lens = [60, 31, 32, 10, 200, 1000, 5000, 9000, 16_000] + [int(x) for x in rng.integers(40, 3000, 130)]
    reads = [rand_seq(rng, n) for n in lens]
    monkeypatch.setenv("MIEKKI_SCAN_TILED", "1")
    monkeypatch.setenv("MIEKKI_SCAN_NARROW_GROUPS", "0")
    st0 = ix.stats()
    counts, surv = ix.query_counts(reads)
    hits = ix.query(reads, 10, 1, 0.0)
    monkeypatch.setenv("MIEKKI_SCAN_TILED", "0")
    counts_ring, surv_ring = ix.query_counts(reads)
    assert np.array_equal(counts, counts_ring) and np.array_equal(surv, surv_ring)
    for i, s in enumerate(reads):
        if len(s) < k:
            assert surv[i] == 0 and not counts[i].any() and len(hits[i]) == 0
            continue
        oc, oa = o.counts(s)
        assert surv[i] == oa, (i, len(s))
        assert np.array_equal(counts[i], oc), (i, len(s))
        oh = o.filter(oc, 10, 1, 0.0)
        assert np.array_equal(hits[i]["genome"], oh["genome"]) and np.array_equal(hits[i]["matches"], oh["matches"])
    ix.close()


def test_tiled_scan_on_case_a_hit_lines(mk, case_a, monkeypatch):
    """Real sketches, real Bloom table, the reference binary's hit lines -- through the tiled scan."""
    d, ix, _ = case_a
    monkeypatch.setenv("MIEKKI_SCAN_TILED", "1")
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    for s in (200, 0, 5000):
        assert gpu_hit_lines(ix, reads, s, chunk=64) == open(os.path.join(d, "hits_s%d.txt" % s)).read()


@pytest.mark.parametrize("k,h", [(31, 12), (21, 10), (5, 6), (2, 3)])
def test_packed_upload_equals_oracle(mk, monkeypatch, k, h):
    """mk_index_add packs long genomes to 2 bits per base on the host (pack.cpp) and ships only the
    words that need the general encoder -- the k-1 prefix, the ragged last word, words with N /
    lower case / any other byte -- as raw bytes.  Forced on for short sequences here and compared
    with the oracle (rows, statistics, every Bloom byte) and with the character path."""
    rng = np.random.default_rng(31 * k + h)
    genomes = []
    for i in range(40):
        n = int(rng.integers(max(k, 1), 5000)) if i % 4 else int(rng.integers(k, k + 40))
        s = np.frombuffer(rand_seq(rng, n, special=(i % 3 == 0)), np.uint8).copy()
        if i % 7 == 0:
            s[: min(n, 5)] = ord("n")                       # a dirty prefix zeroes the whole prefix word
        if i % 11 == 0:
            s[:] = np.frombuffer(bytes(s).lower(), np.uint8)   # packs badly: goes over as characters
        if i == 13:
            s = rng.integers(0, 256, n, dtype=np.uint8)     # arbitrary bytes
        genomes.append(s.tobytes())
    o = orc.Oracle(k=k, h=h, cap=len(genomes))
    for s in genomes:
        o.insert(s)
    exports = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MIEKKI_PACKED_UPLOAD", mode)
        monkeypatch.setenv("MIEKKI_PACK_MIN_LEN", "1")
        ix = mk.Miekki(k=k, h=h, threshold=0)
        ix.insert_sequences(genomes[:17])
        ix.insert_sequences(genomes[17:])
        exports[mode] = ix.export()
        ix.close()
    for mode, e in exports.items():
        assert np.array_equal(e["rows"], o.rows), mode
        assert np.array_equal(e["sketch_size"], o.sketch_size) and np.array_equal(e["genome_size"], o.genome_size), mode
        m = min(len(e["bloom"]), len(o.bloom))
        assert np.array_equal(e["bloom"][:m], o.bloom[:m]), mode


@pytest.mark.parametrize("R", [2, 4, 7])
def test_pipelined_chain_over_shards_on_one_gpu(mk, R):
    """The data path of sharded.pipelined_query without NCCL: R contexts on one GPU hold contiguous
    shards (19 genomes or fewer each at R = 4: one 32-genome group), every batch is scanned with
    mk_scan_async (batch i + 1 before batch i's heap step) and the heap goes through the shards
    with mk_topk_slot on device buffers.  Same genomes and reads as tests/nccl_worker.py."""
    import torch
    from miekki_b200 import sharded, synth
    k, h, G, GL = 31, 12, 75, 60_000
    rng = np.random.default_rng(42)
    genomes = [synth.genome(g, GL) for g in range(G - 15)]
    for j in range(15):
        genomes.append(synth.substitute(np.frombuffer(genomes[3 * j], np.uint8), 0.01 + 0.002 * j, rng).tobytes())
    reads = [s for _, s in synth.sample_reads(genomes, 330, 1500, sub_rate=0.01, block=9)]
    o = orc.Oracle(k=k, h=h, cap=G)
    for s in genomes:
        o.insert(s)
    engines = []
    for r in range(R):
        first, count = sharded.shard_range(G, r, R)
        ix = mk.Miekki(k=k, h=h, threshold=200)
        ix.set_shard(first)
        ix.insert_sequences(genomes[first:first + count])
        engines.append(ix)
    merged = sharded.fold_bloom([torch.from_numpy(e.bloom_get()) for e in engines]).numpy()
    m = min(len(merged), len(o.bloom))
    assert np.array_equal(merged[:m], o.bloom[:m])
    for e in engines:
        e.bloom_set(merged)
    K, nb = 10, 3
    per = len(reads) // nb
    for thr in (200, 0):
        batches = [[e.upload(reads[b * per:(b + 1) * per]) for b in range(nb)] for e in engines]
        bufs = [(torch.zeros((per, K * 24), dtype=torch.uint8, device="cuda"),
                 torch.zeros(per, dtype=torch.int32, device="cuda")) for _ in range(2)]
        got, pending = {}, None

        def finish(i, slots):
            hp, ln = bufs[i & 1]
            for r, e in enumerate(engines):
                e.topk_slot_ptr(slots[r], hp.data_ptr(), ln.data_ptr(), K, 10, 0.5 * thr, chain_in=r > 0,
                                finalize=r == R - 1)
            got[i] = (hp.cpu().numpy().view(mk.HIT_DTYPE).reshape(per, K).copy(), ln.cpu().numpy().view(np.uint32).copy())
        for i in range(nb):
            slots = [e.scan_async(batches[r][i]) for r, e in enumerate(engines)]
            if pending is not None:
                finish(*pending)
            pending = (i, slots)
        finish(*pending)
        for b in range(nb):
            hh, ll = got[b]
            for j in range(per):
                oc, _ = o.counts(reads[b * per + j])
                want = o.filter(oc, K, 10, 0.5 * thr)
                assert orc.format_hit_line(">r", hh[j, :ll[j]]) == orc.format_hit_line(">r", want), (R, thr, b, j)
        for bs in batches:
            for bt in bs:
                bt.free()
    for e in engines:
        e.close()


def test_tiled_scan_flushes_counters_of_long_lists(mk, monkeypatch):
    """Lists longer than the tiled kernel's 14 counter planes can count (16,380 entries): the counts
    are written out in passes (store, then add).  -h 15, sequences of 20-70 kbp (lists of up to
    ~29,000 entries) against a random matrix in which one genome equals the query's own sketch, so
    that counts really exceed 16,383."""
    rng = np.random.default_rng(77)
    k, h, N = 31, 15, 1100
    B = 1 << h
    rows = rng.integers(0, 256, (B, N), dtype=np.uint8)
    reads = [rand_seq(rng, n) for n in (70_000, 20_000, 45_000, 900, 64_000)] + [rand_seq(rng, 3000) for _ in range(60)]
    for col, s in ((5, reads[0]), (700, reads[4])):          # a genome that is the query itself
        fp, _, _ = orc.sketch(s, k, h)
        rows[:, col] = fp
    ss = rng.integers(1, B + 1, N).astype(np.uint32)
    gs = rng.integers(0, 5_000_000, N).astype(np.uint64)
    o = orc.Oracle(k=k, h=h, cap=N)
    bloom = np.ones(len(o.bloom), np.uint8)
    o.load(rows, gs, bloom, ss)
    ix = mk.Miekki(k=k, h=h, threshold=0)
    ix.import_(rows, gs, bloom, ss)
    monkeypatch.setenv("MIEKKI_SCAN_TILED", "1")
    monkeypatch.setenv("MIEKKI_SCAN_NARROW_GROUPS", "0")
    counts, surv = ix.query_counts(reads)
    assert counts[0, 5] == surv[0] and surv[0] > 20_000 and counts[4, 700] == surv[4] > 16_384
    for i, s in enumerate(reads[:8]):
        oc, oa = o.counts(s)
        assert surv[i] == oa and np.array_equal(counts[i], oc), i
    monkeypatch.setenv("MIEKKI_SCAN_TILED", "0")
    c2, s2 = ix.query_counts(reads)
    assert np.array_equal(counts, c2) and np.array_equal(surv, s2)
    ix.close()
