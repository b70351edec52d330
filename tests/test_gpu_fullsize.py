"""Parity at BASELINE.json's full sizes (config 2: 10,000 x 5 Mbp genomes, -h 20, 1 kbp reads)
through checks that do not need the CPU to redo the whole job:

* columns of the index for sampled genomes == the oracle's sketch of the same genome
  (regenerated on the host with the counter-based generator), incl. sketch_size/genome_size;
* counts of sampled reads == numpy recomputation from the EXPORTED matrix and the oracle's
  Bloom-masked read sketch (checks scan + layout at 10,000 columns / 313 groups);
* hit lists == the oracle's filter on those counts; error-free reads rank their source first;
* linearity: scoring against two half-indexes and chaining == scoring against the whole;
* idempotence: the same batch twice gives identical bytes.

MIEKKI_FULLSIZE_GENOMES overrides the genome count (default 10000).
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

N = int(os.environ.get("MIEKKI_FULLSIZE_GENOMES", "10000"))
L, K, H = 5_000_000, 31, 20
SEED = 0x5EED_B200


@pytest.fixture(scope="module")
def big():
    import miekki_b200
    from miekki_b200 import synth
    ix = miekki_b200.Miekki(k=K, h=H, threshold=200)
    ix.reserve(N)
    for g0 in range(0, N, 32):
        b = ix.synth(SEED, g0, min(32, N - g0), L)
        ix.insert_batch(b)
        b.free()
    reads, gs, ps = synth.cb_reads(SEED, N, L, 64, 1000, 0.0, block=99)
    yield miekki_b200, ix, [r.tobytes() for r in reads], gs, ps
    ix.close()


def test_sampled_columns_equal_oracle_sketch(big):
    from miekki_b200 import synth
    mk, ix, _, _, _ = big
    e = ix.export()
    ss, gsz = e["sketch_size"], e["genome_size"]
    for g in (0, 1, 31, 32, N // 2 + 5, N - 1):
        seq = synth.cb_bases(SEED, g, 0, L).tobytes()
        fp, anc, act = orc.sketch(seq, K, H)
        assert np.array_equal(e["rows"][:, g], fp), g
        assert ss[g] == act
        assert gsz[g] == L                      # unsaturated sketch: genome_size == length
    big_rows = e["rows"]
    test_sampled_columns_equal_oracle_sketch.rows = big_rows
    test_sampled_columns_equal_oracle_sketch.bloom = e["bloom"]
    test_sampled_columns_equal_oracle_sketch.ss = ss
    test_sampled_columns_equal_oracle_sketch.gs = gsz


def test_counts_and_hits_equal_numpy_on_exported_matrix(big):
    mk, ix, reads, src, _ = big
    rows = getattr(test_sampled_columns_equal_oracle_sketch, "rows", None)
    if rows is None:
        e = ix.export()
        rows, bloom, ss, gsz = e["rows"], e["bloom"], e["sketch_size"], e["genome_size"]
    else:
        f = test_sampled_columns_equal_oracle_sketch
        bloom, ss, gsz = f.bloom, f.ss, f.gs
    counts, surv = ix.query_counts(reads)
    hits = ix.query(reads, 10, 10, 100.0)
    L_ = orc.lib()
    top_is_source = 0
    for i, s in enumerate(reads):
        fp, anc, _ = orc.sketch(s, K, H)
        act = np.flatnonzero(fp != 255)
        keep = [b for b in act if L_.mko_bloom_check(bloom.ctypes.data, 33, int(anc[b]))]
        assert surv[i] == len(keep)
        want = np.zeros(N, np.uint32)
        for b in keep:
            want += rows[b] == fp[b]
        assert np.array_equal(counts[i], want), i
        hits_o = np.zeros(10, orc.HIT_DTYPE)
        n = L_.mko_filter(want.ctypes.data, N, np.ascontiguousarray(ss).ctypes.data,
                          np.ascontiguousarray(gsz).ctypes.data, 10, 10, 100.0, hits_o.ctypes.data)
        assert np.array_equal(hits[i]["genome"], hits_o[:n]["genome"])
        assert np.array_equal(hits[i]["matches"], hits_o[:n]["matches"])
        np.testing.assert_allclose(hits[i]["intersection"], hits_o[:n]["intersection"], rtol=1e-6)
        top_is_source += int(len(hits[i]) > 0 and hits[i]["genome"][0] == src[i])
    assert top_is_source >= len(reads) - 2      # error-free reads find their genome


def test_linearity_over_shards_and_idempotence(big):
    mk, ix, reads, _, _ = big
    e = ix.export()
    half = (N // 2) // 32 * 32 + 7 if N > 64 else N // 2     # deliberately not group aligned
    a = mk.Miekki(k=K, h=H, threshold=200)
    b = mk.Miekki(k=K, h=H, threshold=200)
    a.import_(e["rows"][:, :half], e["genome_size"][:half], e["bloom"], e["sketch_size"][:half])
    b.import_(e["rows"][:, half:], e["genome_size"][half:], e["bloom"], e["sketch_size"][half:])
    b.set_shard(half)
    ca, _ = a.query_counts(reads)
    cb, _ = b.query_counts(reads)
    cw, _ = ix.query_counts(reads)
    assert np.array_equal(np.concatenate([ca, cb], axis=1), cw)
    heap = np.zeros((len(reads), 10), mk.HIT_DTYPE)
    lens = np.zeros(len(reads), np.uint32)
    ba, bb, bw = a.upload(reads), b.upload(reads), ix.upload(reads)
    a.query_chain(ba, heap, lens, 10, 10, 100.0, finalize=False)
    b.query_chain(bb, heap, lens, 10, 10, 100.0, finalize=True)
    h1, n1 = ix.query_batch(bw, 10, 10, 100.0)
    h2, n2 = ix.query_batch(bw, 10, 10, 100.0)
    assert np.array_equal(n1, n2) and h1.tobytes() == h2.tobytes()
    assert np.array_equal(n1, lens)
    for i in range(len(reads)):
        assert heap[i, :lens[i]].tobytes() == h1[i, :n1[i]].tobytes()
    for x in (ba, bb, bw):
        x.free()
    a.close()
    b.close()
