"""N > 1 host logic on CPU: world_size 2, gloo backend.  The shard protocol of
miekki_b200/sharded.py (contiguous id ranges, Bloom "lowest rank wins", chained bounded heap)
is driven with an oracle-backed engine standing in for the GPU, and must reproduce the golden
hit lines of the unsharded reference run exactly (ties included)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H


class OracleEngine:
    """Scores its shard with the CPU oracle; same topk_ptr contract as binding.Miekki."""

    def __init__(self, genomes, first_id, k, h, b, bloom=None):
        from oracle import oracle as orc
        self.orc = orc
        self.first_id = first_id
        self.o = orc.Oracle(k=k, h=h, b=b, cap=max(1, len(genomes)))
        for s in genomes:
            self.o.insert(s)
        self.counts = None

    def bloom_tensor(self):
        return torch.from_numpy(np.array(self.o.bloom))

    def set_bloom(self, t):
        self.o.bloom[:] = t.numpy()

    def scan(self, reads):
        self.counts = [self.o.counts(s)[0] for s in reads]

    # the pipelined pair of binding.Miekki: two count tiles, alternating
    def sketch_async(self, reads):
        self.sketched_ahead = reads            # mk_sketch_async: the batch the next scan_async gets

    def scan_async(self, reads):
        ahead = getattr(self, "sketched_ahead", None)
        assert ahead is None or ahead is reads, "sketch_async was given another batch than the next scan"
        self.sketched_ahead = None
        self.slot = getattr(self, "slot", 0) ^ 1
        if not hasattr(self, "tiles"):
            self.tiles = {}
        self.tiles[self.slot] = [self.o.counts(s)[0] for s in reads]
        return self.slot

    def topk_slot_ptr(self, slot, heap_ptr, len_ptr, nresults, min_score, min_intersection, chain_in, finalize,
                      first=0, count=None):
        # a range of the batch's reads; the pointers address read `first` (mk_topk_slot_range)
        self.counts = self.tiles[slot][first:None if count is None else first + count]
        self.topk_ptr(heap_ptr, len_ptr, nresults, min_score, min_intersection, chain_in, finalize)

    def topk_ptr(self, heap_ptr, len_ptr, nresults, min_score, min_intersection, chain_in, finalize):
        import ctypes as C
        n = len(self.counts)
        heap = np.ctypeslib.as_array(C.cast(heap_ptr, C.POINTER(C.c_uint8)), shape=(n, nresults * 24))
        lens = np.ctypeslib.as_array(C.cast(len_ptr, C.POINTER(C.c_int32)), shape=(n,))
        for i in range(n):
            row = heap[i].view(self.orc.HIT_DTYPE)
            length = int(lens[i]) if chain_in else 0
            lens[i] = self.orc.filter_chain(self.counts[i], self.first_id, self.o.sketch_size,
                                            self.o.genome_size, nresults, min_score, min_intersection,
                                            row, length, finalize)


def _worker(rank, world, port, case, k, h, b, thresholds, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from miekki_b200 import sharded
        from oracle import oracle as orc
        d = os.path.join(H.GOLDEN, case)
        genomes = [H.genome_like_reference(os.path.join(d, f)) for f in H.load_list(d)]
        first, count = sharded.shard_range(len(genomes), rank, world)
        eng = OracleEngine(genomes[first:first + count], first, k, h, b)
        eng.set_bloom(sharded.merge_bloom(eng.bloom_tensor()))
        reads = H.reads_like_reference(os.path.join(d, "reads.fa"), k)
        eng.scan([s for _, s in reads])
        out = {}
        for s in thresholds:
            heap = torch.zeros((len(reads), 10 * 24), dtype=torch.uint8)
            lens = torch.zeros(len(reads), dtype=torch.int32)
            sharded.chained_topk(eng, heap, lens, 10, 10, 0.5 * s)
            if rank == world - 1:
                hn = heap.numpy()
                out[s] = "".join(orc.format_hit_line(hd, hn[i].view(orc.HIT_DTYPE)[: int(lens[i])])
                                 for i, (hd, _) in enumerate(reads))
        if rank == world - 1:
            q.put((out, np.array(eng.o.bloom)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _pipelined_worker(rank, world, port, case, k, h, b, n_batches, q):
    """sharded.pipelined_query over n_batches distinct read batches (scan of batch i + 1 enqueued
    before batch i's heap is chained; consecutive batches in alternating buffers)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from miekki_b200 import sharded
        from oracle import oracle as orc
        d = os.path.join(H.GOLDEN, case)
        genomes = [H.genome_like_reference(os.path.join(d, f)) for f in H.load_list(d)]
        first, count = sharded.shard_range(len(genomes), rank, world)
        eng = OracleEngine(genomes[first:first + count], first, k, h, b)
        eng.set_bloom(sharded.merge_bloom(eng.bloom_tensor()))
        reads = H.reads_like_reference(os.path.join(d, "reads.fa"), k)
        per = len(reads) // n_batches
        parts = [reads[i * per:(i + 1) * per] for i in range(n_batches)]
        heap = torch.zeros((per, 10 * 24), dtype=torch.uint8)
        lens = torch.zeros(per, dtype=torch.int32)
        lines = {}

        def on_result(i):
            hn = heap.numpy()
            lines[i] = "".join(orc.format_hit_line(hd, hn[j].view(orc.HIT_DTYPE)[: int(lens[j])])
                               for j, (hd, _) in enumerate(parts[i]))
        # the last batch travels in 3 tiles of reads (5 + 4 + 5), the others in 2
        sharded.pipelined_query(eng, ([s for _, s in p] for p in parts), heap, lens, 10, 10, 0.0, on_result=on_result,
                                tiles=2, last_tiles=3)
        if rank == world - 1:
            q.put("".join(lines[i] for i in range(n_batches)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_three_rank_pipelined_batches_reproduce_reference_lines():
    """-s 0 (every candidate competes: heap order decides), 3 ranks, 4 batches of 14 reads."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipelined_worker, args=(r, 3, port, "caseA", 31, 12, 33, 4, q)) for r in range(3)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = open(os.path.join(H.GOLDEN, "caseA", "hits_s0.txt")).read().split("\n")
    n = (len(want) - 1) // 4 * 4
    assert out.split("\n")[:n] == want[:n] and n >= 52


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_shards_reproduce_reference_lines():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "caseA", 31, 12, 33, (200, 0), q)) for r in range(2)]
    for p in procs:
        p.start()
    out, bloom = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    d = os.path.join(H.GOLDEN, "caseA")
    for s in (200, 0):
        assert out[s] == open(os.path.join(d, "hits_s%d.txt" % s)).read()
    # the merged Bloom table equals the single in-order build's (reference dump)
    z = H.load_dump_npz(os.path.join(d, "dump.npz"))
    idx, val = H.bloom_nonzero(bloom)
    assert np.array_equal(idx, z["bloom_idx"]) and np.array_equal(val, z["bloom_val"])


def test_shard_ranges_cover_ids_in_order():
    from miekki_b200 import sharded
    for n in (1, 7, 14, 100_000):
        for world in (1, 2, 3, 8):
            nxt = 0
            for r in range(world):
                first, count = sharded.shard_range(n, r, world)
                assert first == nxt
                nxt += count
            assert nxt == n


def test_fold_bloom_lowest_rank_wins():
    from miekki_b200 import sharded
    a = torch.tensor([0, 4, 0, 1], dtype=torch.uint8)
    b = torch.tensor([2, 8, 0, 0], dtype=torch.uint8)
    c = torch.tensor([16, 16, 16, 16], dtype=torch.uint8)
    assert sharded.fold_bloom([a, b, c]).tolist() == [2, 4, 16, 1]


def test_drain_tiles_and_nccl_defaults(monkeypatch):
    from miekki_b200 import sharded
    assert sharded.drain_tiles(100_000, 8) == 16          # 2 x world tiles ...
    assert sharded.drain_tiles(20_000, 8) == 16
    assert sharded.drain_tiles(5_000, 8) == 4             # ... of at least 1,024 reads
    assert sharded.drain_tiles(14, 3) == 1
    monkeypatch.delenv("NCCL_MAX_P2P_NCHANNELS", raising=False)
    sharded.nccl_env_defaults()
    assert os.environ["NCCL_MAX_P2P_NCHANNELS"] == "1"
    monkeypatch.setenv("NCCL_MAX_P2P_NCHANNELS", "4")     # the user's own setting wins
    sharded.nccl_env_defaults()
    assert os.environ["NCCL_MAX_P2P_NCHANNELS"] == "4"


def test_pipelined_query_without_a_process_group():
    """One shard, no torch.distributed: the streaming path bench.py's end-to-end leg takes on one GPU.
    Three batches (the second and third travel through the second buffer pair / come back through
    on_result in the caller's tensors), hit lines equal the reference's."""
    from miekki_b200 import sharded
    from oracle import oracle as orc
    d = os.path.join(H.GOLDEN, "caseA")
    genomes = [H.genome_like_reference(os.path.join(d, f)) for f in H.load_list(d)]
    eng = OracleEngine(genomes, 0, 31, 12, 33)
    reads = H.reads_like_reference(os.path.join(d, "reads.fa"), 31)
    per = len(reads) // 3
    parts = [reads[i * per:(i + 1) * per] for i in range(3)]
    heap = torch.zeros((per, 10 * 24), dtype=torch.uint8)
    lens = torch.zeros(per, dtype=torch.int32)
    lines, freed = {}, []

    def on_result(i):
        hn = heap.numpy()
        lines[i] = "".join(orc.format_hit_line(hd, hn[j].view(orc.HIT_DTYPE)[: int(lens[j])])
                           for j, (hd, _) in enumerate(parts[i]))
    sharded.pipelined_query(eng, ([s for _, s in p] for p in parts), heap, lens, 10, 10, 100.0,
                            on_result=on_result, after_chain=freed.append, tiles=2, last_tiles=3)
    assert freed == [0, 1, 2]
    want = open(os.path.join(d, "hits_s200.txt")).read().split("\n")
    got = "".join(lines[i] for i in range(3)).split("\n")
    assert got[:3 * per] == want[:3 * per] and 3 * per >= 54
    # without on_result the last batch's lists are left in the caller's tensors
    heap.zero_()
    lens.zero_()
    sharded.pipelined_query(eng, ([s for _, s in p] for p in parts[:2]), heap, lens, 10, 10, 100.0)
    hn = heap.numpy()
    last = "".join(orc.format_hit_line(hd, hn[j].view(orc.HIT_DTYPE)[: int(lens[j])]) for j, (hd, _) in enumerate(parts[1]))
    assert last.split("\n")[:per] == want[per:2 * per]
