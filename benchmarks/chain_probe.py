#!/usr/bin/env python
"""Where one shard's step goes in the pipelined multi-GPU path, without a second GPU: N genomes at
-h 17 (default 50,000 = one of two ranks of BASELINE config 3), R reads of 10 kbp with 5 %
substitutions at -s 0, two batches through mk_scan_async / mk_topk_slot exactly as
sharded.pipelined_query issues them on one rank.  Host wall clock per call, once as the first rank
of a chain (chain_in = 0) and once as a later rank (chain_in = 1, heap state of a previous shard),
whole batch and in tiles of reads; then PROBE_STEPS batches in steady state, pipelined and through
mk_query_batch.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import miekki_b200  # noqa: E402
from miekki_b200 import synth  # noqa: E402

SEED = 0x5EED_B200


def main():
    n = int(os.environ.get("PROBE_GENOMES", "50000"))
    r = int(os.environ.get("PROBE_READS", "20000"))
    h = int(os.environ.get("PROBE_H", "17"))
    rl = int(os.environ.get("PROBE_READ_LEN", "10000"))
    K = 10
    ix = miekki_b200.Miekki(k=31, h=h, threshold=200)
    ix.reserve(n)
    for g0 in range(0, n, 128):
        b = ix.synth(SEED, g0, min(128, n - g0), 5_000_000)
        ix.insert_batch(b)
        b.free()
    reads, _, _ = synth.cb_reads_block(SEED, n, 5_000_000, r, rl, 0.05, block=1)
    seqs = [x.tobytes() for x in reads]
    batches = [ix.upload(seqs), ix.upload(seqs[::-1])]
    heap = np.zeros((r, K), miekki_b200.HIT_DTYPE)
    lens = np.zeros(r, np.uint32)
    out = {"genomes": n, "reads": r, "h": h, "read_len": rl}

    def t(f):
        t0 = time.perf_counter()
        f()
        return round(1e3 * (time.perf_counter() - t0), 2)

    def two_steps(chain_in, tiles, tag):
        """scan(b0); scan(b1); topk(b0); topk(b1): the order pipelined_query issues them in"""
        ix.sync()
        rec = {}
        t_all = time.perf_counter()
        s0 = ix.scan_async(batches[0])
        s1 = ix.scan_async(batches[1])
        rec["enqueue_ms"] = round(1e3 * (time.perf_counter() - t_all), 2)
        for name, slot in (("topk_b0_ms", s0), ("topk_b1_ms", s1)):
            def step():
                for k in range(tiles):
                    a, b = k * r // tiles, (k + 1) * r // tiles
                    ix.topk_slot_ptr(slot, heap[a:].ctypes.data, lens[a:].ctypes.data, K, 10, 0.0,
                                     chain_in=chain_in, finalize=True, first=a, count=b - a)
            rec[name] = t(step)              # b0: includes waiting for its scan; b1: likewise
        ix.sync()
        rec["total_ms"] = round(1e3 * (time.perf_counter() - t_all), 2)
        out[tag] = rec

    two_steps(False, 1, "warmup")
    ix.stats_reset()
    two_steps(False, 1, "first_rank")
    st = ix.stats()
    out["scan_ms_per_batch"] = st["scan_ms"] / 2
    out["read_sketch_ms_per_batch"] = st["read_sketch_ms"] / 2
    out["topk_ms_per_batch_device"] = st["topk_ms"] / 2
    # heap state as a previous shard would hand it over: this shard's own unsorted heaps under other ids
    s = ix.scan_async(batches[0])
    ix.topk_slot_ptr(s, heap.ctypes.data, lens.ctypes.data, K, 10, 0.0, chain_in=False, finalize=False)
    ix.sync()
    ix.set_shard(n)
    two_steps(True, 1, "later_rank")
    two_steps(True, 16, "later_rank_16_tiles")
    # the heap step alone, GPU otherwise idle
    s = ix.scan_async(batches[0])
    ix.sync()
    out["topk_alone_ms"] = t(lambda: ix.topk_slot_ptr(s, heap.ctypes.data, lens.ctypes.data, K, 10, 0.0,
                                                      chain_in=True, finalize=True))
    # steady state: PROBE_STEPS batches through the pipelined pair (scan of batch i + 1 enqueued before
    # batch i's heap step, as sharded.pipelined_query does) against the same batches through mk_query_batch
    steps = int(os.environ.get("PROBE_STEPS", "6"))
    min_int = float(os.environ.get("PROBE_MIN_INT", "0"))

    def pipelined():
        pending = None
        for i in range(steps):
            s_ = ix.scan_async(batches[i & 1])
            if pending is not None:
                ix.topk_slot_ptr(pending, heap.ctypes.data, lens.ctypes.data, K, 10, min_int, chain_in=False, finalize=True)
            pending = s_
        ix.topk_slot_ptr(pending, heap.ctypes.data, lens.ctypes.data, K, 10, min_int, chain_in=False, finalize=True)
        ix.sync()

    def plain():
        for i in range(steps):
            ix.query_batch(batches[i & 1], K, 10, min_int, fetch=False)
        ix.sync()

    ix.set_shard(0)
    for name, f in (("pipelined", pipelined), ("query_batch", plain)):
        f()
        ix.stats_reset()
        ms = t(f)
        st = ix.stats()
        out["steady_" + name] = {"ms_per_step": round(ms / steps, 2), "scan_ms_per_step": round(st["scan_ms"] / steps, 2),
                                 "read_sketch_ms_per_step": round(st["read_sketch_ms"] / steps, 2),
                                 "topk_ms_per_step": round(st["topk_ms"] / steps, 2)}
    print(json.dumps(out))
    ix.close()


if __name__ == "__main__":
    main()
