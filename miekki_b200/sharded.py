"""Genome-sharded runs: one process per GPU, `torch.distributed` for the plumbing.

The index partitions by genome (matrix columns), so a query needs exactly two exchanges
 and no index byte ever crosses NVLink:

* after the build, the global Bloom filter: byte-wise "lowest rank with a non-zero byte wins"
  (exact because each byte belongs to the smallest (genome id, bucket, probe) that maps to
  it and shards are contiguous ascending id ranges);
* after each rank has scanned its own columns, the bounded heap of Miekki.cpp:386-393 is
  passed through the ranks in ascending id order (24 bytes x nresults per read), which
  reproduces the sequential filter exactly, ties included; the last rank applies sort_heap.

The functions are backend-agnostic: tensors live on the GPU with NCCL, or on the CPU with gloo
(tests/test_sharded_cpu.py drives them with world_size 2 and an oracle-backed engine).
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist

HIT_BYTES = 24          # sizeof(mk_hit)


def nccl_env_defaults() -> None:
    """Call before `init_process_group("nccl")`.  A receive posted for the next heap is a kernel
    that spins on its SMs until the previous rank sends, which may be a whole scan later; with
    NCCL's default of many P2P channels it takes that many SMs away from the scan meanwhile
    (config 3 on 2 GPUs: 665 ms per step instead of 604).  One channel moves a 24 MB heap in
    about a millisecond, which is all the chain needs."""
    os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "1")


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous ascending ids [first, first + count) of `rank`: [r N / R, (r+1) N / R)."""
    first = rank * n_total // world
    return first, (rank + 1) * n_total // world - first


def fold_bloom(tables: list[torch.Tensor]) -> torch.Tensor:
    """tables[r] = Bloom bytes of rank r; the lowest rank's non-zero byte wins."""
    merged = tables[0].clone()
    for t in tables[1:]:
        merged = torch.where(merged != 0, merged, t)
    return merged


def merge_bloom(local: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks end with the Bloom table a single in-order build would have produced."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    tables = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(tables, local, group=group)
    return fold_bloom(tables)


def _exchange(op, tensors, peer, group) -> None:
    """One batched P2P call for the tensors of a tile (a single NCCL group launch; unbatched
    sends on the default group are serialised with every other collective)."""
    for w in dist.batch_isend_irecv([dist.P2POp(op, t, peer, group) for t in tensors]):
        w.wait()                  # NCCL: the current stream waits; gloo: the host does


def drain_tiles(n_reads: int, world: int) -> int:
    """How many read tiles the last batch of a run is chained in: enough that a tile's heap step
    is a small fraction of the batch's, never tiles of fewer than 1,024 reads."""
    return max(1, min(2 * world, n_reads // 1024))


def chained_topk(engine, heap: torch.Tensor, lens: torch.Tensor, nresults: int, min_score: int,
                 min_intersection: float, group=None, slot: int | None = None, tiles: int = 1,
                 trace: list | None = None) -> None:
    """heap: uint8 [n_reads, nresults * 24], lens: int32 [n_reads], on the engine's device.
    `engine.topk_ptr(heap_ptr, len_ptr, nresults, min_score, min_intersection, chain_in,
    finalize)` applies this rank's stored counts to the heap state in place
    (`engine.topk_slot_ptr(slot, ...)` when `slot` is given: the pipelined form, where the
    caller has already enqueued the next batch's scan with `scan_async`).
    After the call the LAST rank holds the final hit lists.

    tiles > 1 (needs `slot`): the batch travels in that many tiles of reads, each stepped with
    `topk_slot_ptr(..., first=, count=)` and sent on at once, so rank r steps tile t while rank
    r + 1 steps tile t - 1.  Reads are independent in the heap step, so the result is the same;
    the chain then takes one batch-sized heap step plus (world - 1) tile-sized ones instead of
    `world` batch-sized ones.  Worth it when the GPUs have nothing else to do (the last batch of
    a run); beside a running scan every tile would wait for SMs of its own.

    On GPUs the receive is followed by a wait on the current stream only, so when this runs
    under `torch.cuda.stream(side_stream)` the main stream keeps scanning."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = heap.shape[0]
    tiles = max(1, min(tiles, n)) if slot is not None and world > 1 else 1
    kw = dict(chain_in=rank > 0, finalize=rank == world - 1)
    for t in range(tiles):
        a, b = t * n // tiles, (t + 1) * n // tiles
        h, l = heap[a:b], lens[a:b]
        if rank > 0:
            _exchange(dist.irecv, (h, l), rank - 1, group)
            if heap.is_cuda:
                torch.cuda.current_stream().synchronize()      # the engine reads the buffers next
        if trace is not None:
            trace.append(("received", t, time.perf_counter()))
        if slot is None:
            engine.topk_ptr(heap.data_ptr(), lens.data_ptr(), nresults, min_score, min_intersection, **kw)
        elif tiles == 1:
            engine.topk_slot_ptr(slot, heap.data_ptr(), lens.data_ptr(), nresults, min_score,
                                 min_intersection, **kw)
        else:
            engine.topk_slot_ptr(slot, h.data_ptr(), l.data_ptr(), nresults, min_score, min_intersection,
                                 first=a, count=b - a, **kw)
        if trace is not None:
            trace.append(("stepped", t, time.perf_counter()))
        if rank < world - 1:
            _exchange(dist.isend, (h, l), rank + 1, group)


def pipelined_query(engine, batches, heap: torch.Tensor, lens: torch.Tensor, nresults: int,
                    min_score: int, min_intersection: float, on_result=None, after_chain=None,
                    group=None, last_tiles: int | None = None, tiles: int | None = None) -> None:
    """Software pipeline over read batches: the scan of batch i+1 is enqueued before batch i's
    heap is chained through the ranks, so the cheap, latency-bound chain hides behind the scan.
    `batches` yields engine batches; `on_result(i)` is called on the last rank when batch i's hit
    lists are final in (heap, lens); `after_chain(i)` on every rank once its own step for batch i
    is done (its scan has run by then, so the batch may be freed).

    Consecutive batches travel in alternating buffers: an NCCL send only completes when the next
    rank has posted its receive, which may be a whole scan later, so the buffer a batch was sent
    from must not be written again (by the next batch's top-k on the first rank) before that.
    A buffer is reused two batches later, after the event recorded behind its sends.

    Every batch is chained in `tiles` tiles of reads (default: one per rank, tiles of >= 1,024
    reads), so that a rank steps tile t while the next one steps tile t - 1 and the chain takes
    about two heap steps instead of `world`; the last batch, which has no scan to hide behind, in
    `last_tiles` (default `drain_tiles`).  See chained_topk.

    On GPUs the engine should leave one SM out of the scan's grid (`set_scan_spare_sms(1)`): the
    NCCL kernels of the chain do not fit beside a persistent scan CTA and would otherwise wait
    for the scan of the next batch to end, and the rank waiting for the heap with them."""
    side = torch.cuda.Stream() if heap.is_cuda else None
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bufs = [(heap, lens), (torch.empty_like(heap), torch.empty_like(lens))]
    sent = [None, None]
    # MIEKKI_CHAIN_TRACE=<prefix>: host time stamps of every batch's steps on this rank, one JSON line
    # per call appended to <prefix>.rank<r> (when did the heap arrive, when was it stepped)
    trace_to = os.environ.get("MIEKKI_CHAIN_TRACE")
    events = [] if trace_to else None
    t_origin = time.perf_counter()

    def finish(i, slot, tiles=1):
        h, l = bufs[i & 1]

        def body():
            if sent[i & 1] is not None:
                sent[i & 1].synchronize()           # the sends of batch i - 2 have left this buffer
            tr = [] if events is not None else None
            chained_topk(engine, h, l, nresults, min_score, min_intersection, group, slot, tiles, tr)
            if events is not None:
                events.append({"batch": i, "steps": [(w, t, round(1e3 * (x - t_origin), 2)) for w, t, x in tr]})
            if side is not None and rank < world - 1:
                sent[i & 1] = torch.cuda.Event()
                sent[i & 1].record(torch.cuda.current_stream())
            if on_result is not None and rank == world - 1:
                if h is not heap:                   # results are handed over in the caller's tensors
                    heap.copy_(h)
                    lens.copy_(l)
                    if side is not None:
                        torch.cuda.current_stream().synchronize()
                on_result(i)
            if after_chain is not None:
                after_chain(i)
        if side is not None:
            with torch.cuda.stream(side):
                body()
        else:
            body()

    pending = None
    it = iter(batches)
    nxt = next(it, None)
    sketch_ahead = getattr(engine, "sketch_async", None)
    i = 0
    while nxt is not None:
        slot = engine.scan_async(nxt)
        if events is not None:
            events.append({"batch": i, "scan_enqueued": round(1e3 * (time.perf_counter() - t_origin), 2)})
        # The next batch is sketched now, beside the scan just enqueued: the chain step below may
        # return only when that scan has ended (a later rank waits for the previous rank's heap),
        # and a sketch enqueued then would run with the GPU otherwise idle.
        nxt = next(it, None)
        if nxt is not None and sketch_ahead is not None:
            sketch_ahead(nxt)
        if pending is not None:
            finish(*pending, tiles=max(1, min(world, heap.shape[0] // 1024)) if tiles is None else tiles)
        pending = (i, slot)
        i += 1
    if pending is not None:
        finish(*pending, tiles=drain_tiles(heap.shape[0], world) if last_tiles is None else last_tiles)
        if on_result is None and rank == world - 1 and (pending[0] & 1):
            # the last batch went through the second buffer: its hit lists belong in the caller's
            if side is not None:
                with torch.cuda.stream(side):
                    heap.copy_(bufs[1][0])
                    lens.copy_(bufs[1][1])
            else:
                heap.copy_(bufs[1][0])
                lens.copy_(bufs[1][1])
    if side is not None:
        side.synchronize()
    if events is not None:
        with open("%s.rank%d" % (trace_to, rank), "a") as f:
            f.write(json.dumps({"total_ms": round(1e3 * (time.perf_counter() - t_origin), 2), "events": events}) + "\n")
