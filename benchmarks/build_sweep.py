#!/usr/bin/env python
"""Index-build throughput sweep (BASELINE config 4 shape): synthetic 5 Mbp genomes at
-h 17/19/20 and -k 21/31.  For each point prints one JSON line with
  device_gbp_per_s   genomes generated in HBM (counter-based generator) -> index, CUDA-event time
                     of the build kernels (encode, sketch, resolve, Bloom commit, scatter)
  wall_gbp_per_s     same, wall clock of the mk_index_add_batch calls (launches, final sync; the
                     synthetic generator runs between the calls and is not counted)
  host_gbp_per_s     genomes in pageable host memory -> mk_index_add (pinned ring + H2D) -> index
Run under torchrun for several GPUs: every rank builds its own shard, rank 0 prints per-rank and
summed rates (the build has no exchange except the Bloom fold at the end)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miekki_b200  # noqa: E402
from miekki_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=1000)
ap.add_argument("--genome-len", type=int, default=5_000_000)
ap.add_argument("--host-genomes", type=int, default=64)
ap.add_argument("--batch", type=int, default=128, help="genomes per mk_index_add_batch call")
ap.add_argument("--hs", default="17,19,20")
ap.add_argument("--ks", default="21,31")
ap.add_argument("--trace", action="store_true", help="per-batch wall and device times on stderr")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

hosts = None
for h in map(int, a.hs.split(",")):
    for k in map(int, a.ks.split(",")):
        ix = miekki_b200.Miekki(k=k, h=h, device=local)
        ix.reserve(a.genomes)
        b = ix.synth(1, 0, 64, a.genome_len)
        ix.insert_batch(b)                       # warm-up: one full 64-genome chunk sizes all scratch
        b.free()
        ix.stats_reset()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        trace = []
        insert_wall = 0.0
        for g0 in range(64, a.genomes, a.batch):
            t1 = time.perf_counter()
            b = ix.synth(1, rank * a.genomes + g0, min(a.batch, a.genomes - g0), a.genome_len)
            t2 = time.perf_counter()
            ix.insert_batch(b)
            t3 = time.perf_counter()
            insert_wall += t3 - t2
            b.free()
            if a.trace:
                trace.append((g0, round((t2 - t1) * 1e3, 2), round((t3 - t2) * 1e3, 2), round(ix.stats()["sketch_ms"], 2)))
        if a.trace and rank == 0:
            print("g0, synth ms, insert ms, cumulative device ms:", file=sys.stderr)
            for row in trace[:12] + trace[-12:]:
                print(row, file=sys.stderr)
        wall = time.perf_counter() - t0
        st = ix.stats()
        bases = st["bases_sketched"]
        ix.close()
        # host path
        if hosts is None:
            hosts = [synth.cb_bases(1, g, 0, a.genome_len).tobytes() for g in range(8)]
        hx = miekki_b200.Miekki(k=k, h=h, device=local)
        hx.reserve(2 * a.host_genomes)
        seqs = [hosts[i % 8] for i in range(a.host_genomes)]
        hx.insert_sequences(seqs)
        t0 = time.perf_counter()
        hx.insert_sequences(seqs)
        hwall = time.perf_counter() - t0
        hx.close()
        rec = {"h": h, "k": k, "genomes_per_gpu": a.genomes, "n_gpus": world,
               "device_gbp_per_s": bases / st["sketch_ms"] / 1e6, "wall_gbp_per_s": bases / insert_wall / 1e9,
               "host_gbp_per_s": a.host_genomes * a.genome_len / hwall / 1e9}
        if dist:
            import torch
            t = torch.tensor([rec["device_gbp_per_s"], rec["wall_gbp_per_s"], rec["host_gbp_per_s"]],
                             dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            rec.update(sum_device_gbp_per_s=float(t[0]), sum_wall_gbp_per_s=float(t[1]),
                       sum_host_gbp_per_s=float(t[2]))
        if rank == 0:
            print(json.dumps(rec), flush=True)
if dist:
    dist.destroy_process_group()
