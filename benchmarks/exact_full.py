#!/usr/bin/env python
"""BASELINE.json config 5 at full size on one B200: exact mode (-e) over 10,000 indexed 5 Mbp
genomes and 100,000 error-free 1 kbp reads.

Phases, like Miekki::query_file_exact (Miekki.cpp:723-759) + ground_truth_batch (:792-859):
  1. approximate query with the -e parameters (5 results, min score 10, intersection >= -s);
  2. every hit is a (read, genome) candidate; candidates are grouped by genome;
  3. per genome with candidates: set B of its k-mers is built on the device and each candidate
     read's set A is intersected with it (mk_exact_batch): nb_inter, nb_union, real Jaccard.
Genomes are regenerated in HBM by the counter-based generator (50 Gbp do not fit host memory
comfortably); the reads travel from host memory.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miekki_b200  # noqa: E402
from miekki_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=10_000)
ap.add_argument("--genome-len", type=int, default=5_000_000)
ap.add_argument("--reads", type=int, default=100_000)
ap.add_argument("--read-len", type=int, default=1000)
ap.add_argument("--h", type=int, default=20)
ap.add_argument("--k", type=int, default=31)
ap.add_argument("--threshold", type=int, default=200)
a = ap.parse_args()
SEED = 1

ix = miekki_b200.Miekki(k=a.k, h=a.h, threshold=a.threshold)
ix.reserve(a.genomes)
t0 = time.perf_counter()
for g0 in range(0, a.genomes, 128):
    b = ix.synth(SEED, g0, min(128, a.genomes - g0), a.genome_len)
    ix.insert_batch(b)
    b.free()
build_s = time.perf_counter() - t0

stride = (a.read_len + 15) // 16 * 16
data = np.zeros((a.reads, stride), np.uint8)
src = np.zeros(a.reads, np.int64)
for b0 in range(0, a.reads, 10_000):
    m = min(10_000, a.reads - b0)
    r, gs, _ = synth.cb_reads(SEED, a.genomes, a.genome_len, m, a.read_len, 0.0, block=b0 // 10_000)
    data[b0:b0 + m, :a.read_len] = r
    src[b0:b0 + m] = gs
offsets = np.arange(a.reads, dtype=np.uint64) * np.uint64(stride)
lens = np.full(a.reads, a.read_len, np.uint64)

# warm-up of both paths (scratch allocation)
wb = ix.upload_flat(data[:256], offsets[:256], lens[:256])
ix.query_batch(wb, 5, 10, float(a.threshold))
gb = ix.synth(SEED, 0, 1, a.genome_len)
ix.exact_batch(gb, wb)
gb.free()
wb.free()
ix.stats_reset()

t0 = time.perf_counter()
rb = ix.upload_flat(data, offsets, lens)
hits, nh = ix.query_batch(rb, 5, 10, float(a.threshold))          # Miekki.cpp:417 (-e parameters)
rb.free()
t_query = time.perf_counter() - t0

# candidates grouped by genome (Miekki.cpp:744-752 groups them by genome file)
rid = np.repeat(np.arange(a.reads), nh[:a.reads])
mask = np.arange(hits.shape[1])[None, :] < nh[:a.reads, None]
gid = hits["genome"][:a.reads][mask]
order = np.argsort(gid, kind="stable")
gid, rid = gid[order], rid[order]
bounds = np.flatnonzero(np.diff(gid)) + 1
starts = np.concatenate(([0], bounds))
ends = np.concatenate((bounds, [len(gid)]))
t_group = time.perf_counter() - t0 - t_query

pairs = len(gid)
true_pairs = int((gid == src[rid]).sum())
positive = 0
sum_inter = 0
jacc_sum = 0.0
for s, e in zip(starts, ends):
    g = int(gid[s])
    rr = rid[s:e]
    gb = ix.synth(SEED, g, 1, a.genome_len)
    cb = ix.upload_flat(data, offsets[rr], lens[rr])
    nB, inter, uni = ix.exact_batch(gb, cb)
    cb.free()
    gb.free()
    pos = inter > 0                                                 # Miekki.cpp:852: only printed if nb_inter > 0
    positive += int(pos.sum())
    sum_inter += int(inter.sum())
    jacc_sum += float((inter[pos] / uni[pos]).sum())
t_total = time.perf_counter() - t0
st = ix.stats()
t_exact = t_total - t_query - t_group
print(json.dumps({
    "workload": "C5: %d x %.1f Mbp genomes -k %d -h %d, %d error-free %d bp reads, -e -s %d" %
                (a.genomes, a.genome_len / 1e6, a.k, a.h, a.reads, a.read_len, a.threshold),
    "build_s": build_s, "query_s": t_query, "group_s": t_group, "exact_s": t_exact, "total_s": t_total,
    "candidate_pairs": pairs, "pairs_on_source_genome": true_pairs, "genomes_with_candidates": int(len(starts)),
    "lines_printed": positive, "sum_nb_inter": sum_inter, "mean_real_jaccard": jacc_sum / max(1, positive),
    "reads_per_s_end_to_end": a.reads / t_total, "pairs_per_s_exact_phase": pairs / max(t_exact, 1e-9),
    "genomes_per_s_exact_phase": len(starts) / max(t_exact, 1e-9),
    "exact_device_ms": st["exact_ms"], "exact_device_ms_per_genome": st["exact_ms"] / max(1, len(starts)),
    "genome_kmers_per_s_device": len(starts) * a.genome_len / max(st["exact_ms"], 1e-9) * 1e3,
    "expected_nb_inter_per_true_pair": a.read_len - a.k + 1}))
