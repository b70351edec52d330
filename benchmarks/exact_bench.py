#!/usr/bin/env python
"""Exact mode (-e, BASELINE config 5 shape) throughput: per genome, set B of a 5 Mbp genome is
built and `reads_per_genome` 1 kbp candidate reads are intersected with it (mk_exact).
Prints one JSON line: genomes/s, (read, genome) pairs/s, device ms per genome."""
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
ap.add_argument("--genomes", type=int, default=40)
ap.add_argument("--genome-len", type=int, default=5_000_000)
ap.add_argument("--reads-per-genome", type=int, default=50)
ap.add_argument("--read-len", type=int, default=1000)
a = ap.parse_args()

ix = miekki_b200.Miekki(k=31, h=20)
rng = np.random.default_rng(1)
genomes = [synth.cb_bases(7, g, 0, a.genome_len).tobytes() for g in range(min(a.genomes, 8))]
work = []
for g in range(a.genomes):
    s = genomes[g % len(genomes)]
    reads = []
    for _ in range(a.reads_per_genome):
        p = int(rng.integers(0, a.genome_len - a.read_len))
        reads.append(s[p:p + a.read_len])
    work.append((s, reads))
ix.exact([work[0][0]], work[0][1])          # warm-up
ix.stats_reset()
t0 = time.perf_counter()
tot_inter = 0
for s, reads in work:
    nB, inter, uni = ix.exact([s], reads)
    tot_inter += int(inter.sum())
dt = time.perf_counter() - t0
st = ix.stats()
print(json.dumps({"genomes": a.genomes, "pairs": a.genomes * a.reads_per_genome, "wall_s": dt,
                  "genomes_per_s": a.genomes / dt, "pairs_per_s": a.genomes * a.reads_per_genome / dt,
                  "device_ms_per_genome": st["exact_ms"] / a.genomes,
                  "genome_kmers_per_s_device": a.genomes * a.genome_len / (st["exact_ms"] / 1e3),
                  "check_inter": tot_inter}))
