#!/usr/bin/env python
"""One GPU's share of BASELINE config 3 in small: N genomes at -h 17 (default 12,500), R reads of
10 kbp with 5 % substitutions (default 4,000), a few passes of the query path; prints the scan
kernel's time and algorithmic GB/s.  Short enough to sit under ncu:

    python benchmarks/c3_scan_probe.py && ncu --set full --clock-control none --import-source on \
        -k regex:scan_tiled -c 1 -o gpurun_out/tiled python benchmarks/c3_scan_probe.py

MIEKKI_SCAN_TILED=0 gives the ring kernel on the same input."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import miekki_b200  # noqa: E402
from miekki_b200 import synth  # noqa: E402

SEED = 0x5EED_B200


def main():
    n = int(os.environ.get("PROBE_GENOMES", "12500"))
    r = int(os.environ.get("PROBE_READS", "4000"))
    h = int(os.environ.get("PROBE_H", "17"))
    rl = int(os.environ.get("PROBE_READ_LEN", "10000"))
    passes = int(os.environ.get("PROBE_PASSES", "3"))
    ix = miekki_b200.Miekki(k=31, h=h, threshold=200)
    ix.reserve(n)
    for g0 in range(0, n, 128):
        b = ix.synth(SEED, g0, min(128, n - g0), 5_000_000)
        ix.insert_batch(b)
        b.free()
    reads, _, _ = synth.cb_reads_block(SEED, n, 5_000_000, r, rl, 0.05, block=1)
    batch = ix.upload([x.tobytes() for x in reads])
    ix.query_batch(batch, 10, 10, 100.0, fetch=False)          # warm-up (default -s: empty lists, cheap top-k)
    ix.stats_reset()
    import time
    t0 = time.perf_counter()
    for _ in range(passes):
        ix.query_batch(batch, 10, 10, 100.0, fetch=False)
    wall_ms = 1e3 * (time.perf_counter() - t0) / passes
    st = ix.stats()
    print(json.dumps({"genomes": n, "reads": r, "h": h, "read_len": rl,
                      "query_ms_per_pass": wall_ms, "scan_ms_per_pass": st["scan_ms"] / passes,
                      "scan_gbs_algorithmic": st["scan_row_bytes"] / max(st["scan_ms"], 1e-9) / 1e6,
                      "read_sketch_ms_per_pass": st["read_sketch_ms"] / passes,
                      "topk_ms_per_pass": st["topk_ms"] / passes,
                      "rows_per_read": st["scan_rows"] / passes / r,
                      "tiled": os.environ.get("MIEKKI_SCAN_TILED", "auto")}))
    ix.close()


if __name__ == "__main__":
    main()
