#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric on its config 2 (SURVEY.md 8d "C2"):

    10,000 synthetic 5 Mbp genomes indexed at -k 31 -h 20 (per GPU), 100,000 error-free
    1 kbp reads; metric = query throughput in kbp/s.

One step = one pass of the query hot path (read sketch + Bloom mask -> fingerprint scan of
the bucket-major matrix -> threshold + bounded-heap top-k) over all reads.

  value     reads already resident in HBM, hit lists left in HBM (device-timed, CUDA events)
  e2e       the C-ABI call a host program makes: reads in pinned host memory -> H2D ->
            query -> hit lists D2H, all inside the timed region
  roofline  the scan kernel: algorithmic bytes (sum_q A(q) * N, counted by the kernel's
            producer of the lists) / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference binary (oracle/_ref/Miekki, all host threads) on a
            bounded sample of the same reads against the same index

N > 1 (torchrun, one rank per GPU): genome-sharded, weak scaling -- every rank holds its own
10,000-genome shard (N x 10,000 genomes in total), every rank scans all reads against its
shard, the bounded heap is chained through the ranks in ascending id order (NCCL send/recv
of 24 MB), `value` = sum over ranks of the read kbp each scored per second.

`--impl reference` times the reference's own CPU implementation on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x5EED_B200
METRIC = "query_throughput_kbp_per_s"
UNIT = "kbp/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # workload overrides (defaults = BASELINE config 2)
    ap.add_argument("--genomes", type=int, default=10_000, help="genomes per GPU")
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--read-len", type=int, default=1_000)
    ap.add_argument("--sub-rate", type=float, default=0.0)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-H", "--hbits", type=int, default=20)
    ap.add_argument("--threshold", type=int, default=200)
    ap.add_argument("--cpu-reads", type=int, default=2_000, help="reads in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--build-e2e-genomes", type=int, default=64,
                    help="genomes pushed through the host->index path to report build Gbp/s e2e")
    return ap.parse_args()


def workload_name(a):
    return ("C2: %d x %.1f Mbp synthetic genomes per GPU, -k %d -h %d, %d error-free %d bp reads"
            % (a.genomes, a.genome_len / 1e6, a.k, a.hbits, a.reads, a.read_len)
            if a.sub_rate == 0 else
            "%d x %.1f Mbp genomes per GPU, -k %d -h %d, %d reads of %d bp with %.0f%% substitutions"
            % (a.genomes, a.genome_len / 1e6, a.k, a.hbits, a.reads, a.read_len, 100 * a.sub_rate))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per scan launch from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


# ---- synthetic reads (host) -------------------------------------------------------------

def make_reads(a, n_genomes_total):
    """Reads cut from the counter-based genomes; returns (uint8 [n, stride] pinned-friendly
    array with 16-byte aligned rows, offsets u64[n+1])."""
    from miekki_b200 import synth
    stride = (a.read_len + 15) // 16 * 16
    data = np.zeros((a.reads, stride), np.uint8)
    blk = 10_000
    for b0 in range(0, a.reads, blk):
        m = min(blk, a.reads - b0)
        r, _, _ = synth.cb_reads(SEED, n_genomes_total, a.genome_len, m, a.read_len, a.sub_rate,
                                 block=b0 // blk)
        data[b0:b0 + m, :a.read_len] = r
    offsets = (np.arange(a.reads, dtype=np.uint64) * np.uint64(stride))
    lens = np.full(a.reads, a.read_len, np.uint64)
    return data, offsets, lens


# ---- reference arm / CPU baseline ---------------------------------------------------------

def write_dump(path, k, h, b, threshold, rows, genome_size, bloom_window_bytes, sketch_size):
    """Uncompressed dump (zstr reads plain files, zstr.hpp:154-167); layout: Miekki.cpp:651-676."""
    n = rows.shape[1]
    with open(path, "wb") as f:
        f.write(np.array([k, h, 8, 5, n, b], "<u4").tobytes())
        f.write(np.array([1 << b], "<u8").tobytes())
        f.write(bytes([0, 0]))
        f.write(np.array([threshold], "<u4").tobytes())
        f.write(bytes([0]))
        rows.tofile(f)
        np.ascontiguousarray(genome_size, "<u8").tofile(f)
        f.write(bloom_window_bytes.tobytes())
        pad = (1 << b) // 8 - len(bloom_window_bytes)
        chunk = bytes(1 << 24)
        while pad > 0:
            f.write(chunk[:min(pad, len(chunk))])
            pad -= len(chunk)
        np.ascontiguousarray(sketch_size, "<u4").tofile(f)


def scratch_dir(need_bytes):
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(d).free > need_bytes * 1.2:
                return tempfile.mkdtemp(prefix="miekki_bench_", dir=d)
        except Exception:
            continue
    return None


def reference_query(a, ix, reads_data, n_sample, steps, warmup, threads):
    """Times oracle/_ref/Miekki (unmodified reference, -t threads) querying the first n_sample
    reads against the SAME index, loaded through its own -i loader from a dump of the index
    (load not timed: the binary reports query time separately, main.cpp:232-234)."""
    from oracle import oracle as orc
    ref = orc.RefBinary()
    if not ref.available:
        return None, "oracle/_ref/Miekki missing (run make -C oracle ref where /root/reference exists)"
    n = ix.n
    need = (1 << a.hbits) * n + (1 << 33) // 8 + 64 * n
    d = scratch_dir(need)
    if d is None:
        return None, "no scratch space for a %.1f GB dump" % (need / 1e9)
    try:
        e = ix.export()
        dump = os.path.join(d, "index.dump")
        write_dump(dump, a.k, a.hbits, 33, a.threshold, e["rows"], e["genome_size"], e["bloom"],
                   e["sketch_size"])
        del e
        fa = os.path.join(d, "sample.fa")
        with open(fa, "wb") as f:
            for i in range(n_sample):
                f.write(b">r%d\n" % i + reads_data[i, :a.read_len].tobytes() + b"\n")
        times = []
        for s in range(warmup + steps):
            out = ref.run(["-i", dump, "-a", fa, "-o", os.path.join(d, "ref_out.txt"), "-t", threads],
                          cwd=d, timeout=3600)
            el = ref.elapsed(out)
            if len(el) < 2:
                return None, "reference output had no timing lines"
            if s >= warmup:
                times.append(el[1])
        return times, None
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---- main ---------------------------------------------------------------------------------

def main():
    a = parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference" and rank != 0:
        return 0                                   # rank 0 alone runs the CPU reference

    if not torch.cuda.is_available():
        print(json.dumps({"impl": a.impl, "error": "no CUDA device: this bench needs a B200"}))
        return 1
    torch.cuda.set_device(local)
    if world > 1 and a.impl == "ours":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import miekki_b200
    nproc = os.cpu_count() or 1

    # ---- setup (untimed): the index, like weights, is state ------------------------------
    ix = miekki_b200.Miekki(k=a.k, h=a.hbits, b=33, threshold=a.threshold, device=local)
    # a high-priority stream: the scan outranks the top-k / NCCL kernels that run beside it
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    ix.set_stream(stream.cuda_stream)
    ix.reserve(a.genomes)
    first = rank * a.genomes if a.impl == "ours" else 0
    ix.set_shard(first)
    t0 = time.perf_counter()
    ix.stats_reset()
    step_g = 128
    insert_wall = 0.0           # wall clock of the insert calls alone (the generator is bench tooling)
    for g0 in range(0, a.genomes, step_g):
        m = min(step_g, a.genomes - g0)
        b = ix.synth(SEED, first + g0, m, a.genome_len)
        t1 = time.perf_counter()
        ix.insert_batch(b)
        insert_wall += time.perf_counter() - t1
        b.free()
    st_build = ix.stats()
    build_wall = time.perf_counter() - t0
    build_kernel_gbps = st_build["bases_sketched"] / max(st_build["sketch_ms"], 1e-9) / 1e6

    total_genomes = a.genomes * (world if a.impl == "ours" else 1)
    if world > 1 and a.impl == "ours":
        # global Bloom filter: byte-wise "lowest rank wins" 
        from miekki_b200 import sharded
        w = ix.bloom_window()
        mine = torch.empty(w, dtype=torch.uint8, device="cuda")
        ix.bloom_get_ptr(mine.data_ptr(), w)
        merged = sharded.merge_bloom(mine).contiguous()
        ix.bloom_set_ptr(merged.data_ptr(), w)
        del merged, mine
        # whole-job build rate: all shards were built concurrently, the slowest rank sets the time
        tb = torch.tensor([st_build["sketch_ms"], insert_wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        build_all = {"kernel_gbp_per_s": world * st_build["bases_sketched"] / max(float(tb[0]), 1e-9) / 1e6,
                     "device_resident_wall_gbp_per_s": world * st_build["bases_sketched"] / float(tb[1]) / 1e6}
    else:
        build_all = {"kernel_gbp_per_s": build_kernel_gbps,
                     "device_resident_wall_gbp_per_s": st_build["bases_sketched"] / insert_wall / 1e9}

    reads_np, offsets, rlens = make_reads(a, total_genomes)
    read_kbp = a.reads * a.read_len / 1e3

    if a.impl == "reference":
        times, why = reference_query(a, ix, reads_np, min(a.cpu_reads, a.reads), a.steps,
                                     min(a.warmup, 1), nproc)
        if times is None:
            print(json.dumps({"impl": "reference", "unavailable": why}))
            return 0
        sample_kbp = min(a.cpu_reads, a.reads) * a.read_len / 1e3
        v = sample_kbp * len(times) / sum(times)
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "genomes_indexed": a.genomes,
                       "reads_per_step": min(a.cpu_reads, a.reads)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": nproc, "kind": "reference",
                             "sample": "first %d of the %d reads per step, against the full %d-genome index "
                                       "(built on the GPU, bit-identical to a reference build, loaded by the "
                                       "reference's own -i loader; load not timed)"
                                       % (min(a.cpu_reads, a.reads), a.reads, a.genomes)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ---- our arm -------------------------------------------------------------------------
    K = 10
    min_int = 0.5 * a.threshold
    pinned = torch.from_numpy(reads_np).pin_memory()
    hits_host = torch.empty((a.reads, K * 24), dtype=torch.uint8).pin_memory()
    nh_host = torch.empty(a.reads, dtype=torch.int32).pin_memory()
    d_heap = torch.zeros((a.reads, K * 24), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(a.reads, dtype=torch.int32, device="cuda")
    resident = ix.upload_flat_ptr(pinned.data_ptr(), offsets, rlens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from miekki_b200 import sharded

    def run_resident(steps):
        """`steps` passes over the reads already in HBM; hit lists stay in HBM."""
        if world == 1:
            for _ in range(steps):
                ix.query_batch(resident, K, 10, min_int, fetch=False)
        else:
            # scan of step i+1 is in flight while step i's heap is chained through the ranks
            sharded.pipelined_query(ix, (resident for _ in range(steps)), d_heap, d_len, K, 10, min_int)

    def run_e2e(steps):
        """Same through host buffers: every step uploads its reads from pinned host memory and
        brings the hit lists back to the host."""
        if world == 1:
            import ctypes as C
            lib = miekki_b200.lib()
            for _ in range(steps):
                b = ix.upload_flat_ptr(pinned.data_ptr(), offsets, rlens)      # H2D of this step's reads
                ix._ck(lib.mk_query_batch(ix._ctx, b._h, K, 10, float(min_int),
                                          C.c_void_p(hits_host.data_ptr()), C.c_void_p(nh_host.data_ptr())))
                b.free()
            return
        live = {}

        def uploads():
            for i in range(steps):
                live[i] = ix.upload_flat_ptr(pinned.data_ptr(), offsets, rlens)
                yield live[i]

        def fetched(i):                                  # last rank: D2H of batch i's hit lists
            hits_host.copy_(d_heap, non_blocking=True)
            nh_host.copy_(d_len, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        sharded.pipelined_query(ix, uploads(), d_heap, d_len, K, 10, min_int, on_result=fetched,
                                after_chain=lambda i: live.pop(i).free())   # its scan has run by then

    def timed(run, steps, warmup):
        run(warmup)
        barrier()
        ix.stats_reset()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        run(steps)
        torch.cuda.synchronize()                         # side / aux streams of the last step
        ev1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        st = ix.stats()
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), st

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms, st = timed(run_resident, a.steps, a.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_dev_ms, e2e_wall_ms, st_e2e = timed(run_e2e, a.steps, max(3, a.warmup))

    # the index exceeds L2 (10.5 GB vs 126 MB) and rows are touched in hash order: no flush needed
    value = world * read_kbp * a.steps / (dev_ms / 1e3)
    e2e_value = world * read_kbp * a.steps / (e2e_wall_ms / 1e3)
    peak, peak_src = measured_peak()
    scan_gbs = st["scan_row_bytes"] / max(st["scan_ms"], 1e-9) / 1e6
    traffic = ncu_traffic()

    # ---- build throughput through the host path (extra, not the headline) ----------------
    build_e2e = None
    if a.build_e2e_genomes > 0 and rank == 0:
        from miekki_b200 import synth
        m = a.build_e2e_genomes
        bx = miekki_b200.Miekki(k=a.k, h=a.hbits, b=33, threshold=a.threshold, device=local)
        bx.reserve(2 * m)
        hosts = [synth.cb_bases(SEED, g, 0, a.genome_len).tobytes() for g in range(min(m, 8))]
        seqs = [hosts[i % len(hosts)] for i in range(m)]
        bx.insert_sequences(seqs)                          # warm-up: same call, sizes all scratch
        t0 = time.perf_counter()
        bx.insert_sequences(seqs)
        dt = time.perf_counter() - t0
        build_e2e = m * a.genome_len / dt / 1e9
        bx.close()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        times, why = reference_query(a, ix, reads_np, min(a.cpu_reads, a.reads), 1, 0, nproc)
        if times is None:
            cpu = {"value": None, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": "unavailable: " + why}
        else:
            cpu = {"value": min(a.cpu_reads, a.reads) * a.read_len / 1e3 / times[0], "unit": UNIT,
                   "cores": nproc, "kind": "reference",
                   "sample": "first %d of the %d reads, against the full %d-genome index (dumped from the GPU "
                             "build, loaded by the reference's -i loader; load not timed); query seconds = %.2f"
                             % (min(a.cpu_reads, a.reads), a.reads, a.genomes, times[0])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": workload_name(a), "genomes_per_gpu": a.genomes,
                "genomes_indexed_total": total_genomes, "reads_per_step": a.reads,
                "index_bytes_per_gpu": int((1 << a.hbits) * a.genomes),
                "l2": "index >> L2 and rows are hit in hash order; no flush between steps",
                "value_definition": "sum over ranks of read kbp scored against that rank's shard per second "
                                    "(weak scaling: job read throughput = value / n_gpus on an n_gpus x larger index)",
                "sharding": "genomes (matrix columns), contiguous ascending ids per rank",
            },
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] // a.steps),
                    "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] // a.steps + (a.reads * (K * 24 + 4) if world > 1 else 0)),
                    "ms_per_step": e2e_wall_ms / a.steps},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "scan_kernel", "achieved": scan_gbs, "peak": peak,
                         "unit": "GB/s", "frac": scan_gbs / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(st["scan_row_bytes"] // max(1, st["scan_launches"])),
                         "launch_ms": st["scan_ms"] / max(1, st["scan_launches"]),
                         # ncu's DRAM bytes per algorithmic byte (one capture), scaled to this launch size
                         "traffic": (int(traffic["dram_bytes_per_algorithmic_byte"] *
                                         (st["scan_row_bytes"] // max(1, st["scan_launches"])))
                                     if traffic else None),
                         "traffic_source": (traffic or {}).get("source")},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "phases_ms_per_step": {"read_sketch": st["read_sketch_ms"] / a.steps, "scan": st["scan_ms"] / a.steps,
                                   "topk": st["topk_ms"] / a.steps},
            "surviving_buckets_per_read": st["scan_rows"] / max(1, a.steps * a.reads),
            "build": {"kernel_gbp_per_s": build_kernel_gbps, "setup_wall_s": build_wall,
                      "e2e_host_to_index_gbp_per_s": build_e2e, "genomes": a.genomes,
                      "all_ranks": build_all},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
