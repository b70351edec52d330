#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's metric on its config 2 (SURVEY.md 8d "C2"):

    10,000 synthetic 5 Mbp genomes indexed at -k 31 -h 20 (per GPU), 100,000 error-free
    1 kbp reads; metric = query throughput in kbp/s.

One step = one pass of the query hot path (read sketch + Bloom mask -> fingerprint scan of
the bucket-major matrix -> threshold + bounded-heap top-k) over all reads.

  value     reads already resident in HBM, hit lists left in HBM (device-timed, CUDA events)
  e2e       the calls a host program makes: every step's reads go from pinned host memory to
            the GPU (mk_batch_upload_flat), are queried (mk_sketch_async / mk_scan_async /
            mk_topk_slot, as sharded.pipelined_query issues them) and the hit lists come back
            to the host, all inside the timed region; steps are streamed: the copies of
            step i +- 1 run beside the scan of step i (MIEKKI_BENCH_E2E_SERIAL=1 on one GPU:
            one blocking mk_query_batch per step instead)
  roofline  the scan kernel: algorithmic bytes (sum_q A(q) * N, counted by the kernel's
            producer of the lists) / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference binary (oracle/_ref/Miekki, all host threads) on a
            bounded sample of the same reads against the same index
  parity_checked  (untimed) the hit lists the timed path produced for sampled reads, compared
            with the oracle's filter chained over every rank's counts (tests/ hold the full suite)

N > 1 (torchrun, one rank per GPU): genome-sharded, weak scaling -- every rank holds its own
10,000-genome shard (N x 10,000 genomes in total), every rank scans all reads against its
shard, the bounded heap is chained through the ranks in ascending id order (NCCL send/recv
of 24 MB per step, in tiles of reads), `value` = sum over ranks of the read kbp each scored
per second.

Extra blocks on the same line, measured in the same run (none of them inside the timed region
of `value`): `c3_strong` (BASELINE config 3: 100,000 genomes at -h 17 in total, split over the
ranks, 10 kbp reads with 5 % substitutions: strong scaling), `build` (config 4: Gbp/s of the
sketch kernels, of the device-resident build and of host memory -> index), `exact` (config 5 in
small: true k-mer intersections of the hits), `query_vs_genomes` (metric vs #genomes indexed).

`--impl reference` times the reference's own CPU implementation on the same config; that
process never loads libmiekki_b200.so (the index it queries is prepared by a separate helper
process, benchmarks/make_dump).
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
K_RESULTS = 10


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
    ap.add_argument("--parity-reads", type=int, default=64, help="reads whose hit lists are checked (untimed)")
    ap.add_argument("--build-e2e-genomes", type=int, default=256,
                    help="genomes pushed through the host->index path to report build Gbp/s e2e")
    # extra blocks (BASELINE configs 3, 4, 5 and the metric's "vs #genomes indexed")
    ap.add_argument("--no-extras", action="store_true", help="headline only")
    ap.add_argument("--c3-genomes", type=int, default=100_000, help="genomes of config 3 in TOTAL (split over the ranks)")
    ap.add_argument("--c3-reads", type=int, default=20_000)
    ap.add_argument("--c3-read-len", type=int, default=10_000)
    ap.add_argument("--c3-sub-rate", type=float, default=0.05)
    ap.add_argument("--c3-hbits", type=int, default=17)
    ap.add_argument("--c3-steps", type=int, default=2)
    ap.add_argument("--exact-reads", type=int, default=2_000)
    ap.add_argument("--exact-genomes", type=int, default=48, help="genomes intersected in the exact-mode block")
    ap.add_argument("--sweep-genomes", default="100,1000,3000")
    ap.add_argument("--sweep-reads", type=int, default=20_000)
    return ap.parse_args()


def workload_name(a):
    return ("C2: %d x %.1f Mbp synthetic genomes per GPU, -k %d -h %d, %d error-free %d bp reads"
            % (a.genomes, a.genome_len / 1e6, a.k, a.hbits, a.reads, a.read_len)
            if a.sub_rate == 0 else
            "%d x %.1f Mbp genomes per GPU, -k %d -h %d, %d reads of %d bp with %.0f%% substitutions"
            % (a.genomes, a.genome_len / 1e6, a.k, a.hbits, a.reads, a.read_len, 100 * a.sub_rate))


def config_of(a, world):
    """The `config` object of the JSON line: identical for both arms."""
    return {
        "workload": workload_name(a), "genomes_per_gpu": a.genomes,
        "genomes_indexed_total": a.genomes * world, "reads_per_step": a.reads,
        "index_bytes_per_gpu": int((1 << a.hbits) * a.genomes),
        "l2": "index >> L2 and rows are hit in hash order; no flush between steps",
        "value_definition": "sum over ranks of read kbp scored against that rank's shard per second "
                            "(weak scaling: job read throughput = value / n_gpus on an n_gpus x larger index)",
        "sharding": "genomes (matrix columns), contiguous ascending ids per rank",
    }


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_config):
    """DRAM bytes per algorithmic byte from a committed ncu capture of THIS configuration
    (profiles/scan_traffic.json holds one entry per captured configuration), else None: a ratio
    captured on another shape says nothing about this one (L2 serves part of the rows at -h 17)."""
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    try:
        entries = json.load(open(p))
    except Exception:
        return None
    if isinstance(entries, dict):
        entries = [entries]
    for e in entries:
        if all(e.get("config", {}).get(k) == v for k, v in kernel_config.items()):
            return e
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

def make_reads(n_reads, read_len, sub_rate, n_genomes_total, genome_len, long_reads=False):
    """Reads cut from the counter-based genomes; returns (uint8 [n, stride] pinned-friendly
    array with 16-byte aligned rows, offsets u64[n], lens u64[n])."""
    from miekki_b200 import synth
    stride = (read_len + 15) // 16 * 16
    data = np.zeros((n_reads, stride), np.uint8)
    blk = 10_000
    for b0 in range(0, n_reads, blk):
        m = min(blk, n_reads - b0)
        gen = synth.cb_reads_block if long_reads else synth.cb_reads
        r, _, _ = gen(SEED, n_genomes_total, genome_len, m, read_len, sub_rate, block=b0 // blk)
        data[b0:b0 + m, :read_len] = r
    offsets = (np.arange(n_reads, dtype=np.uint64) * np.uint64(stride))
    lens = np.full(n_reads, read_len, np.uint64)
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


def write_sample_fasta(path, reads_data, read_len, n_sample):
    with open(path, "wb") as f:
        for i in range(n_sample):
            f.write(b">r%d\n" % i + reads_data[i, :read_len].tobytes() + b"\n")


def time_reference_binary(ref, dump, fa, workdir, threads, steps, warmup):
    """`steps` timed runs of the unmodified binary (-i dump -a sample), `warmup` untimed ones
    before; the time of a run is the binary's own second `elapsed time:` line = query_file
    (main.cpp:232-234), so the index load is outside."""
    times = []
    for s in range(warmup + steps):
        out = ref.run(["-i", dump, "-a", fa, "-o", os.path.join(workdir, "ref_out.txt"), "-t", threads],
                      cwd=workdir, timeout=3600)
        el = ref.elapsed(out)
        if len(el) < 2:
            return None
        if s >= warmup:
            times.append(el[1])
    return times


def reference_arm(a):
    """--impl reference: oracle/_ref/Miekki (the unmodified reference, all host threads) on a
    bounded sample of the config's reads against the config's full index.  This process maps
    neither libmiekki_b200.so nor CUDA: the index file is written by benchmarks/make_dump, a
    separate process (GPU build, bit-identical to a reference -t 1 build -- tests/test_gpu_*)."""
    from oracle import oracle as orc
    nproc = os.cpu_count() or 1

    def unavailable(why):
        print(json.dumps({"impl": "reference", "unavailable": why}))
        return 0

    ref = orc.RefBinary()
    if not ref.available:
        return unavailable("oracle/_ref/Miekki missing (run make -C oracle ref where /root/reference exists)")
    helper = os.path.join(ROOT, "benchmarks", "make_dump")
    if not os.path.exists(helper):
        return unavailable("benchmarks/make_dump missing (run make)")
    need = (1 << a.hbits) * a.genomes + (1 << 33) // 8 + 64 * a.genomes
    d = scratch_dir(need)
    if d is None:
        return unavailable("no scratch space for a %.1f GB dump" % (need / 1e9))
    try:
        dump = os.path.join(d, "index.dump")
        t0 = time.perf_counter()
        r = subprocess.run([helper, str(a.k), str(a.hbits), "33", str(a.threshold), str(a.genomes),
                            str(a.genome_len), str(SEED), "0", dump], capture_output=True, text=True, timeout=1800)
        if r.returncode != 0:
            return unavailable("make_dump failed: " + (r.stderr.strip().split("\n") or ["?"])[-1][:200])
        prep_s = time.perf_counter() - t0
        n_sample = min(a.cpu_reads, a.reads)
        reads_np, _, _ = make_reads(min(a.reads, (n_sample + 9_999) // 10_000 * 10_000), a.read_len, a.sub_rate,
                                    a.genomes, a.genome_len)
        fa = os.path.join(d, "sample.fa")
        write_sample_fasta(fa, reads_np, a.read_len, n_sample)
        times = time_reference_binary(ref, dump, fa, d, nproc, a.steps, a.warmup)
        if times is None:
            return unavailable("reference output had no timing lines")
    finally:
        shutil.rmtree(d, ignore_errors=True)
    sample_kbp = n_sample * a.read_len / 1e3
    v = sample_kbp * len(times) / sum(times)
    sample = ("each step = the first %d of the config's %d reads (bounded sample: one full step takes minutes on "
              "these cores), against the config's full %d-genome index; timed = the binary's own query time "
              "(2nd `elapsed time:`), index load excluded; index file prepared by benchmarks/make_dump in a "
              "separate process (%.0f s, untimed); this process loads no GPU code"
              % (n_sample, a.reads, a.genomes, prep_s))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": config_of(a, 1),
        "sample_reads_per_step": n_sample,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if a.gpus > 1:
        line["note"] = ("the reference has no multi-GPU path: this is the same 1 x %d-genome CPU run at every N; "
                        "the GPU arm's value at N > 1 is summed over N shards (an N x larger index), so a ratio "
                        "of the two values is only meaningful at N = 1" % a.genomes)
    print(json.dumps(line))
    return 0


# ---- parity of the timed path's output (untimed) ---------------------------------------------

def check_hit_lists(ix, sample_seqs, sample_idx, got_hits, got_len, first_id, min_score, min_int,
                    world, rank, dist, torch):
    """Hit lists of the sampled reads as the timed path produced them (on the LAST rank:
    got_hits [n, K] HIT_DTYPE, got_len [n]) against the oracle's filter (Miekki.cpp:376-397 with
    libstdc++'s heap) chained in rank order over every rank's shared-fingerprint counts.
    -> dict for the JSON line (identical on all ranks)."""
    from oracle import oracle as orc
    import miekki_b200
    K = K_RESULTS
    counts, _ = ix.query_counts(sample_seqs)                      # [m, n_local]
    ss, gs = ix.stats_arrays()
    m, n_local = counts.shape
    firsts, widths = [first_id], [n_local]
    if world > 1:
        meta = torch.tensor([first_id, n_local], dtype=torch.int64, device="cuda")
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta)
        firsts, widths = [int(x[0]) for x in metas], [int(x[1]) for x in metas]
        wmax = max(widths)

        def gather(x, dtype, cols):                               # shards may differ in width: pad
            pad = np.zeros((x.shape[0], wmax), dtype)
            pad[:, :cols] = x
            t = torch.from_numpy(pad.view(np.uint8).reshape(-1)).cuda()
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [o.cpu().numpy().view(dtype).reshape(x.shape[0], wmax)[:, :widths[r]] for r, o in enumerate(out)]
        all_counts = gather(counts, np.uint32, n_local)
        all_ss = [x[0] for x in gather(ss[None, :], np.uint32, n_local)]
        all_gs = [x[0] for x in gather(gs[None, :], np.uint64, n_local)]
    else:
        all_counts, all_ss, all_gs = [counts], [ss], [gs]
    ok, bad, floats_identical = True, [], True
    if rank == world - 1:
        for j, i in enumerate(sample_idx):
            heap = np.zeros(K, miekki_b200.HIT_DTYPE)
            n = 0
            for r in range(world):
                n = orc.filter_chain(np.ascontiguousarray(all_counts[r][j]), firsts[r], all_ss[r], all_gs[r], K,
                                     min_score, min_int, heap, n, r == world - 1)
            g = got_hits[i, :got_len[i]]
            same = (n == got_len[i] and np.array_equal(g["genome"], heap["genome"][:n])
                    and np.array_equal(g["matches"], heap["matches"][:n])
                    and np.allclose(g["jaccard"], heap["jaccard"][:n], rtol=1e-6, atol=0)
                    and np.allclose(g["intersection"], heap["intersection"][:n], rtol=1e-6, atol=0))
            if same:
                floats_identical = floats_identical and g.tobytes() == heap[:n].tobytes()
            else:
                ok = False
                bad.append(int(i))
    if world > 1:
        flag = torch.tensor([1 if ok else 0, 1 if floats_identical else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok, floats_identical = bool(flag[0].item()), bool(flag[1].item())
    return {"reads": len(sample_idx), "ok": ok, "floats_bit_identical": floats_identical,
            "against": "oracle filter (libstdc++ heap) chained in rank order over every rank's counts",
            "mismatching_reads": bad[:8]}


# ---- main ---------------------------------------------------------------------------------

def main():
    a = parse_args()
    if a.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0                               # rank 0 alone runs the CPU reference
        return reference_arm(a)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if not torch.cuda.is_available():
        print(json.dumps({"impl": a.impl, "error": "no CUDA device: this bench needs a B200"}))
        return 1
    torch.cuda.set_device(local)
    import miekki_b200
    from miekki_b200 import sharded, synth
    if world > 1:
        sharded.nccl_env_defaults()              # one P2P channel: a waiting receive holds one SM, not many
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nproc = os.cpu_count() or 1
    K = K_RESULTS

    # config-3 reads are cut on the host while the GPU builds and runs config 2
    c3 = {"reads": None}
    c3_on = not a.no_extras and a.c3_genomes >= world and a.c3_reads > 0

    def gen_c3():
        c3["reads"] = make_reads(a.c3_reads, a.c3_read_len, a.c3_sub_rate, a.c3_genomes, a.genome_len,
                                 long_reads=True)
    c3_thread = threading.Thread(target=gen_c3, daemon=True)
    if c3_on:
        c3_thread.start()

    # ---- setup (untimed): the index, like weights, is state ------------------------------
    ix = miekki_b200.Miekki(k=a.k, h=a.hbits, b=33, threshold=a.threshold, device=local)
    # a high-priority stream: the scan outranks the top-k / NCCL kernels that run beside it
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    ix.set_stream(stream.cuda_stream)
    if world > 1:
        # one SM for the NCCL kernels of the heap chain: they do not fit beside a ring-scan CTA and
        # would wait for the next batch's scan to end (2 GPUs: 136.4 -> 134.3 ms per step)
        ix.set_scan_spare_sms(1)

    def build_shard(index, first, count, genome_len):
        """count synthetic genomes (ids first ..) generated on the device and inserted;
        -> (stats of the build, wall seconds of the insert calls alone)."""
        index.reserve(count)
        index.set_shard(first)
        index.stats_reset()
        insert_wall = 0.0
        for g0 in range(0, count, 128):
            m = min(128, count - g0)
            b = index.synth(SEED, first + g0, m, genome_len)
            t1 = time.perf_counter()
            index.insert_batch(b)
            insert_wall += time.perf_counter() - t1
            b.free()
        return index.stats(), insert_wall

    def merge_bloom_all(index):
        """global Bloom filter: byte-wise "lowest rank wins" """
        if world == 1:
            return
        w = index.bloom_window()
        mine = torch.empty(w, dtype=torch.uint8, device="cuda")
        index.bloom_get_ptr(mine.data_ptr(), w)
        merged = sharded.merge_bloom(mine).contiguous()
        torch.cuda.current_stream().synchronize()      # the fold must have run before the context copies it
        index.bloom_set_ptr(merged.data_ptr(), w)

    def build_rates(st, insert_wall):
        """kernel and device-resident build Gbp/s, this rank and whole job (slowest rank sets the time)"""
        mine = {"kernel_gbp_per_s": st["bases_sketched"] / max(st["sketch_ms"], 1e-9) / 1e6,
                "device_resident_wall_gbp_per_s": st["bases_sketched"] / max(insert_wall, 1e-9) / 1e9}
        if world == 1:
            return mine, dict(mine)
        tb = torch.tensor([st["sketch_ms"], insert_wall * 1e3, float(st["bases_sketched"])], dtype=torch.float64,
                          device="cuda")
        mx = tb.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tb.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return mine, {"kernel_gbp_per_s": float(sm[2]) / max(float(mx[0]), 1e-9) / 1e6,
                      "device_resident_wall_gbp_per_s": float(sm[2]) / max(float(mx[1]), 1e-9) / 1e6}

    t0 = time.perf_counter()
    first = rank * a.genomes
    st_build, insert_wall = build_shard(ix, first, a.genomes, a.genome_len)
    build_wall = time.perf_counter() - t0
    merge_bloom_all(ix)
    build_mine, build_all = build_rates(st_build, insert_wall)
    total_genomes = a.genomes * world

    reads_np, offsets, rlens = make_reads(a.reads, a.read_len, a.sub_rate, total_genomes, a.genome_len)
    read_kbp = a.reads * a.read_len / 1e3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Workload:
        """One index + one read set: resident and end-to-end passes, timing, parity check."""

        def __init__(self, index, reads_np, offsets, rlens, min_int):
            self.ix, self.min_int = index, float(min_int)
            self.n = len(rlens)
            self.offsets, self.rlens = offsets, rlens
            self.reads_np = reads_np
            # Two orders of the same reads, used by alternate steps (B = A reversed): consecutive
            # batches then differ, so a hit list that belongs to another batch (a stale buffer in
            # the pipelined multi-GPU path) cannot pass the parity check by coincidence.
            self.pinned = [torch.from_numpy(reads_np).pin_memory(),
                           torch.from_numpy(np.ascontiguousarray(reads_np[::-1])).pin_memory()]
            self.hits_host = torch.empty((self.n, K * 24), dtype=torch.uint8).pin_memory()
            self.nh_host = torch.empty(self.n, dtype=torch.int32).pin_memory()
            self.d_heap = torch.zeros((self.n, K * 24), dtype=torch.uint8, device="cuda")
            self.d_len = torch.zeros(self.n, dtype=torch.int32, device="cuda")
            self.resident = [index.upload_flat_ptr(p.data_ptr(), offsets, rlens) for p in self.pinned]
            self.step_no = 0            # steps run so far: step s uses order s & 1
            self.last_order = 0         # order of the reads of the last end-to-end step

        def run_resident(self, steps):
            """`steps` passes over the reads already in HBM; hit lists stay in HBM."""
            first_step, self.step_no = self.step_no, self.step_no + steps
            if world == 1:
                for s_ in range(first_step, first_step + steps):
                    self.ix.query_batch(self.resident[s_ & 1], K, 10, self.min_int, fetch=False)
            else:
                # scan of step i+1 is in flight while step i's heap is chained through the ranks
                sharded.pipelined_query(self.ix, (self.resident[s_ & 1] for s_ in range(first_step, first_step + steps)),
                                        self.d_heap, self.d_len, K, 10, self.min_int)

        def run_e2e(self, steps):
            """Same through host buffers: every step uploads its reads from pinned host memory and
            brings the hit lists back to the host."""
            first_step, self.step_no = self.step_no, self.step_no + steps
            if steps:
                self.last_order = (first_step + steps - 1) & 1
            if world == 1 and os.environ.get("MIEKKI_BENCH_E2E_SERIAL"):
                # one blocking call per step (upload, query, read back in sequence): the simplest use
                import ctypes as C
                lib = miekki_b200.lib()
                for s_ in range(first_step, first_step + steps):
                    b = self.ix.upload_flat_ptr(self.pinned[s_ & 1].data_ptr(), self.offsets, self.rlens)   # H2D
                    self.ix._ck(lib.mk_query_batch(self.ix._ctx, b._h, K, 10, self.min_int,
                                                   C.c_void_p(self.hits_host.data_ptr()),
                                                   C.c_void_p(self.nh_host.data_ptr())))
                    b.free()
                return
            # Streaming use, one GPU or many: the reads of step i + 1 are uploaded (and sketched) while
            # step i scans, the hit lists of step i come back while step i + 1 scans.
            live = {}

            def uploads():
                for i in range(steps):
                    live[i] = self.ix.upload_flat_ptr(self.pinned[(first_step + i) & 1].data_ptr(), self.offsets,
                                                      self.rlens)
                    yield live[i]

            def fetched(i):                              # last rank: D2H of batch i's hit lists
                self.hits_host.copy_(self.d_heap, non_blocking=True)
                self.nh_host.copy_(self.d_len, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            sharded.pipelined_query(self.ix, uploads(), self.d_heap, self.d_len, K, 10, self.min_int,
                                    on_result=fetched, after_chain=lambda i: live.pop(i).free())

        def timed(self, run, steps, warmup):
            run(warmup)
            barrier()
            self.ix.stats_reset()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record(stream)
            run(steps)
            torch.cuda.synchronize()                     # side / aux streams of the last step
            ev1.record(stream)
            barrier()
            wall = time.perf_counter() - t0
            dev_ms = ev0.elapsed_time(ev1)
            st = self.ix.stats()
            t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
            # every rank's own scan time and step time: under the power cap GPUs of one box differ by a
            # few per cent and the job runs at the pace of the slowest
            mine = torch.tensor([st["scan_ms"] / max(1, steps), dev_ms / max(1, steps)], dtype=torch.float64,
                                device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                every = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(every, mine)
            else:
                every = [mine]
            self.per_rank = {"scan_ms_per_step": [round(float(e[0]), 2) for e in every],
                             "step_ms": [round(float(e[1]), 2) for e in every]}
            return float(t[0]), float(t[1]), st

        def parity(self, n_sample, first_id):
            """the hit lists of the LAST end-to-end step (hits_host / nh_host on the last rank)"""
            n_sample = max(1, min(n_sample, self.n))
            idx = np.unique(np.linspace(0, self.n - 1, n_sample).astype(np.int64))   # positions in the last batch
            L = int(self.rlens[0])
            src = (self.n - 1 - idx) if self.last_order else idx                   # the reads at those positions
            seqs = [self.reads_np[i, :L].tobytes() for i in src]
            got = self.hits_host.numpy().view(miekki_b200.HIT_DTYPE).reshape(self.n, K)
            return check_hit_lists(self.ix, seqs, idx, got, self.nh_host.numpy().view(np.uint32), first_id, 10,
                                   self.min_int, world, rank, dist, torch)

        def close(self):
            for r in self.resident:
                r.free()

    # ---- headline: config 2 ---------------------------------------------------------------
    w2 = Workload(ix, reads_np, offsets, rlens, 0.5 * a.threshold)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms, st = w2.timed(w2.run_resident, a.steps, a.warmup)
    per_rank = w2.per_rank
    clocks = sampler.stop() if rank == 0 else None
    e2e_dev_ms, e2e_wall_ms, st_e2e = w2.timed(w2.run_e2e, a.steps, max(3, a.warmup))
    parity = w2.parity(a.parity_reads, first)

    # the index exceeds L2 (10.5 GB vs 126 MB) and rows are touched in hash order: no flush needed
    value = world * read_kbp * a.steps / (dev_ms / 1e3)
    e2e_value = world * read_kbp * a.steps / (e2e_wall_ms / 1e3)
    peak, peak_src = measured_peak()
    scan_gbs = st["scan_row_bytes"] / max(st["scan_ms"], 1e-9) / 1e6
    traffic = ncu_traffic({"h": a.hbits, "genomes": a.genomes, "read_len": a.read_len})
    alg_per_launch = int(st["scan_row_bytes"] // max(1, st["scan_launches"]))

    extras = {}

    # ---- config 3: 100,000 genomes at -h 17 in total, 10 kbp reads with 5 % substitutions ----
    if c3_on:
        gfirst, gcount = sharded.shard_range(a.c3_genomes, rank, world)
        ix3 = miekki_b200.Miekki(k=a.k, h=a.c3_hbits, b=33, threshold=a.threshold, device=local)
        ix3.set_stream(stream.cuda_stream)
        t0 = time.perf_counter()
        st3b, wall3b = build_shard(ix3, gfirst, gcount, a.genome_len)
        merge_bloom_all(ix3)
        c3_build_s = time.perf_counter() - t0
        _, c3_build_all = build_rates(st3b, wall3b)
        c3_thread.join()
        r3, o3, l3 = c3["reads"]
        # -s 0: at -h 17 every 5 Mbp sketch is saturated, genome_size = 0 (quirk G2) and the default
        # -s leaves every list empty; -s 0 keeps the all-ties heap order in play
        w3 = Workload(ix3, r3, o3, l3, 0.0)
        # two warm-up steps: the pipelined path alternates between two sets of buffers, both are sized before timing
        d3_ms, _, st3 = w3.timed(w3.run_resident, a.c3_steps, 2)
        per_rank3 = w3.per_rank
        _, e3_wall_ms, _ = w3.timed(w3.run_e2e, 2, 2)
        par3 = w3.parity(16, gfirst)
        c3_kbp = a.c3_reads * a.c3_read_len / 1e3
        extras["c3_strong"] = {
            "workload": "C3: %d x %.1f Mbp genomes in total at -h %d split over %d GPU(s) (%d on this rank), %d reads "
                        "of %d bp with %.0f%% substitutions, -s 0" % (a.c3_genomes, a.genome_len / 1e6, a.c3_hbits,
                                                                      world, gcount, a.c3_reads, a.c3_read_len,
                                                                      100 * a.c3_sub_rate),
            "scaling": "strong", "steps": a.c3_steps, "ms_per_step": d3_ms / a.c3_steps,
            "job_kbp_per_s": c3_kbp * a.c3_steps / (d3_ms / 1e3),
            "e2e_job_kbp_per_s": 2 * c3_kbp / (e3_wall_ms / 1e3),
            "scan_gbs_algorithmic_this_rank": st3["scan_row_bytes"] / max(st3["scan_ms"], 1e-9) / 1e6,
            "scan_frac_of_peak": st3["scan_row_bytes"] / max(st3["scan_ms"], 1e-9) / 1e6 / peak,
            "phases_ms_per_step": {"read_sketch": st3["read_sketch_ms"] / a.c3_steps,
                                   "scan": st3["scan_ms"] / a.c3_steps, "topk": st3["topk_ms"] / a.c3_steps},
            "surviving_buckets_per_read": st3["scan_rows"] / max(1, a.c3_steps * a.c3_reads),
            "per_rank": per_rank3,
            # DRAM bytes per algorithmic byte of the scan from the committed ncu capture of this shape
            # (tiled kernel), when there is one; not measured in this run
            "dram_bytes_per_algorithmic_byte_ncu": (ncu_traffic({"h": a.c3_hbits, "genomes": gcount,
                                                                 "read_len": a.c3_read_len, "tiled": True})
                                                    or {}).get("dram_bytes_per_algorithmic_byte"),
            # what binds the tiled kernel, from its committed ncu capture (12,500 genomes: the pipe
            # utilisation is a property of the kernel, not of the shard's width)
            "scan_bound_ncu": {k: (ncu_traffic({"h": a.c3_hbits, "read_len": a.c3_read_len, "tiled": True})
                                   or {}).get(k)
                               for k in ("shared_memory_pipe_pct", "issue_slots_busy_pct",
                                         "dram_throughput_pct_of_theoretical", "source")},
            "scan_kernel": "scan_tiled_kernel (rows staged once per 46 reads; DESIGN.md section 4)",
            "build_s": c3_build_s, "build_all_ranks": c3_build_all,
            "parity_checked": par3,
        }
        w3.close()
        ix3.close()
        del w3, ix3
        torch.cuda.empty_cache()

    # ---- config 5 in small: exact mode on the hits of a read sample (rank 0's shard) ----------
    if not a.no_extras and a.exact_reads > 0 and rank == 0:
        from oracle import oracle as orc
        ne = min(a.exact_reads, a.reads)
        seqs = [reads_np[i, :a.read_len].tobytes() for i in range(ne)]
        t0 = time.perf_counter()
        hits = ix.query(seqs, 5, 10, float(a.threshold))             # Miekki.cpp:741
        per_genome = {}
        for i, hh in enumerate(hits):
            for g in hh["genome"]:
                per_genome.setdefault(int(g), []).append(i)
        genomes = sorted(g for g in per_genome if first <= g < first + a.genomes)[:a.exact_genomes]
        ix.stats_reset()
        pairs, t_exact, checked, exact_ok, kmers = 0, 0.0, 0, True, 0
        sort_ms, sort_same, exact_ms_total = 0.0, True, 0.0
        for g in genomes:
            rec = ix.synth(SEED, g, 1, a.genome_len)                 # the genome file's one record, in HBM
            rb = ix.upload([seqs[i] for i in per_genome[g]])
            # the sort-merge variant of set B (radix sort + binary search) on the same pair, for comparison
            os.environ["MIEKKI_EXACT_SORT"] = "1"
            before = ix.stats()["exact_ms"]
            sB, s_inter, s_uni = ix.exact_batch(rec, rb)
            sort_ms += ix.stats()["exact_ms"] - before
            os.environ["MIEKKI_EXACT_SORT"] = "0"
            ix.stats_reset()
            t1 = time.perf_counter()
            nB, inter, uni = ix.exact_batch(rec, rb)
            t_exact += time.perf_counter() - t1
            exact_ms_total += ix.stats()["exact_ms"]
            sort_same = sort_same and sB == nB and np.array_equal(s_inter, inter) and np.array_equal(s_uni, uni)
            pairs += len(per_genome[g])
            kmers += a.genome_len - a.k + 1
            if checked < 2:                                          # the oracle redoes two genomes on the CPU
                host = rec.download(0, a.genome_len)
                oB, opairs = orc.exact([host], [seqs[i] for i in per_genome[g]], a.k)
                exact_ok = exact_ok and oB == nB and [(int(x), int(y)) for x, y in zip(inter, uni)] == opairs
                checked += 1
            rec.free()
            rb.free()
        st_x = {"exact_ms": exact_ms_total}
        extras["exact"] = {
            "workload": "C5 in small: hits (top 5, -s %d) of the first %d reads, true k-mer intersection against %d of "
                        "their genomes (5 Mbp each, resident in HBM)" % (a.threshold, ne, len(genomes)),
            "pairs": pairs, "genomes": len(genomes),
            "device_ms_per_genome": st_x["exact_ms"] / max(1, len(genomes)),
            "call_ms_per_genome": 1e3 * t_exact / max(1, len(genomes)),
            "pairs_per_s": pairs / max(t_exact, 1e-9),
            "genome_kmers_per_s_device": kmers / max(st_x["exact_ms"], 1e-9) * 1e3,
            "roofline": {"bound": "hbm", "kernel": "exact_insert_kernel",
                         "algorithmic_bytes_per_genome": 8 * (a.genome_len - a.k + 1),
                         "achieved": 8 * kmers / max(st_x["exact_ms"], 1e-9) / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": 8 * kmers / max(st_x["exact_ms"], 1e-9) / 1e6 / peak,
                         "note": "8 B per genome k-mer (the key a sort or a set must move at least once) over the "
                                 "device time of the whole exact phase; the kernel is a random-access insert, "
                                 "bound by L2 atomics, not by streaming bandwidth"},
            "sort_merge_variant": {"device_ms_per_genome": sort_ms / max(1, len(genomes)), "same_results": sort_same,
                                   "what": "set B as a radix-sorted array (cub::DeviceRadixSort over 2k bits) with "
                                           "binary search, MIEKKI_EXACT_SORT=1; the hash sets above are the default"},
            "parity_checked": {"genomes": checked, "ok": exact_ok and sort_same,
                               "against": "oracle (CPU sets) on the same pairs"},
        }

    # ---- metric vs #genomes indexed (rank 0, single GPU shapes) --------------------------------
    if not a.no_extras and a.sweep_genomes and rank == 0:
        sweep = []
        for n_g in [int(x) for x in a.sweep_genomes.split(",") if x]:
            ixs = miekki_b200.Miekki(k=a.k, h=a.hbits, b=33, threshold=a.threshold, device=local)
            ixs.set_stream(stream.cuda_stream)
            build_shard(ixs, 0, n_g, a.genome_len)
            rs, os_, ls = make_reads(a.sweep_reads, a.read_len, 0.0, n_g, a.genome_len)
            pin = torch.from_numpy(rs).pin_memory()
            res = ixs.upload_flat_ptr(pin.data_ptr(), os_, ls)
            for _ in range(2):
                ixs.query_batch(res, K, 10, 0.5 * a.threshold, fetch=False)
            torch.cuda.synchronize()
            ixs.stats_reset()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(3):
                ixs.query_batch(res, K, 10, 0.5 * a.threshold, fetch=False)
            ev1.record(stream)
            torch.cuda.synchronize()
            sts = ixs.stats()
            ms = ev0.elapsed_time(ev1) / 3
            sweep.append({"genomes": n_g, "reads": a.sweep_reads, "ms_per_step": ms,
                          "kbp_per_s": a.sweep_reads * a.read_len / 1e3 / (ms / 1e3),
                          "scan_gbs_algorithmic": sts["scan_row_bytes"] / max(sts["scan_ms"], 1e-9) / 1e6})
            res.free()
            ixs.close()
        sweep.append({"genomes": a.genomes, "reads": a.reads, "ms_per_step": dev_ms / a.steps,
                      "kbp_per_s": value / world, "scan_gbs_algorithmic": scan_gbs})
        extras["query_vs_genomes"] = sweep

    # ---- build throughput through the host path (extra, not the headline) ----------------
    build_e2e = None
    if a.build_e2e_genomes > 0 and rank == 0:
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
        from oracle import oracle as orc
        ref = orc.RefBinary()
        n_sample = min(a.cpu_reads, a.reads)
        why, times = None, None
        if not ref.available:
            why = "oracle/_ref/Miekki missing (run make -C oracle ref where /root/reference exists)"
        else:
            need = (1 << a.hbits) * ix.n + (1 << 33) // 8 + 64 * ix.n
            d = scratch_dir(need)
            if d is None:
                why = "no scratch space for a %.1f GB dump" % (need / 1e9)
            else:
                try:
                    e = ix.export()
                    dump = os.path.join(d, "index.dump")
                    write_dump(dump, a.k, a.hbits, 33, a.threshold, e["rows"], e["genome_size"], e["bloom"],
                               e["sketch_size"])
                    del e
                    fa = os.path.join(d, "sample.fa")
                    write_sample_fasta(fa, reads_np, a.read_len, n_sample)
                    times = time_reference_binary(ref, dump, fa, d, nproc, 1, 0)
                    if times is None:
                        why = "reference output had no timing lines"
                finally:
                    shutil.rmtree(d, ignore_errors=True)
        if times is None:
            cpu = {"value": None, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": "unavailable: " + why}
        else:
            cpu = {"value": n_sample * a.read_len / 1e3 / times[0], "unit": UNIT,
                   "cores": nproc, "kind": "reference",
                   "sample": "first %d of the %d reads, against the full %d-genome index (dumped from the GPU "
                             "build, loaded by the reference's -i loader; load not timed); query seconds = %.2f"
                             % (n_sample, a.reads, a.genomes, times[0])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_of(a, world),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] // a.steps),
                    # the streaming path reads the hit lists back with its own copy (not in the library's count)
                    "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] // a.steps +
                                              (0 if (world == 1 and os.environ.get("MIEKKI_BENCH_E2E_SERIAL"))
                                               else a.reads * (K * 24 + 4))),
                    "path": ("one blocking mk_query_batch per step" if (world == 1 and os.environ.get("MIEKKI_BENCH_E2E_SERIAL"))
                             else "sharded.pipelined_query: upload of step i + 1 and read-back of step i beside the scan"),
                    "ms_per_step": e2e_wall_ms / a.steps},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "scan_kernel", "achieved": scan_gbs, "peak": peak,
                         "unit": "GB/s", "frac": scan_gbs / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_per_launch,
                         "launch_ms": st["scan_ms"] / max(1, st["scan_launches"]),
                         # DRAM bytes are not measured in this run (that needs ncu): the figure is the
                         # committed capture of the same configuration scaled to this launch size, or null
                         "traffic": (int(traffic["dram_bytes_per_algorithmic_byte"] * alg_per_launch)
                                     if traffic else None),
                         "traffic_source": ((traffic or {}).get("source", "") +
                                            " -- ncu capture of this configuration, scaled by the algorithmic "
                                            "bytes of this run's launches; not measured in this run")
                                           if traffic else "no ncu capture of this configuration is committed",
                         # SURVEY 8(d)'s second bound: one byte compare per algorithmic byte, against the issue rate
                         "second_bound": {"what": "byte compares (one per algorithmic byte; 32 genomes per lane "
                                                  "and LOP3 in the bit-plane layout) against the SM issue rate",
                                          "byte_compares_per_s": scan_gbs * 1e9,
                                          "issue_slots_busy_pct_ncu": (traffic or {}).get("issue_slots_busy_pct"),
                                          "source": "same ncu capture as traffic"}},
            "parity_checked": parity,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "phases_ms_per_step": {"read_sketch": st["read_sketch_ms"] / a.steps, "scan": st["scan_ms"] / a.steps,
                                   "topk": st["topk_ms"] / a.steps},
            "surviving_buckets_per_read": st["scan_rows"] / max(1, a.steps * a.reads),
            "per_rank": per_rank,
            "build": {"workload": "C4 point: %d x %.1f Mbp genomes per GPU, -k %d -h %d (config 3 adds -h %d)"
                                  % (a.genomes, a.genome_len / 1e6, a.k, a.hbits, a.c3_hbits),
                      "kernel_gbp_per_s": build_mine["kernel_gbp_per_s"],
                      "device_resident_wall_gbp_per_s": build_mine["device_resident_wall_gbp_per_s"],
                      "setup_wall_s": build_wall,
                      "e2e_host_to_index_gbp_per_s": build_e2e,
                      "e2e_host_to_index_note": "%d inserts of %d distinct host genomes (pageable memory -> pinned "
                                                "staging -> H2D -> sketch), second call timed"
                                                % (a.build_e2e_genomes, min(a.build_e2e_genomes, 8)),
                      "genomes": a.genomes, "all_ranks": build_all,
                      "ncu": "profiles/ holds the per-kernel ncu details (issue slots, L2 %) of the build kernels"},
        }
        line.update(extras)
        print(json.dumps(line))
    ok = parity["ok"] and all(extras.get(k, {}).get("parity_checked", {}).get("ok", True) for k in extras
                              if isinstance(extras.get(k), dict))
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
