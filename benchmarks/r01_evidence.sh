#!/bin/bash
# Round-1 measurement run on one B200 (gpurun -- 'bash benchmarks/r01_evidence.sh').  Everything
# lands in gpurun_out/ev_*; the summaries judged live in profiles/.  Each ncu capture follows a
# plain run of the same command that exited 0; no number printed under ncu is a bench value.
set -x
O=gpurun_out
mkdir -p $O
Q='--steps 2 --warmup 3 --no-cpu-baseline --build-e2e-genomes 0'

# 1. query step: plain run, launch list of the query kernels, full capture of one scan launch
python bench.py $Q > $O/ev_bench_plain.json 2> $O/ev_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|topk_kernel|sketch_reads|account_rows' \
    -c 400 --csv --log-file $O/ev_launches_query_step.csv python bench.py $Q > $O/ev_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_kernel --launch-skip 9 -c 1 \
    -o $O/ev_scan python bench.py --steps 1 --warmup 1 --no-cpu-baseline --build-e2e-genomes 0 > $O/ev_ncu2.log 2>&1

# 2. build: plain run, launch list of late chunks, full captures of the hashing and streaming kernels
B='benchmarks/build_sweep.py --genomes 4160 --host-genomes 0 --hs 20 --ks 31'
python $B > $O/ev_build_plain.json 2> $O/ev_build_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 -c 120 --csv \
    --log-file $O/ev_launches_build_late.csv python $B > $O/ev_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sketch_dense|resolve_fast|scatter_planes|encode_planes' \
    --launch-skip 120 -c 4 -o $O/ev_build python $B > $O/ev_ncu4.log 2>&1

# 3. sweeps (CUDA events / wall clock, no profiler)
python benchmarks/build_sweep.py --genomes 1024 > $O/ev_sweep_1k.jsonl 2> $O/ev_sweep.err
python benchmarks/build_sweep.py --genomes 10048 --host-genomes 0 > $O/ev_sweep_10k.jsonl 2>> $O/ev_sweep.err
python benchmarks/build_sweep.py --genomes 100032 --host-genomes 0 --hs 17 > $O/ev_sweep_100k_h17.jsonl 2>> $O/ev_sweep.err
python benchmarks/build_sweep.py --genomes 100032 --host-genomes 0 --hs 20 --ks 31 > $O/ev_sweep_100k_h20.jsonl 2>> $O/ev_sweep.err
for n in 100 1000 3000 30000 100000; do
    python bench.py --genomes $n --no-cpu-baseline --build-e2e-genomes 0 --steps 3 > $O/ev_query_n$n.json 2> $O/ev_query_n$n.err
done
python benchmarks/exact_full.py > $O/ev_exact_full.json 2> $O/ev_exact_full.err
python benchmarks/config1_cli.py --check > $O/ev_c1_cli.json 2> $O/ev_c1_cli.err
ls -la $O | tail -40
