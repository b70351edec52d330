#!/bin/bash
# Round 2: the commands behind profiles/r02_* and gpurun_out/r02_* (each block was one gpurun call;
# an ncu run only ever follows the same command run plainly in the same call).
set -x
mkdir -p gpurun_out
# --- parity
python -m pytest tests -m gpu -x -q
python -c "import __graft_entry__ as g; g.smoke()"
# --- headline + extra blocks, both arms (1 GPU)
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_n1.json
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json
# --- N GPUs (gpurun --gpus N):  tests/test_gpu_nccl.py runs the NCCL chain and the one-process CLI on distinct devices
# python -m pytest tests/test_gpu_nccl.py -x -q
# python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus N --steps 5 --warmup 3
# --- config 3's shape: tiled scan vs ring scan, CUDA events, then ncu
export PROBE_READS=20000
python benchmarks/c3_scan_probe.py > gpurun_out/r02_probe_tiled.json &&
ncu --set full --clock-control none --import-source on -k regex:scan_tiled -s 1 -c 1 -o gpurun_out/r02_tiled_v5 python benchmarks/c3_scan_probe.py
MIEKKI_SCAN_TILED=0 python benchmarks/c3_scan_probe.py > gpurun_out/r02_probe_ring.json &&
MIEKKI_SCAN_TILED=0 ncu --set full --clock-control none -k regex:scan_kernel -s 1 -c 1 -o gpurun_out/r02_ring_h17 python benchmarks/c3_scan_probe.py
python benchmarks/c3_scan_probe.py > /dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|topk|sketch_reads|sort_lists|account" -c 400 --csv \
    --log-file gpurun_out/r02_launches_c3_probe.csv python benchmarks/c3_scan_probe.py
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --build-e2e-genomes 0 > /dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|topk|sketch_reads|sort_lists|account" -c 400 --csv \
    --log-file gpurun_out/r02_launches_query_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --build-e2e-genomes 0
# --- file -> index through the CLI (1,000 genome files in /dev/shm)
python benchmarks/file_to_index.py > gpurun_out/r02_file_to_index.json
# here, without a GPU:
#   ncu -i gpurun_out/r02_tiled_v5.ncu-rep --page details --csv > profiles/r02_scan_tiled_ncu_details.csv
#   python benchmarks/sass_histogram.py > profiles/r02_sass_histogram.md
# --- later in the round
# the ring scan at config 2 under ncu --set full (a full count tile: launch 22)
A="--steps 2 --warmup 3 --no-cpu-baseline --no-extras --build-e2e-genomes 0"
python bench.py $A > /dev/null &&
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 21 -c 1 -o gpurun_out/r02_scan_c2 python bench.py $A
# one rank's share of the pipelined multi-GPU path without a second GPU (config 3 on 2 GPUs, config 2)
python benchmarks/chain_probe.py > gpurun_out/r02_chain_probe.json
PROBE_H=20 PROBE_GENOMES=10000 PROBE_READS=100000 PROBE_READ_LEN=1000 PROBE_MIN_INT=100 python benchmarks/chain_probe.py > gpurun_out/r02_chain_probe_c2.json
# build sweep, exact mode and file -> index through the command line
python benchmarks/build_sweep.py --genomes 4096 --host-genomes 256 > gpurun_out/r02_build_sweep_n1.jsonl
python benchmarks/exact_cli.py > gpurun_out/r02_exact_cli.json
python benchmarks/file_to_index.py > gpurun_out/r02_file_to_index_c.json
# N GPUs with the chain's time stamps:  MIEKKI_CHAIN_TRACE=gpurun_out/chain_trace python -m torch.distributed.run ... bench.py --gpus N
#   ncu -i gpurun_out/r02_scan_c2.ncu-rep --page details --csv > profiles/r02_scan_ring_c2_ncu_details.csv
