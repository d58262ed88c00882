#!/bin/bash
# tools/ncu_capture.sh <tag> [kernel regex]: one `ncu --set full` capture of the bench step's kernels (run under gpurun).
# Writes gpurun_out/<tag>.ncu-rep; read it with tools/ncu_read.py on the CPU box.
tag=${1:-prof}
rx=${2:-simulate_pairs_tp|format_fastq}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$rx" -s ${SKIP:-4} -c ${COUNT:-2} -f -o gpurun_out/$tag \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-e2e-cli --min-seconds 0 ${BENCH_ARGS:-} > gpurun_out/$tag.log 2>&1
tail -2 gpurun_out/$tag.log
