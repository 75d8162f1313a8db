#!/bin/bash
# Per-kernel time table of one bench configuration (run on the GPU box):  tools/kprof.sh [size]
# 181 launches per step (178 kernels; memsets are not kernels) -> skip the setup forward + 3 warm-up steps, capture 2 steps.
SIZE=${1:-512}
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --size $SIZE"
mkdir -p gpurun_out
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 605 -c 356 --csv --log-file gpurun_out/launches_$SIZE.csv $B > gpurun_out/ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_$SIZE.csv 2
