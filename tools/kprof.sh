#!/bin/bash
# Per-kernel time table of one bench configuration (run on the GPU box):  tools/kprof.sh [size]
# Captures a window of launches past the setup forward and the warm-up steps; tools/ncu_summary.py cuts whole
# iterations out of it (pack_weights_kernel marks the start of one).
SIZE=${1:-512}
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --concurrent 1 --size $SIZE"
mkdir -p gpurun_out
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 700 --csv --log-file gpurun_out/launches_$SIZE.csv $B > gpurun_out/ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_$SIZE.csv
