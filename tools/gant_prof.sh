#!/bin/bash
# ncu launch list of the SRGAN training step (one GPU): per-launch gpu__time_duration of 3 whole steps.
# Run under gpurun; outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python tools/gant_step.py 3 > gpurun_out/gant_step_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gant_launches.csv \
    python tools/gant_step.py 3 > gpurun_out/gant_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/gant_launches.csv --steps 3 > gpurun_out/gant_kernel_table.txt 2>&1 || true
head -40 gpurun_out/gant_kernel_table.txt
