#!/bin/bash
# ncu evidence of the SRGAN training step, second pass: launch list of 3 steps + --set full of the dense head kernels.
set -x
mkdir -p gpurun_out
python tools/gant_step.py 3 > gpurun_out/gant_step_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gant_launches.csv \
    python tools/gant_step.py 3 > gpurun_out/gant_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/gant_launches.csv --steps 3 > gpurun_out/gant_kernel_table.txt 2>&1 || true
head -40 gpurun_out/gant_kernel_table.txt
for k in g_dense1_fwd_kernel g_dense1_bwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 --launch-count 1 \
      -f -o gpurun_out/gant_$k python tools/gant_step.py 1 > gpurun_out/gant_ncu_full_$k.log 2>&1
  ncu -i gpurun_out/gant_$k.ncu-rep --page raw --csv > gpurun_out/gant_${k}_raw.csv 2>/dev/null
done
