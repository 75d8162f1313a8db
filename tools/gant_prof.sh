#!/bin/bash
# ncu evidence of the SRGAN training step (one GPU): per-launch gpu__time_duration of 3 whole steps, then --set full
# captures of two VGG convolutions (gconv_kernel) and the largest weight gradient (gwgrad_kernel).
# Run under gpurun; outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python tools/gant_step.py 3 > gpurun_out/gant_step_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gant_launches.csv \
    python tools/gant_step.py 3 > gpurun_out/gant_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/gant_launches.csv --steps 3 > gpurun_out/gant_kernel_table.txt 2>&1 || true
head -30 gpurun_out/gant_kernel_table.txt
# gconv launch 93 of a step = VGG conv1_2 (64 -> 64 at 224 x 224, batch 8: 29.6 GFLOP); 97 = conv3_2 (256 -> 256 at 56 x 56)
for skip in 93 97; do
  ncu --set full --clock-control none --import-source on -k regex:gconv_kernel --launch-skip $skip --launch-count 1 \
      -f -o gpurun_out/gant_gconv_$skip python tools/gant_step.py 1 > gpurun_out/gant_ncu_full_$skip.log 2>&1
  ncu -i gpurun_out/gant_gconv_$skip.ncu-rep --page raw --csv > gpurun_out/gant_gconv_${skip}_raw.csv 2>/dev/null
done
ncu --set full --clock-control none --import-source on -k regex:gwgrad_kernel --launch-skip 6 --launch-count 1 \
    -f -o gpurun_out/gant_gwgrad python tools/gant_step.py 1 > gpurun_out/gant_ncu_full_wg.log 2>&1
ncu -i gpurun_out/gant_gwgrad.ncu-rep --page raw --csv > gpurun_out/gant_gwgrad_raw.csv 2>/dev/null
ls -la gpurun_out/gant_*
