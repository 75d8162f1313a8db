#!/bin/bash
# End-of-round evidence run (GPU box): tests, smoke, the bench workloads, ncu launch lists of the DIP iteration and of
# the SRGAN training step.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -s > gpurun_out/f_tests.txt 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/f_bench_dip.log 2>&1
timeout 300 python bench.py --workload gan_eval > gpurun_out/f_bench_gan.log 2>&1
timeout 300 python bench.py --workload gan_train > gpurun_out/f_bench_gan_train.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --concurrent 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 700 --csv --log-file gpurun_out/f_launches_512.csv $B > gpurun_out/f_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/f_launches_512.csv > gpurun_out/f_ktable.txt 2>&1
DSR_GAN_ONE_STREAM=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_gant_launches.csv \
    python tools/gant_step.py 3 > gpurun_out/f_gant_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/f_gant_launches.csv --steps 3 > gpurun_out/f_gant_ktable.txt 2>&1
tail -3 gpurun_out/f_tests.txt; tail -3 gpurun_out/f_smoke.txt; cut -c1-200 gpurun_out/f_bench_dip.log | tail -2; cut -c1-200 gpurun_out/f_bench_gan_train.log | tail -1
