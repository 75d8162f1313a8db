#!/bin/bash
# End-of-round evidence run (GPU box): tests, both bench workloads, ncu launch list of the DIP iteration.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/f_tests.txt 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/f_bench_dip.log 2>&1
timeout 300 python bench.py --workload gan_eval > gpurun_out/f_bench_gan.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --concurrent 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 700 --csv --log-file gpurun_out/f_launches_512.csv $B > gpurun_out/f_ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/f_launches_512.csv > gpurun_out/f_ktable.txt 2>&1
