#!/bin/bash
# Last check of a build on the GPU box (no ncu): tests, smoke, the default bench line, the SRGAN training bench line.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -s > gpurun_out/f_tests.txt 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/f_bench_dip.log 2>&1
timeout 300 python bench.py --workload gan_train > gpurun_out/f_bench_gan_train.log 2>&1
tail -3 gpurun_out/f_tests.txt; tail -3 gpurun_out/f_smoke.txt; cut -c1-200 gpurun_out/f_bench_dip.log | tail -2; cut -c1-200 gpurun_out/f_bench_gan_train.log | tail -1
