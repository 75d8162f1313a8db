#!/bin/bash
# A/B of environment switches: iteration rate AND the weight-gradient class time of the bench line
for cfg in "$@"; do
  v=$(env $cfg python bench.py --steps 60 --warmup 10 --no-e2e --no-cpu --no-secondary --concurrent 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
c = d['roofline']['classes']
print('%.1f it/s  %.4f ms   wgrad %.3f ms (%.3f of burst)  halo2 %.3f ms  gemm %.3f ms' % (d['value'], d['ms_per_step'], c['wgrad_halo_kernel']['ms_per_step'], c['wgrad_halo_kernel']['frac'], c['conv_halo2_kernel']['ms_per_step'], c['conv_gemm_kernel']['ms_per_step']))")
  echo "[$cfg] $v"
done
