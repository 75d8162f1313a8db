#!/bin/bash
# A/B of environment switches on the device-resident iteration rate:  tools/ab_env.sh "VAR=1 VAR2=3" "VAR=2" ...
# prints one line per configuration ("" = default)
for cfg in "$@"; do
  v=$(env $cfg python bench.py --steps 60 --warmup 10 --no-e2e --no-cpu --no-secondary --concurrent 1 2>/dev/null | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f it/s  %.4f ms' % (d['value'], d['ms_per_step']))")
  echo "[$cfg] $v"
done
