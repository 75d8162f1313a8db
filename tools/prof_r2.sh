#!/bin/bash
# Round-2 ncu --set full captures (GPU box) of the largest bandwidth kernels of the 512x512 iteration and of the
# cluster weight-gradient kernel; reports land in gpurun_out/ (summaries: tools/ncu_stalls.py, ncu --page raw --csv).
mkdir -p gpurun_out
timeout 200 python tools/one_iter.py 512 1 > gpurun_out/p_plain.log 2>&1 || exit 1
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/r02_$1 -f python tools/one_iter.py 512 1 > gpurun_out/p_$1.log 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1_full_raw.csv 2>/dev/null
}
cap upcat_bwd_a upcat_bwd_a_kernel 0
cap upcat_apply upcat_apply_kernel 4
cap upcat_stats upcat_stats_merged_kernel 4
cap bn_bwd_fast_stats bn_bwd_fast_kernel 0
cap wgrad wgrad_halo_kernel 1
cap bn_bwd_top_stats bn_bwd_top_kernel 0
