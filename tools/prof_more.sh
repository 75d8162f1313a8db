#!/bin/bash
# ncu --set full captures of the L0 instances of the element-wise kernels (second iteration of tools/one_iter.py)
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -o gpurun_out/v5_$1 -f python tools/one_iter.py 512 2 > gpurun_out/p_$1.log 2>&1
}
cap upcat_apply upcat_apply_kernel 9
cap upcat_bwd_c upcat_bwd_c_kernel 5
cap upcat_stats upcat_stats_merged 9
cap bn_act "bn_act_kernel" 38
cap input_pack input_pack32 1
