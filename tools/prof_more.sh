timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_halo -s 23 -c 1 -o gpurun_out/v4_wgrad_l1u1 -f python tools/one_iter.py 512 2 > gpurun_out/p9.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 9 -c 1 -o gpurun_out/v4_gemm_l0d1 -f python tools/one_iter.py 512 2 > gpurun_out/p10.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:bn_bwd_fast_kernel<true>|bn_bwd_fast_kernelILb1" -s 15 -c 1 -o gpurun_out/v4_bnbwd_apply -f python tools/one_iter.py 512 2 > gpurun_out/p11.log 2>&1
