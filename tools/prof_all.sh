#!/bin/bash
# End-of-round profile refresh (GPU box): ncu --set full of the dominant DIP launch and of the generator kernels.
mkdir -p gpurun_out
timeout 200 python tools/one_iter.py 512 2 > gpurun_out/p_plain.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo2_kernel -s 43 -c 1 -o gpurun_out/v3_halo2_l0 -f python tools/one_iter.py 512 2 > gpurun_out/p1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo2_kernel -s 44 -c 1 -o gpurun_out/v3_halo2_l0_1x1 -f python tools/one_iter.py 512 2 > gpurun_out/p2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo2_kernel -s 40 -c 1 -o gpurun_out/v3_gan_halo2_96 -f python tools/gan_one.py 32 > gpurun_out/p3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo2_kernel -s 76 -c 1 -o gpurun_out/v3_gan_halo2_384 -f python tools/gan_one.py 32 > gpurun_out/p4.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/v3_gan_launches.csv python tools/gan_one.py 32 > gpurun_out/p5.log 2>&1
