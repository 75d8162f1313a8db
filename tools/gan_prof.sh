set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r_tests.txt 2>&1
timeout 300 python bench.py --workload gan_eval --steps 5 > gpurun_out/gan_b3.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/gan_launches2.csv python tools/gan_one.py 32 > gpurun_out/gan_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv9_out -s 1 -c 1 -o gpurun_out/gan_conv9 -f python tools/gan_one.py 32 > gpurun_out/gan_ncu3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo2_kernel -s 76 -c 1 -o gpurun_out/gan_halo2_384 -f python tools/gan_one.py 32 > gpurun_out/gan_ncu4.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_halo2_kernel -s 40 -c 1 -o gpurun_out/gan_halo2_96 -f python tools/gan_one.py 32 > gpurun_out/gan_ncu5.log 2>&1
