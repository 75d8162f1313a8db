#!/bin/bash
# Round-2 final-build evidence (GPU box): bench line, reference arm, ncu launch list + DRAM-byte table, in-kernel stamp
# timelines, --set full captures of the kernels that changed late in the round.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
python bench.py > gpurun_out/r02_bench_dip.json 2> gpurun_out/r02_bench_dip.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --concurrent 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 320 -c 700 --csv --log-file gpurun_out/r02_launches_bench_512.csv $B > gpurun_out/ncu_b.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_launches_bench_512.csv > gpurun_out/r02_kernel_table_512.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 320 -c 500 --csv --log-file gpurun_out/r02_bw.csv $B > gpurun_out/ncu_b2.log 2>&1
python tools/ncu_bw_table.py gpurun_out/r02_bw.csv > gpurun_out/r02_dram_bytes_per_launch_512.txt
export DSR_B200_LIB=$PWD/deep-super-resolution_b200/libdsr_b200_kstamp.so DSR_TIMELINE=2
DSR_EXP_SKIP_WGRAD=1 python tools/scales_exp.py 512 30 5 > gpurun_out/ks_final_512_nowg.txt 2>&1
DSR_EXP_SKIP_WGRAD=1 python tools/scales_exp.py 64 30 5 > gpurun_out/ks_final_64_nowg.txt 2>&1
unset DSR_B200_LIB DSR_TIMELINE
cap() {
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/r02_$1 -f python tools/one_iter.py 512 1 > gpurun_out/p_$1.log 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1_full_raw.csv 2>/dev/null
  python tools/ncu_stalls.py gpurun_out/r02_$1.ncu-rep 14 > gpurun_out/r02_$1_stalls.txt 2>&1
}
cap upcat_apply8 upcat_apply8_kernel 4
cap upcat_bwd_a upcat_bwd_a_kernel 0
cap bn_bwd_fast_stats bn_bwd_fast_kernel 0
cap bn_bwd_fast_apply bn_bwd_fast_kernel 1
cap bn_act bn_act_kernel 18
