#!/bin/bash
# tools/ncu_kernels.sh TAG -- ncu --set full of one warm launch of every mid-fraction kernel on the BENCHMARK3 grid (2048x256x30), step2d included
mkdir -p gpurun_out; O=gpurun_out/$1
for k in step2d_kernel rhs3d_kernel rhs3d_sum_kernel uv3dmix2_kernel uv3dmix2_sum_kernel geo_dTdz_kernel t3dmix2_geo_kernel prsgrd_T_kernel prsgrd_P_kernel prsgrd_ruv_kernel pre_step3d_t_kernel pre_step3d_uv_kernel step3d_uv1_kernel step3d_uv2_kernel kpp_levels_kernel kpp_spline_kernel; do
  s=4; [ $k = step2d_kernel ] && s=300
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s $s -c 1 -o ${O}_$k python tools/time_phases.py 2048 256 30 1 > ${O}_ncu_$k.log 2>&1
  ncu -i ${O}_$k.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct 2>/dev/null | tail -1 > ${O}_sum_$k.csv
  echo "$k: $(cat ${O}_sum_$k.csv | cut -c1-400)"
done
