#!/bin/bash
# tools/r2_evidence.sh TAG -- the round-2 evidence set on one B200 (copied into profiles/r02_* afterwards):
# GPU suite, warm per-kernel tables on both grids, ncu launch list of bench.py, ncu --set full of the graded kernel and of the
# restructured kernels on the large grid, persistent-loop stamps, the bench line and the reference arm.
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 1200 python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -2 ${O}_pytest.log
python tools/time_phases.py > ${O}_phases_b1.log 2>&1
python tools/time_phases.py 2048 256 30 5 > ${O}_phases_b3.log 2>&1
ROMS_B200_S2_PERSIST=1 ROMS_B200_S2_PROF=1 python tools/time_phases.py > ${O}_persist.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-roofline --no-check > ${O}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step3d_t_v8 -s 4 -c 1 -o ${O}_v8_b3 python tools/prof_step3d_t.py 2048 256 30 > ${O}_ncu_v8.log 2>&1
for k in t3dmix2_geo_roll_kernel pre_step3d_t_roll_kernel uv3dmix2_roll_kernel rhs3d_roll_kernel pre_step3d_uv_march_kernel step2d_kernel; do
  s=4; [ $k = step2d_kernel ] && s=300
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o ${O}_$k python tools/time_phases.py 2048 256 30 1 > ${O}_ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step2d_kernel -s 300 -c 1 -o ${O}_step2d_b1 python tools/time_phases.py 512 64 30 1 > ${O}_ncu_step2d_b1.log 2>&1
timeout 900 python bench.py > ${O}_bench.log 2>&1
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > ${O}_bench_ref.log 2>&1
tail -1 ${O}_bench.log | cut -c1-400; tail -1 ${O}_bench_ref.log | cut -c1-300
grep -h "step2d\|persistent" ${O}_persist.log
