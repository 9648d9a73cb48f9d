#!/bin/bash
# tools/r2_first_call.sh -- everything round 2 should measure FIRST, in one B200 call (about 3 GPU-minutes):
#   tools/gpu.sh 420 'bash tools/r2_first_call.sh'
# 1. the GPU suite including the experimental step3d_t variant (never run on hardware in round 1),
# 2. warm per-kernel times of the code restructured after round 1's last measurement (t3dmix2, uv3dmix2, prsgrd, step2d),
# 3. step3d_t: production vs the experimental variant with each knob (ROMS_B200_S3T_EXP bit 0 = decoupled producers,
#    bit 1 = x-neighbours by shuffle) and a stagger sweep,
# 4. the bench line, 5. the launch list.
mkdir -p gpurun_out; O=gpurun_out/r2a
ROMS_B200_TEST_V7=1 timeout 300 python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
python tools/time_phases.py > ${O}_phases.log 2>&1
python tools/prof_step3d_t.py 2048 256 30 > ${O}_s3t_v6_b3.log 2>&1
python tools/prof_step3d_t.py 1024 512 50 > ${O}_s3t_v6_n50.log 2>&1
for e in 0 1 2 3; do
  ROMS_B200_STEP3D_T_V7=1 ROMS_B200_S3T_EXP=$e timeout 60 python tools/prof_step3d_t.py 2048 256 30 > ${O}_s3t_v7_exp${e}_b3.log 2>&1
done
for s in 0 200 800 1600; do
  ROMS_B200_STEP3D_T_V7=1 ROMS_B200_S3T_EXP=3 ROMS_B200_S3T_STAGGER=$s timeout 60 python tools/prof_step3d_t.py 2048 256 30 > ${O}_s3t_v7_stag${s}_b3.log 2>&1
done
ROMS_B200_STEP3D_T_V7=1 timeout 60 python tools/prof_step3d_t.py 1024 512 50 > ${O}_s3t_v7_n50.log 2>&1
python bench.py --no-cpu > ${O}_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-roofline > /dev/null 2>&1
tail -3 ${O}_pytest.log; cat ${O}_phases.log; for f in ${O}_s3t_*.log; do echo "$f: $(tail -1 $f)"; done; tail -1 ${O}_bench.log | cut -c1-600
