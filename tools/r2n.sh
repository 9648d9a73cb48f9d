#!/bin/bash
# tools/r2n.sh TAG -- GPU suite, step3d_t timings + ncu, per-kernel tables on BENCHMARK1 and the BENCHMARK3 grid, bench (N=1)
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "2048 256 30" "1024 512 50" "512 64 30"; do
  n=$(echo $g | tr ' ' x)
  ROMS_B200_S3T_VERBOSE=1 timeout 120 python tools/prof_step3d_t.py $g > ${O}_s3t_$n.log 2>&1
done
python tools/time_phases.py > ${O}_phases_b1.log 2>&1
python tools/time_phases.py 2048 256 30 5 > ${O}_phases_b3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step3d_t_v8 -s 4 -c 1 -o ${O}_v8_b3 python tools/prof_step3d_t.py 2048 256 30 > ${O}_ncu.log 2>&1
timeout 900 python bench.py > ${O}_bench.log 2>&1
for f in ${O}_s3t_*.log; do echo "$f: $(grep -h 'step3d_t ' $f | grep ms | tail -1)"; done
cat ${O}_phases_b3.log; tail -1 ${O}_bench.log | cut -c1-700
