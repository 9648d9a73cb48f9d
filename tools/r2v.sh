#!/bin/bash
# tools/r2v.sh TAG "name:ENV=v,ENV=v ..." -- parity suite + step3d_t timings of env-selected variants on 2048x256x30 (+ defaults on
# the other grids) + one ncu capture of the default
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 600 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "2048 256 30" "1024 512 50" "512 64 30"; do
  n=$(echo $g | tr ' ' x)
  ROMS_B200_S3T_VERBOSE=1 timeout 120 python tools/prof_step3d_t.py $g > ${O}_def_$n.log 2>&1
done
for v in $2; do
  name=${v%%:*}; envs=$(echo ${v#*:} | tr ',' ' ')
  env $envs ROMS_B200_S3T_VERBOSE=1 timeout 120 python tools/prof_step3d_t.py 2048 256 30 > ${O}_${name}.log 2>&1
done
if [ -z "$3" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step3d_t_v8 -s 4 -c 1 -o ${O}_v8_b3 python tools/prof_step3d_t.py 2048 256 30 > ${O}_ncu.log 2>&1
fi
for f in ${O}_*.log; do echo "$f: $(grep -h 'step3d_t v8' $f | tail -1 | cut -c13-120) | $(grep -h 'step3d_t ' $f | grep ms | tail -1)"; done
