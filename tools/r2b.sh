#!/bin/bash
# tools/r2b.sh -- hardware run of the TMA/mbarrier step3d_t (k_step3d_t8.cu): parity suite, timings, ncu capture
mkdir -p gpurun_out; O=gpurun_out/${1:-r2b}
timeout 600 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "2048 256 30" "1024 512 50" "512 64 30"; do
  n=$(echo $g | tr ' ' x)
  ROMS_B200_S3T_VERBOSE=1 timeout 120 python tools/prof_step3d_t.py $g > ${O}_v8_$n.log 2>&1
done
for j in 8 16; do ROMS_B200_S3T_JCH=$j timeout 120 python tools/prof_step3d_t.py 2048 256 30 > ${O}_v8_jch$j.log 2>&1; done
for s in $2; do ROMS_B200_S3T_SLOTS=$s timeout 120 python tools/prof_step3d_t.py 2048 256 30 > ${O}_v8_slots$s.log 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step3d_t_v8 -s 4 -c 1 -o ${O}_v8_b3 python tools/prof_step3d_t.py 2048 256 30 > ${O}_ncu.log 2>&1
for f in ${O}_v*.log; do echo "$f: $(grep -h 'step3d_t ' $f | tail -1)"; done
