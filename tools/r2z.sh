#!/bin/bash
# tools/r2z.sh TAG -- marching pre_step3d momentum: parity suite, timings, bench
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "512 64 30 20" "2048 256 30 5"; do
  for f in 0 2 4 8; do
    echo "FILL=$f $g: $(ROMS_B200_PRE3DUV_FILL=$f python tools/time_phases.py $g 2>&1 | grep -E 'pre_step3d|rhs3d' | tr '\n' ' ')"
  done
done
timeout 900 python bench.py --no-cpu --no-roofline > ${O}_bench.log 2>&1
tail -1 ${O}_bench.log | cut -c1-300
