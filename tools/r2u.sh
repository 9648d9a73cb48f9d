#!/bin/bash
# tools/r2u.sh TAG -- rolling-window t3dmix2_geo: GPU parity suite, then its time for 2/3/4 resident blocks x fill factors, both grids
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "512 64 30 20" "2048 256 30 5"; do
  for m in 2 3; do for f in 2 4 8; do
    echo "MINB=$m FILL=$f $g: $(ROMS_B200_T3DMIX_MINB=$m ROMS_B200_T3DMIX_FILL=$f python tools/time_phases.py $g 2>&1 | grep t3dmix2)"
  done; done
done
