#!/bin/bash
# tools/r2t.sh TAG -- GPU suite; per-kernel tables on both grids (level-major block order); step2d with 1/2/3 resident blocks per SM
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
python tools/time_phases.py 2048 256 30 5 > ${O}_phases_b3.log 2>&1; cat ${O}_phases_b3.log
python tools/time_phases.py > ${O}_phases_b1.log 2>&1; cat ${O}_phases_b1.log
for m in 1 3; do for g in "512 64 30 20" "2048 256 30 5"; do echo "MINB=$m $g: $(ROMS_B200_S2_MINB=$m python tools/time_phases.py $g 2>&1 | grep step2d_loop)"; done; done
timeout 900 python bench.py > ${O}_bench.log 2>&1
tail -1 ${O}_bench.log | cut -c1-300
