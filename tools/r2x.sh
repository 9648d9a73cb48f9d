#!/bin/bash
# tools/r2x.sh TAG -- marching uv3dmix2: GPU parity suite, timings (fill factors) on both grids against the per-level form
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "512 64 30 20" "2048 256 30 5"; do
  echo "per-level $g: $(ROMS_B200_UVMIX_PERLEVEL=1 python tools/time_phases.py $g 2>&1 | grep -E 'uv3dmix2' | tr '\n' ' ')"
  for f in 1 2 4; do
    echo "FILL=$f $g: $(ROMS_B200_UVMIX_FILL=$f python tools/time_phases.py $g 2>&1 | grep -E 'uv3dmix2' | tr '\n' ' ')"
  done
done
timeout 900 python bench.py --no-cpu --no-roofline > ${O}_bench.log 2>&1
tail -1 ${O}_bench.log | cut -c1-300
