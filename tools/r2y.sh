#!/bin/bash
# tools/r2y.sh TAG -- marching rhs3d (and uv3dmix2 full-column on the small grid): parity suite, timings
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "512 64 30 20" "2048 256 30 5"; do
  echo "per-level $g: $(ROMS_B200_RHS3D_PERLEVEL=1 python tools/time_phases.py $g 2>&1 | grep -E 'rhs3d' | tr '\n' ' ')"
  for f in 0 1 2; do
    echo "FILL=$f $g: $(ROMS_B200_RHS3D_FILL=$f ROMS_B200_UVMIX_FILL=$f python tools/time_phases.py $g 2>&1 | grep -E 'rhs3d|uv3dmix2' | tr '\n' ' ')"
  done
done
timeout 900 python bench.py --no-cpu --no-roofline > ${O}_bench.log 2>&1
tail -1 ${O}_bench.log | cut -c1-300
