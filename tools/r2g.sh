#!/bin/bash
# tools/r2g.sh TAG -- GPU suite, step3d_t timings, bench line (N=1)
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
for g in "2048 256 30" "1024 512 50" "512 64 30"; do
  n=$(echo $g | tr ' ' x)
  ROMS_B200_S3T_VERBOSE=1 timeout 120 python tools/prof_step3d_t.py $g > ${O}_s3t_$n.log 2>&1
done
timeout 900 python bench.py > ${O}_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > ${O}_bench_ref.log 2>&1
python tools/time_phases.py > ${O}_phases.log 2>&1
for f in ${O}_s3t_*.log; do echo "$f: $(grep -h 'step3d_t ' $f | grep ms | tail -1)"; done
tail -1 ${O}_bench.log | cut -c1-3000; tail -1 ${O}_bench_ref.log | cut -c1-400; cat ${O}_phases.log
