#!/bin/bash
# tools/r2aa.sh TAG -- diag on the second stream, KPP levels in chunks: parity suite, phase table, bench
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
python tools/time_phases.py > ${O}_phases_b1.log 2>&1; grep -E "lmd_vmix|sum" ${O}_phases_b1.log
python tools/time_phases.py 2048 256 30 5 > ${O}_phases_b3.log 2>&1; grep -E "lmd_vmix|sum" ${O}_phases_b3.log
timeout 900 python bench.py --no-cpu --no-roofline > ${O}_bench.log 2>&1
tail -1 ${O}_bench.log | cut -c1-300
