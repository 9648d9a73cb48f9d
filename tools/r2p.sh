#!/bin/bash
# tools/r2p.sh TAG -- persistent fast loop: parity tests that cover step2d, BENCHMARK1 phase table with and without it, bench
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 900 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log
tail -3 ${O}_pytest.log
python tools/time_phases.py > ${O}_phases_b1.log 2>&1
ROMS_B200_S2_PERSIST=0 python tools/time_phases.py > ${O}_phases_b1_graph.log 2>&1
grep -h "step2d" ${O}_phases_b1.log ${O}_phases_b1_graph.log
timeout 900 python bench.py > ${O}_bench.log 2>&1
tail -1 ${O}_bench.log | cut -c1-900
