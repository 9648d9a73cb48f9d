#!/bin/bash
# tools/r2_final.sh TAG -- what the driver runs at round end on one GPU: GPU suite, smoke(), bench.py (both arms)
mkdir -p gpurun_out; O=gpurun_out/$1
timeout 1200 python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -2 ${O}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > ${O}_smoke.log 2>&1; tail -1 ${O}_smoke.log
( time timeout 900 python bench.py > ${O}_bench.log 2>&1 ) 2>&1 | grep real
tail -1 ${O}_bench.log | cut -c1-260
