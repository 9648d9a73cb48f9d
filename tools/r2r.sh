#!/bin/bash
# tools/r2r.sh TAG -- where a sub-step of the persistent fast loop spends its time (clock stamps of the middle block)
mkdir -p gpurun_out; O=gpurun_out/$1
ROMS_B200_S2_PROF=1 python tools/time_phases.py > ${O}_prof.log 2>&1
grep -h "persistent\|step2d" ${O}_prof.log
