#!/bin/bash
# tools/r2ac.sh TAG -- step3d_t v8: fewer, fatter producer warps (4 levels per thread)
mkdir -p gpurun_out; O=gpurun_out/$1
for cfg in "" "ROMS_B200_S3T_NP=4" "ROMS_B200_S3T_NP=4 ROMS_B200_S3T_NC=2" "ROMS_B200_S3T_NP=4 ROMS_B200_S3T_NC=4" "ROMS_B200_S3T_NP=5" "ROMS_B200_S3T_NP=6"; do
  echo "[$cfg] $(env $cfg ROMS_B200_S3T_VERBOSE=1 timeout 120 python tools/prof_step3d_t.py 2048 256 30 2>&1 | grep -E 'ms/launch|v8:' | tail -2 | tr '\n' ' ' | cut -c1-260)"
done
