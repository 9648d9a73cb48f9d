#!/bin/bash
# tools/r2ag.sh TAG -- step3d_t v8 on the N=50 tile with 3 slots (forced) against the v6 fallback
mkdir -p gpurun_out; O=gpurun_out/$1
for cfg in "X=0" "ROMS_B200_S3T_SLOTS=3" "ROMS_B200_S3T_SLOTS=3 ROMS_B200_S3T_NC=2" "ROMS_B200_S3T_SLOTS=3 ROMS_B200_S3T_TMEM=1" "ROMS_B200_S3T_SLOTS=2"; do
  echo "[$cfg] $(env $cfg ROMS_B200_S3T_VERBOSE=1 timeout 200 python tools/prof_step3d_t.py 1024 512 50 2>&1 | grep -E 'ms/launch|v8:|rror' | tail -2 | tr '\n' ' ' | cut -c1-300)"
done
