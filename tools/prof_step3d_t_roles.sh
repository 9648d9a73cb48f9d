#!/bin/bash
# tools/prof_step3d_t_roles.sh [GRID] -- where the warps of step3d_t_v8_kernel spend their time: builds a copy of the library with
# -DS3T_PROF (clock64 counters per role and phase, never in the shipped library), runs tools/prof_step3d_t.py with it.
set -e
cd "$(dirname "$0")/.."
G=${1:-"2048 256 30"}
mkdir -p /tmp/s3tprof && cp -r roms_b200 include tools /tmp/s3tprof/ && cd /tmp/s3tprof/roms_b200/csrc
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr"
$NV -DS3T_PROF -c -o k_step3d_t8.o k_step3d_t8.cu
nvcc -shared -o ../libroms_b200.so *.o -lcudart -ldl 2>/dev/null
cd /tmp/s3tprof && python tools/prof_step3d_t.py $G 2>&1 | grep -E "S3T_PROF|step3d_t " | tail -3
