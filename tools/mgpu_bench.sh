#!/bin/bash
# tools/mgpu_bench.sh TAG NGPU [extra bench args] -- multi-GPU: tiling check tool + bench line on NGPU GPUs of one box
mkdir -p gpurun_out; O=gpurun_out/$1; N=$2; shift; shift
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
case $N in 2) TI="2 1";; 4) TI="2 2";; 8) TI="4 2";; esac
timeout 300 $T tools/mgpu_check.py --tiles $TI --grid 96 40 30 --steps 6 > ${O}_mgpu.log 2>&1; tail -2 ${O}_mgpu.log
timeout 900 $T bench.py --gpus $N --steps 50 --warmup 5 --no-roofline "$@" > ${O}_bench.log 2>&1
tail -1 ${O}_bench.log | cut -c1-2500
