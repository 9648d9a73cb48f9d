#!/bin/bash
# tools/r2ad.sh TAG -- 2 GPUs: programmatic dependent launch of the sub-steps with neighbours (ROMS_B200_PDL=1) against the default
mkdir -p gpurun_out; O=gpurun_out/$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for e in "X=1" "ROMS_B200_PDL=1"; do
  env $e timeout 600 $T bench.py --gpus 2 --steps 50 --warmup 5 --no-roofline --no-cpu > ${O}_$e.log 2>&1
  echo "$e: $(tail -1 ${O}_$e.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('tiling_bit_identical'))")"
done
