#!/bin/bash
# tools/r2ae.sh TAG N -- N GPUs: halo kernel as a programmatic dependent launch against an ordinary launch
mkdir -p gpurun_out; O=gpurun_out/$1; N=$2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for e in "ROMS_B200_HALO_PDL=0" "ROMS_B200_HALO_PDL=1"; do
  env $e timeout 600 $T bench.py --gpus $N --steps 50 --warmup 5 --no-roofline --no-cpu --no-also > ${O}_$e.log 2>&1
  echo "$e: $(tail -1 ${O}_$e.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('tiling_bit_identical'))" 2>&1 | tail -1)"
done
