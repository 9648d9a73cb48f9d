#!/bin/bash
# tools/mgpu_ab.sh TAG N -- N GPUs: halo strips chunked over blocks against one block per strip
mkdir -p gpurun_out; O=gpurun_out/$1; N=$2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for e in $EVS; do
  env $e timeout 600 $T bench.py --gpus $N --steps 50 --warmup 5 --no-roofline --no-cpu --no-also > ${O}_$e.log 2>&1
  echo "$e: $(tail -1 ${O}_$e.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('tiling_bit_identical'))" 2>&1 | tail -1)"
done
