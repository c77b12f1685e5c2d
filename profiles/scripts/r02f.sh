#!/bin/bash
# round-2 GPU pass f (8 GPUs): BASELINE configs[2] (1000 x 10M over 8 GPUs, full EM step) and configs[4]
# (10,000 x 1M at 8 GPUs; 5,000 at 4; 2,500 at 2; fixed parameters, E-step + Viterbi)
OUT=gpurun_out
run() {  # n_gpus config tag extra...
  local n=$1 c=$2 tag=$3; shift 3
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n + c)) \
    bench.py --gpus $n --config $c --steps 3 --warmup 3 --no_cpu_baseline "$@" > $OUT/r02f_${tag}.json 2> $OUT/r02f_${tag}.err
  echo "$tag rc=$?"
}
run 8 2 c2_8gpu
run 8 4 c4_8gpu
run 4 4 c4_4gpu
run 2 4 c4_2gpu
ls -la $OUT | tail -10
