#!/bin/bash
# round-2 GPU pass q (8 GPUs): configs[2] with both exchange modes, configs[4] at 8/4/2 GPUs with local posteriors,
# configs[1] weak scaling at 8, the 2-GPU parity check, the drop-in binary on 2 GPUs
OUT=gpurun_out
run() {  # n_gpus config tag env...
  local n=$1 c=$2 tag=$3; shift 3
  env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $n --config $c --steps 3 --warmup 3 --no_cpu_baseline > $OUT/r02q_${tag}.json 2> $OUT/r02q_${tag}.err
  echo "$tag rc=$?"
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu_check.py > $OUT/r02q_multi_gpu_check.log 2>&1
tail -2 $OUT/r02q_multi_gpu_check.log
run 8 2 c2_8gpu_direct X=1
run 8 2 c2_8gpu_nccl NFH_PEER_DIRECT=0
run 8 4 c4_8gpu X=1
run 4 4 c4_4gpu X=1
run 2 4 c4_2gpu X=1
run 8 1 c1_8gpu X=1
timeout 600 python -m pytest tests/test_gpu_cli.py tests/test_gpu_multi.py -m gpu -x -q -k "n_gpus or multi" 2>&1 | tail -4
