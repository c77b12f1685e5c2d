#!/bin/bash
# round-2 GPU pass t (8 GPUs): per-rank phase trace of configs[1] weak scaling at 8 GPUs, fused and NCCL exchange
OUT=gpurun_out
run() {
  local n=$1 c=$2 tag=$3; shift 3
  env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $n --config $c --steps 6 --warmup 3 --no_cpu_baseline --trace > $OUT/r02t_${tag}.json 2> $OUT/r02t_${tag}.err
  echo "$tag rc=$?"
}
run 8 1 c1_8gpu_direct X=1
run 8 1 c1_8gpu_nccl NFH_PEER_DIRECT=0
python - <<'PY'
import json
for t in ("c1_8gpu_direct", "c1_8gpu_nccl"):
    for l in open(f"gpurun_out/r02t_{t}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(t, round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["kernel_ms_per_step"].items() if v})
            for r in d["rank_trace"]["per_rank"]:
                print("   ", r)
PY
