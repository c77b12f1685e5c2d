#!/bin/bash
# round-2 GPU pass w (8 GPUs): configs[1] weak scaling at 8 and 4 GPUs with the mixed exchange (default), direct for
# comparison; configs[2] at 8 GPUs (auto = NCCL)
OUT=gpurun_out
run() {
  local n=$1 c=$2 tag=$3; shift 3
  env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $n --config $c --no_cpu_baseline --trace > $OUT/r02w_${tag}.json 2> $OUT/r02w_${tag}.err
  echo "$tag rc=$?"
}
run 8 1 c1_8gpu_mixed X=1
run 8 1 c1_8gpu_direct NFH_EXCHANGE=direct
run 4 1 c1_4gpu_mixed X=1
run 2 1 c1_2gpu_mixed X=1
python - <<'PY'
import json
for t in ("c1_8gpu_mixed", "c1_8gpu_direct", "c1_4gpu_mixed", "c1_2gpu_mixed"):
    for l in open(f"gpurun_out/r02w_{t}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(t, "dev", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), {k: round(v, 2) for k, v in d["kernel_ms_per_step"].items() if v}, d["parity_check"]["ok"])
            for r in d["rank_trace"]["per_rank"]:
                print("   ", r)
PY
