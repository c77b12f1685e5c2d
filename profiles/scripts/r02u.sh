#!/bin/bash
# round-2 GPU pass u (2 GPUs): mixed exchange (posteriors by all-to-all behind the BFGS rounds, emission ratios by peer
# stores) - parity in all three modes, then configs[1] at 2 GPUs per mode with the phase trace
OUT=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu_check.py > $OUT/r02u_multi_gpu_check.log 2>&1
tail -4 $OUT/r02u_multi_gpu_check.log | cut -c1-260
for m in mixed direct nccl; do
  NFH_EXCHANGE=$m timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus 2 --steps 6 --warmup 3 --no_cpu_baseline --trace > $OUT/r02u_c1_2gpu_$m.json 2> $OUT/r02u_c1_2gpu_$m.err
  python - <<PY
import json
for l in open("gpurun_out/r02u_c1_2gpu_$m.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$m", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["kernel_ms_per_step"].items() if v}, d.get("parity_check", {}).get("ok"), d["rank_trace"]["per_rank"])
PY
done
