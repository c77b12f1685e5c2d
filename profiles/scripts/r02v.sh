#!/bin/bash
# round-2 GPU pass v (2 GPUs): why is the device-timed leg 1.5 ms per step slower than the end-to-end leg on 2 GPUs?
OUT=gpurun_out
for t in 1; do
  NFH_BENCH_NO_FAMILY_TIMING=$t NFH_EXCHANGE=mixed timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus 2 --steps 6 --warmup 3 --no_cpu_baseline --trace > $OUT/r02v_t$t.json 2> $OUT/r02v_t$t.err
  python - <<PY
import json
for l in open("gpurun_out/r02v_t$t.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("no_family_timing=$t dev", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "trace leg", round(d["rank_trace"]["wall_ms_per_step_of_this_leg"], 2), d["rank_trace"]["per_rank"][0])
PY
done
