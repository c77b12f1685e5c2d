#!/bin/bash
# round-2 GPU pass o: whole GPU suite + default bench after the shape model / wide hybrid / local-posterior changes
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/r02o_tests.log
cat $OUT/r02o_tests.log
timeout 600 python bench.py > $OUT/r02o_bench.json 2> $OUT/r02o_bench.err
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 --no_cpu_baseline > $OUT/r02o_bench_c2.json 2> $OUT/r02o_bench_c2.err
python - <<'PY'
import json
for f in ("r02o_bench", "r02o_bench_c2"):
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["roofline_estep"]["frac"], d["e2e"]["value"])
PY
