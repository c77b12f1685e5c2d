timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r01f_full.json 2> gpurun_out/bench_r01f.err; tail -c 600 gpurun_out/bench_r01f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01f_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['kernel_ms_per_step'])
print(d['roofline'])
print(d['e2e'], d['cpu_baseline'], d['clocks'])
PY
