timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "freq or emission or tiny" 2>&1 | tail -3
for cfg in "--n_ind 100 --n_sites 1000000" "--n_ind 200 --n_sites 500000" "--n_ind 400 --n_sites 250000" "--n_ind 800 --n_sites 125000"; do
  echo "== $cfg"
  timeout 600 python bench.py $cfg --steps 3 --warmup 2 --no_cpu_baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['kernel_ms_per_step'])"
done
