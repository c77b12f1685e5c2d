#!/bin/bash
# round-2 GPU pass e: full per-GPU sizes of configs[2] and configs[4] on ONE GPU (memory / time check before the 8-GPU
# run) and the stand-alone E-step at S = 10M with the current kernels
OUT=gpurun_out
python profiles/scripts/estep_bench.py --alpha 0.01,0.2 --lkl > $OUT/r02e_estep_1m.jsonl 2> $OUT/r02e_estep_1m.err
python profiles/scripts/estep_bench.py --n_ind 125 --n_sites 10000000 --alpha 0.01,0.2 --reps 5 > $OUT/r02e_estep_10m.jsonl 2> $OUT/r02e_estep_10m.err
timeout 900 python bench.py --config 2 --steps 3 --warmup 3 --no_cpu_baseline > $OUT/r02e_bench_c2.json 2> $OUT/r02e_bench_c2.err
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no_cpu_baseline > $OUT/r02e_bench_c4.json 2> $OUT/r02e_bench_c4.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $OUT/r02e_mem.txt
ls -la $OUT | tail -8
