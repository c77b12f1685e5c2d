#!/bin/bash
# round-2 GPU pass d: all GPU tests (group driver, CLI --n_gpus / --call_geno), bench config 1 and small smokes of configs 2 / 4
OUT=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/r02d_tests.log
python bench.py > $OUT/r02d_bench.json 2> $OUT/r02d_bench.err
python bench.py --config 2 --n_ind 8 --n_sites 2000000 --steps 2 --warmup 1 --no_cpu_baseline > $OUT/r02d_bench_c2smoke.json 2> $OUT/r02d_bench_c2smoke.err
python bench.py --config 4 --n_ind 300 --n_sites 200000 --steps 2 --warmup 1 --cpu_sites 2000 > $OUT/r02d_bench_c4smoke.json 2> $OUT/r02d_bench_c4smoke.err
ls -la $OUT | tail -8
