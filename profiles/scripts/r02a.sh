#!/bin/bash
# round-2 first GPU pass: parity tests, E-step micro-benchmark (new vs v1), ncu of the new kernels
OUT=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/r02a_tests.log
python profiles/scripts/estep_bench.py --alpha 0.01,0.2 --lkl > $OUT/r02a_estep_new.jsonl 2> $OUT/r02a_estep_new.err
NFH_ESTEP_V1=1 python profiles/scripts/estep_bench.py --alpha 0.01,0.2 --lkl > $OUT/r02a_estep_v1.jsonl 2> $OUT/r02a_estep_v1.err
python profiles/scripts/estep_bench.py --n_ind 12 --n_sites 10000000 --alpha 0.01,0.2 > $OUT/r02a_estep_10m.jsonl 2> $OUT/r02a_estep_10m.err
ncu --set full --clock-control none --import-source on -k regex:estep_chunk -s 6 -c 2 -f -o $OUT/prof_estep_r02a \
  python profiles/scripts/estep_bench.py --n_ind 40 --n_sites 500000 --alpha 0.01 --reps 1 > $OUT/r02a_ncu.log 2>&1
ls -la $OUT
