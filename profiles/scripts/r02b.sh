#!/bin/bash
# round-2 second GPU pass: persistent double-buffered apply + 6-CTA products
OUT=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/r02b_tests.log
python profiles/scripts/estep_bench.py --alpha 0.01,0.2 --lkl > $OUT/r02b_estep_new.jsonl 2> $OUT/r02b_estep_new.err
python profiles/scripts/estep_bench.py --n_ind 125 --n_sites 10000000 --alpha 0.01,0.2 --reps 5 > $OUT/r02b_estep_10m.jsonl 2> $OUT/r02b_estep_10m.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:estep -s 9 -c 9 --csv --log-file $OUT/r02b_launches.csv \
  python profiles/scripts/estep_bench.py --alpha 0.01 --reps 3 > $OUT/r02b_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_chunk -s 6 -c 2 -f -o $OUT/prof_estep_r02b \
  python profiles/scripts/estep_bench.py --n_ind 40 --n_sites 500000 --alpha 0.01 --reps 1 > $OUT/r02b_ncu.log 2>&1
ls -la $OUT
