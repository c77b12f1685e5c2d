#!/bin/bash
# round-2 GPU pass n: wider hybrid shapes (n_ind <= 1024), shape cost model, local posteriors for fixed frequencies
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q -k "freq or group" 2>&1 | tail -15 > $OUT/r02n_tests.log
cat $OUT/r02n_tests.log
python profiles/scripts/freq_bench.py --n_ind 125,128,900,960,992,1000,1024 --ind_sites 5e7 --reps 2 > $OUT/r02n_freq.jsonl 2> $OUT/r02n.err
python profiles/scripts/freq_bench.py --n_ind 1000 --ind_sites 1.25e9 --reps 1 >> $OUT/r02n_freq.jsonl 2>> $OUT/r02n.err
cut -c1-230 $OUT/r02n_freq.jsonl; tail -3 $OUT/r02n.err
