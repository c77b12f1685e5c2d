#!/bin/bash
# round-2 GPU pass s: hybrid kernel with 16-byte coefficient pairs and L2 prefetch of the next tile
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "freq" 2>&1 | tail -5 > $OUT/r02s_tests.log
cat $OUT/r02s_tests.log
python profiles/scripts/freq_bench.py --n_ind 600,800,832,1000 --ind_sites 5e7 --reps 2 > $OUT/r02s_freq.jsonl 2> $OUT/r02s.err
python profiles/scripts/freq_bench.py --n_ind 800 --ind_sites 8e8 --reps 1 >> $OUT/r02s_freq.jsonl 2>> $OUT/r02s.err
python profiles/scripts/freq_bench.py --n_ind 1000 --ind_sites 1.25e9 --reps 1 >> $OUT/r02s_freq.jsonl 2>> $OUT/r02s.err
cut -c1-200 $OUT/r02s_freq.jsonl; tail -3 $OUT/r02s.err
