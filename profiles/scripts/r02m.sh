#!/bin/bash
# round-2 GPU pass m: frequency EM at the shapes of configs[2] on 8 GPUs (1000 individuals x 1.25M sites per rank) - long rows
OUT=gpurun_out
python profiles/scripts/freq_bench.py --n_ind 1000 --ind_sites 1.25e9 --reps 1 > $OUT/r02m_freq_long.jsonl 2> $OUT/r02m.err
python profiles/scripts/freq_bench.py --n_ind 1000 --ind_sites 2.5e8 --reps 1 >> $OUT/r02m_freq_long.jsonl 2>> $OUT/r02m.err
python profiles/scripts/freq_bench.py --n_ind 800 --ind_sites 8e8 --reps 1 >> $OUT/r02m_freq_long.jsonl 2>> $OUT/r02m.err
NFH_FREQ_G=16 python profiles/scripts/freq_bench.py --n_ind 125,128 --ind_sites 5e7 --reps 2 >> $OUT/r02m_freq_long.jsonl 2>> $OUT/r02m.err
cut -c1-230 $OUT/r02m_freq_long.jsonl; tail -3 $OUT/r02m.err
