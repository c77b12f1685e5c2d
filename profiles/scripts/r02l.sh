#!/bin/bash
# round-2 GPU pass l: frequency EM against the number of individuals per site (shape selection)
OUT=gpurun_out
python profiles/scripts/freq_bench.py --n_ind 100,104,125,128,200,208,400,416,512,800,832,1000,1250 --ind_sites 5e7 --reps 2 > $OUT/r02l_freq_sweep.jsonl 2> $OUT/r02l.err
NFH_FREQ_G=8 python profiles/scripts/freq_bench.py --n_ind 125,128 --ind_sites 5e7 --reps 2 >> $OUT/r02l_freq_sweep.jsonl 2>> $OUT/r02l.err
NFH_FREQ_G=32 python profiles/scripts/freq_bench.py --n_ind 125,200,400 --ind_sites 5e7 --reps 2 >> $OUT/r02l_freq_sweep.jsonl 2>> $OUT/r02l.err
NFH_FREQ_NO_HYBRID=1 python profiles/scripts/freq_bench.py --n_ind 800 --ind_sites 5e7 --reps 2 >> $OUT/r02l_freq_sweep.jsonl 2>> $OUT/r02l.err
cut -c1-230 $OUT/r02l_freq_sweep.jsonl; tail -3 $OUT/r02l.err
