#!/bin/bash
# round-2 GPU pass i: ncu --set full of the three-launch E-step kernels and of the single-launch kernel
OUT=gpurun_out
NFH_ESTEP_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:estep_chunk -s 6 -c 2 -f -o $OUT/prof_estep_3launch_r02i \
  python profiles/scripts/estep_bench.py --n_ind 100 --n_sites 1000000 --alpha 0.01 --reps 1 > $OUT/r02i_ncu_a.log 2>&1
NFH_ESTEP_WAVE_ROWS=100 NFH_ESTEP_HINTS=0 ncu --set full --clock-control none --import-source on -k regex:estep_fused -s 3 -c 1 -f -o $OUT/prof_estep_fused_r02i \
  python profiles/scripts/estep_bench.py --n_ind 100 --n_sites 1000000 --alpha 0.01 --reps 1 > $OUT/r02i_ncu_b.log 2>&1
ls -la $OUT/prof_estep_*r02i*
