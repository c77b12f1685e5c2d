#!/bin/bash
# round-2 GPU pass r: launch list + ncu --set full of the hot kernels (run_ncu.sh, tag r02) and of the wide hybrid kernel
# at 1,000 individuals per site; the reports are summarised on the box (they exceed what gpurun copies back)
OUT=gpurun_out
KERNELS="freq_emission_warp lkl_tile_products estep_chunk_apply estep_chunk_products" bash profiles/run_ncu.sh r02 > $OUT/r02r_run_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:freq_emission_hybrid -s 1 -c 1 -f -o $OUT/prof_freq_emission_hybrid_r02 \
  python profiles/scripts/freq_bench.py --n_ind 1000 --ind_sites 3e7 --reps 1 > $OUT/r02r_ncu_hybrid.log 2>&1
for k in freq_emission_warp lkl_tile_products estep_chunk_apply estep_chunk_products freq_emission_hybrid; do
  python profiles/ncu_rows.py $OUT/prof_${k}_r02.ncu-rep $OUT/ncu_${k}_r02_summary.txt > /dev/null 2>&1
  python profiles/ncu_hot.py $OUT/prof_${k}_r02.ncu-rep $k 25 > $OUT/ncu_${k}_r02_hot.txt 2>&1
done
rm -f $OUT/prof_freq_emission_warp_r02.ncu-rep $OUT/prof_freq_emission_hybrid_r02.ncu-rep $OUT/prof_lkl_tile_products_r02.ncu-rep
ls -la $OUT | tail -20
