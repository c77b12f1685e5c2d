#!/bin/bash
# round-2 GPU pass k: does the second sweep of a wave hit L2?  DRAM bytes of estep_fused against the wave size
OUT=gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
for cfg in "8 0" "8 200" "16 200" "24 200" "32 200" "40 200" "40 888" "64 888" "16 888"; do
  set -- $cfg
  NFH_ESTEP_WAVE_MB=$1 NFH_ESTEP_LOOKAHEAD=$2 NFH_ESTEP_HINTS=1 ncu --metrics $M --clock-control none -k regex:estep_fused -s 3 -c 1 --csv \
    python profiles/scripts/estep_bench.py --n_ind 100 --n_sites 1000000 --alpha 0.01 --reps 1 2>/dev/null | grep estep_fused | awk -F'","' -v c="mb=$1 la=$2" '{print c, $(NF-2), $(NF)}' | tr -d '"' >> $OUT/r02k_sweep.txt
done
NFH_ESTEP_WAVE_MB=16 NFH_ESTEP_LOOKAHEAD=200 NFH_ESTEP_HINTS=0 ncu --metrics $M --clock-control none -k regex:estep_fused -s 3 -c 1 --csv \
    python profiles/scripts/estep_bench.py --n_ind 100 --n_sites 1000000 --alpha 0.01 --reps 1 2>/dev/null | grep estep_fused | awk -F'","' -v c="mb=16 la=200 nohints" '{print c, $(NF-2), $(NF)}' | tr -d '"' >> $OUT/r02k_sweep.txt
cat $OUT/r02k_sweep.txt
