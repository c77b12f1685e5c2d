#!/bin/bash
# Profiling passes run on the GPU box (B200_PROFILING.md recipe).  Usage: profiles/run_ncu.sh <tag>
# 1) launch list of the bench command (every launch + device time)
# 2) one --set full capture of each hot kernel
TAG=${1:-r01}
OUT=gpurun_out
SITES=${SITES:-400000}
CMD="python bench.py --n_sites $SITES --steps 1 --warmup 1 --no_cpu_baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
for K in ${KERNELS:-freq_emission_warp estep_chunk_apply estep_chunk_products lkl_tile_products viterbi_chunk_pointers}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $OUT/prof_${K}_$TAG $CMD > $OUT/ncu_${K}_$TAG.log 2>&1
done
ls -la $OUT
