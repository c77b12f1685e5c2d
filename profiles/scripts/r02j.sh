#!/bin/bash
# round-2 GPU pass j: single-launch E-step v3 (batched tickets, finalisation off the feeding warp, early readiness)
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "estep" 2>&1 | tail -15 > $OUT/r02j_tests.log
cat $OUT/r02j_tests.log
run() { tag=$1; shift; env "$@" timeout 300 python profiles/scripts/estep_bench.py $ARGS > $OUT/r02j_$tag.jsonl 2>> $OUT/r02j.err; echo "$tag $(cut -c90-200 $OUT/r02j_$tag.jsonl | tr '\n' ' ')"; }
ARGS="--alpha 0.01"
run 1m_3launch NFH_ESTEP_FUSED=0
run 1m_default X=1
run 1m_onewave NFH_ESTEP_WAVE_ROWS=100 NFH_ESTEP_HINTS=0
run 1m_mb24 NFH_ESTEP_WAVE_MB=24
run 1m_mb64 NFH_ESTEP_WAVE_MB=64
run 1m_mb40_nohints NFH_ESTEP_HINTS=0
run 1m_mb40_la0 NFH_ESTEP_LOOKAHEAD=0
run 1m_mb40_la444 NFH_ESTEP_LOOKAHEAD=444
ARGS="--n_ind 125 --n_sites 10000000 --alpha 0.01 --reps 5"
run 10m_default X=1
run 10m_onewave NFH_ESTEP_WAVE_ROWS=125
NFH_ESTEP_WAVE_ROWS=100 NFH_ESTEP_HINTS=0 ncu --set full --clock-control none --import-source on -k regex:estep_fused -s 3 -c 1 -f -o $OUT/prof_estep_fused_r02j \
  python profiles/scripts/estep_bench.py --n_ind 100 --n_sites 1000000 --alpha 0.01 --reps 1 > $OUT/r02j_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_fused -s 3 -c 1 -f -o $OUT/prof_estep_fused_waves_r02j \
  python profiles/scripts/estep_bench.py --n_ind 100 --n_sites 1000000 --alpha 0.01 --reps 1 > $OUT/r02j_ncu_c.log 2>&1
