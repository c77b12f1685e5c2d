#!/bin/bash
# round-2 GPU pass h: single-launch E-step after the bookkeeping rewrite - parity, timings, one ncu capture
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "estep" 2>&1 | tail -15 > $OUT/r02h_tests.log
cat $OUT/r02h_tests.log
run() { tag=$1; shift; env "$@" timeout 300 python profiles/scripts/estep_bench.py $ARGS > $OUT/r02h_$tag.jsonl 2>> $OUT/r02h.err; echo "$tag $(cut -c90-200 $OUT/r02h_$tag.jsonl | tr '\n' ' ')"; }
ARGS="--alpha 0.01"
run 1m_3launch NFH_ESTEP_FUSED=0
run 1m_default X=1
run 1m_onewave NFH_ESTEP_WAVE_ROWS=100 NFH_ESTEP_HINTS=0
run 1m_mb24 NFH_ESTEP_WAVE_MB=24
run 1m_mb64 NFH_ESTEP_WAVE_MB=64
run 1m_mb40_nohints NFH_ESTEP_HINTS=0
run 1m_mb40_la0 NFH_ESTEP_LOOKAHEAD=0
ARGS="--n_ind 125 --n_sites 10000000 --alpha 0.01 --reps 5"
run 10m_3launch NFH_ESTEP_FUSED=0
run 10m_default X=1
run 10m_onewave NFH_ESTEP_WAVE_ROWS=125
run 10m_wr32 NFH_ESTEP_WAVE_ROWS=32
ncu --set full --clock-control none --import-source on -k regex:estep_fused -s 4 -c 1 -f -o $OUT/prof_estep_fused_r02h \
  python profiles/scripts/estep_bench.py --n_ind 40 --n_sites 1000000 --alpha 0.01 --reps 1 > $OUT/r02h_ncu.log 2>&1
ls -la $OUT/prof_estep_fused_r02h*
