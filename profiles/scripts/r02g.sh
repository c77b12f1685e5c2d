#!/bin/bash
# round-2 GPU pass g: single-launch E-step (estep_fused) - parity tests under a timeout, then timings
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "estep" 2>&1 | tail -15 > $OUT/r02g_tests.log
cat $OUT/r02g_tests.log
for mode in 1 0; do
  NFH_ESTEP_FUSED=$mode timeout 300 python profiles/scripts/estep_bench.py --alpha 0.01,0.2 > $OUT/r02g_estep_1m_fused$mode.jsonl 2> $OUT/r02g_estep_1m_fused$mode.err
  NFH_ESTEP_FUSED=$mode timeout 300 python profiles/scripts/estep_bench.py --n_ind 125 --n_sites 10000000 --alpha 0.01 --reps 5 > $OUT/r02g_estep_10m_fused$mode.jsonl 2> $OUT/r02g_estep_10m_fused$mode.err
done
for mb in 16 24 32 48 64 96; do
  NFH_ESTEP_WAVE_MB=$mb timeout 300 python profiles/scripts/estep_bench.py --alpha 0.01 > $OUT/r02g_estep_1m_mb$mb.jsonl 2>> $OUT/r02g_sweep.err
done
NFH_ESTEP_HINTS=0 timeout 300 python profiles/scripts/estep_bench.py --alpha 0.01 > $OUT/r02g_estep_1m_nohints.jsonl 2>> $OUT/r02g_sweep.err
for wr in 4 8 32 125; do
  NFH_ESTEP_WAVE_ROWS=$wr timeout 300 python profiles/scripts/estep_bench.py --n_ind 125 --n_sites 10000000 --alpha 0.01 --reps 5 > $OUT/r02g_estep_10m_wr$wr.jsonl 2>> $OUT/r02g_sweep.err
done
for f in $OUT/r02g_estep_*.jsonl; do echo "$f $(cut -c1-180 $f | tr "\n" " ")"; done
