#!/bin/bash
# round-2 GPU pass p: products kernel with two threads per chunk; reference patched per INTEGRATION.md section B
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_cli.py -m gpu -x -q -k "estep or patched or golden or fixture or full_size or long_sequence or tiny" 2>&1 | tail -15 > $OUT/r02p_tests.log
cat $OUT/r02p_tests.log
run() { tag=$1; shift; env "$@" timeout 300 python profiles/scripts/estep_bench.py $ARGS > $OUT/r02p_$tag.jsonl 2>> $OUT/r02p.err; echo "$tag $(cut -c100-215 $OUT/r02p_$tag.jsonl | tr '\n' ' ')"; }
ARGS="--alpha 0.01,0.2"
run 1m_prod1 NFH_ESTEP_PROD2=0
run 1m_prod2 NFH_ESTEP_PROD2=1
ARGS="--n_ind 125 --n_sites 10000000 --alpha 0.01,0.2 --reps 5"
run 10m_prod1 NFH_ESTEP_PROD2=0
run 10m_prod2 NFH_ESTEP_PROD2=1
ARGS="--n_ind 1250 --n_sites 1000000 --alpha 0.01 --reps 5"
run c4_prod1 NFH_ESTEP_PROD2=0
run c4_prod2 NFH_ESTEP_PROD2=1
