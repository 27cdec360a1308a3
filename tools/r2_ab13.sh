#!/bin/bash
# QP prediction passes: structural zeros of the dense rows skipped at compile time (CLIK_QP_ADNZ=0: every entry)
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18), 2 streams"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'zeros skipped (default):' 'every entry (before):CLIK_QP_ADNZ=0'
echo "== ur5_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'zeros skipped (default):' 'every entry (before):CLIK_QP_ADNZ=0'
echo "== ur5_moe2016_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'zeros skipped (default):' 'every entry (before):CLIK_QP_ADNZ=0'
echo "== ur5_moe2016_qp (2^23), 2 streams"
TUNE_STEPS=20 python tools/tune.py ur5_moe2016_qp 8388608 'zeros skipped (default):' 'every entry (before):CLIK_QP_ADNZ=0'
} > gpurun_out/r2_ab13.txt 2>&1
cat gpurun_out/r2_ab13.txt | cut -c1-110
