#!/bin/bash
# QP prediction passes: values of fixed variables carried from pass to pass (CLIK_QP_CARRY_XF) instead of a 12-row select chain per pass
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18), 2 streams"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'carried (cap 3, 168 regs):' 're-selected per pass (cap 4, 128 regs; before):CLIK_QP_CARRY_XF=0' 'carried, cap 4:CLIK_QP_FAST_MINBLOCKS=4' 're-selected, cap 3:CLIK_QP_CARRY_XF=0,CLIK_QP_FAST_MINBLOCKS=3'
echo "== ur5_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'carried (cap 3, 168 regs):' 're-selected per pass (cap 4, 128 regs; before):CLIK_QP_CARRY_XF=0' 'carried, cap 4:CLIK_QP_FAST_MINBLOCKS=4'
echo "== ur5_moe2016_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'carried:' 're-selected per pass (before):CLIK_QP_CARRY_XF=0'
} > gpurun_out/r2_ab16.txt 2>&1
cat gpurun_out/r2_ab16.txt | cut -c1-170
