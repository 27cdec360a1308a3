#!/bin/bash
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18): prediction passes in the fast launch (the tail continues the rest of the 12)"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'fast passes 4 (default):' 'fast passes 3:CLIK_QP_FAST_PASSES=3' 'fast passes 5:CLIK_QP_FAST_PASSES=5' 'fast passes 6:CLIK_QP_FAST_PASSES=6' 'fast passes 12 (round-1 behaviour):CLIK_QP_FAST_PASSES=12' 'fast passes 2:CLIK_QP_FAST_PASSES=2'
echo "== ur5_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'fast passes 4 (default):' 'fast passes 3:CLIK_QP_FAST_PASSES=3' 'fast passes 12 (round-1 behaviour):CLIK_QP_FAST_PASSES=12'
echo "== ur5_moe2016_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'fast passes 4 (default):' 'fast passes 3:CLIK_QP_FAST_PASSES=3' 'fast passes 5:CLIK_QP_FAST_PASSES=5' 'fast passes 12 (round-1 behaviour):CLIK_QP_FAST_PASSES=12'
echo "== ur5_track / moe pinv with the occupancy rule"
python tools/tune.py ur5_track 1048576 'default:'
python tools/tune.py ur5_moe2016_pinv 1048576 'default:'
python tools/tune.py iiwa_multitask 1048576 'default:'
} > gpurun_out/r2_ab4.txt 2>&1
cat gpurun_out/r2_ab4.txt | cut -c1-120
