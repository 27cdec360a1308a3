#!/bin/bash
# QP prediction: a released variable may go straight to another bound in the first passes (CLIK_QP_FLIP_PASSES) x fast-pass budget
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18), 2 streams"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'flip 2, fast passes 4 (default):' 'flip 2, fast passes 3:CLIK_QP_FAST_PASSES=3' 'flip 2, fast passes 5:CLIK_QP_FAST_PASSES=5' 'flip 0, fast passes 4 (before):CLIK_QP_FLIP_PASSES=0' 'flip 2, fast 3, 1 stream plain:CLIK_QP_FAST_PASSES=3,CLIK_BENCH_STREAMS=1' 'flip 0, fast 4, 1 stream plain:CLIK_QP_FLIP_PASSES=0,CLIK_BENCH_STREAMS=1'
echo "== ur5_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'flip 2, fast passes 4 (default):' 'flip 2, fast passes 3:CLIK_QP_FAST_PASSES=3' 'flip 0, fast passes 4 (before):CLIK_QP_FLIP_PASSES=0'
echo "== ur5_moe2016_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'flip 2, fast passes 4 (default):' 'flip 2, fast passes 3:CLIK_QP_FAST_PASSES=3' 'flip 2, fast passes 2:CLIK_QP_FAST_PASSES=2' 'flip 0, fast passes 4 (before):CLIK_QP_FLIP_PASSES=0'
echo "== ur5_moe2016_qp (2^23), 2 streams"
TUNE_STEPS=20 python tools/tune.py ur5_moe2016_qp 8388608 'flip 2, fast passes 4 (default):' 'flip 2, fast passes 3:CLIK_QP_FAST_PASSES=3' 'flip 2, fast passes 2:CLIK_QP_FAST_PASSES=2'
} > gpurun_out/r2_ab11.txt 2>&1
cat gpurun_out/r2_ab11.txt | cut -c1-110
