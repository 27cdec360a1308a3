#!/bin/bash
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18): slow instances continued inside the fast kernel's CTA (no waiting tail launch)"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'default (4 passes, then in-CTA continuation; cap 168 regs):' 'fast passes 3:CLIK_QP_FAST_PASSES=3' 'fast passes 5:CLIK_QP_FAST_PASSES=5' 'fast passes 6:CLIK_QP_FAST_PASSES=6' 'no register cap (208 regs):CLIK_QP_FAST_MINBLOCKS=0' 'one-row passes off:CLIK_QP_CRASH_SINGLE=0'
echo "== ur5_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'default:' 'fast passes 3:CLIK_QP_FAST_PASSES=3' 'fast passes 5:CLIK_QP_FAST_PASSES=5'
echo "== ur5_moe2016_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'default:' 'fast passes 3:CLIK_QP_FAST_PASSES=3' 'fast passes 6:CLIK_QP_FAST_PASSES=6' 'no register cap:CLIK_QP_FAST_MINBLOCKS=0' 'one-row passes off:CLIK_QP_CRASH_SINGLE=0'
} > gpurun_out/r2_ab6.txt 2>&1
cat gpurun_out/r2_ab6.txt | cut -c1-130
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:clik_qp -s 6 -c 8 --csv --log-file gpurun_out/r2_qp_launches.csv python bench.py --secondary-only ur5_qp > /dev/null 2>&1
grep clik_qp gpurun_out/r2_qp_launches.csv | awk -F'","' '{print $5, $(NF)}' | head -8
