#!/bin/bash
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18): one-row-per-pass passes before the Goldfarb-Idnani iteration (tail latency)"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'single-change passes on (default):' 'off:CLIK_QP_CRASH_SINGLE=0' 'on, fast passes 12:CLIK_QP_FAST_PASSES=12' 'on, fast passes 6:CLIK_QP_FAST_PASSES=6'
echo "== ur5_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'single-change passes on (default):' 'off:CLIK_QP_CRASH_SINGLE=0'
echo "== ur5_moe2016_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'single-change passes on (default):' 'off:CLIK_QP_CRASH_SINGLE=0'
} > gpurun_out/r2_ab5.txt 2>&1
cat gpurun_out/r2_ab5.txt | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:clik_qp -s 6 -c 8 --csv --log-file gpurun_out/r2_qp_launches.csv python bench.py --secondary-only ur5_qp > /dev/null 2>&1
grep clik_qp gpurun_out/r2_qp_launches.csv | cut -d, -f5,15 | head -10
