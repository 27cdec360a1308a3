#!/bin/bash
# programmatic dependent launch (CLIK_PDL): 0 plain stream order, 1 the two launches of one step overlap, 2 steps overlap
mkdir -p gpurun_out
{
echo "== ur5_track (2^20)"
TUNE_STEPS=200 python tools/tune.py ur5_track 1048576 'PDL 1 (default):' 'PDL 0:CLIK_PDL=0' 'PDL 2 (steps overlap):CLIK_PDL=2' 'PDL 2, direct launches:CLIK_PDL=2,CLIK_BENCH_GRAPH=0' 'PDL 0, direct launches:CLIK_PDL=0,CLIK_BENCH_GRAPH=0'
echo "== iiwa_multitask (2^20)"
TUNE_STEPS=100 python tools/tune.py iiwa_multitask 1048576 'PDL 1 (default):' 'PDL 2:CLIK_PDL=2'
echo "== ur5_moe2016_pinv (2^20)"
TUNE_STEPS=100 python tools/tune.py ur5_moe2016_pinv 1048576 'PDL 1 (default):' 'PDL 2:CLIK_PDL=2'
echo "== ur5_qp (2^18)"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'PDL 1 (default):' 'PDL 0:CLIK_PDL=0' 'PDL 2:CLIK_PDL=2'
echo "== ur5_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'PDL 1 (default):' 'PDL 0:CLIK_PDL=0' 'PDL 2:CLIK_PDL=2'
echo "== ur5_moe2016_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'PDL 1 (default):' 'PDL 0:CLIK_PDL=0' 'PDL 2:CLIK_PDL=2'
} > gpurun_out/r2_ab7.txt 2>&1
cat gpurun_out/r2_ab7.txt | cut -c1-110
