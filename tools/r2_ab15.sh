#!/bin/bash
# headline kernel, staged persistent variant on two streams: stages, block size, streams
mkdir -p gpurun_out
{
echo "== ur5_track (2^20)"
TUNE_STEPS=200 python tools/tune.py ur5_track 1048576 'staged, 2 stages, 2 streams (default):' '3 stages:CLIK_TMA=1,CLIK_STAGES=3' '4 stages:CLIK_TMA=1,CLIK_STAGES=4' '3 streams:CLIK_BENCH_STREAMS=3' '4 streams:CLIK_BENCH_STREAMS=4' 'block 256:CLIK_BLOCK=256' 'block 64:CLIK_BLOCK=64' '2 streams + overlap level 2:CLIK_PDL=2' 'plain kernel, 2 streams:CLIK_BENCH_STAGED=off'
echo "== ur5_moe2016_pinv (2^23)"
TUNE_STEPS=30 python tools/tune.py ur5_moe2016_pinv 8388608 'plain (default):' 'staged:CLIK_BENCH_STAGED=on'
} > gpurun_out/r2_ab15.txt 2>&1
cat gpurun_out/r2_ab15.txt | cut -c1-150
