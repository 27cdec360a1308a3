#!/bin/bash
# headline kernel: inputs staged through shared memory by bulk async copies, T tiles per CTA, hardware CTA scheduling
mkdir -p gpurun_out
{
echo "== ur5_track (2^20), 2 streams"
TUNE_STEPS=200 python tools/tune.py ur5_track 1048576 'plain kernel (default):' 'staged, 2 tiles/CTA, 2 stages:CLIK_TMA=1,CLIK_TMA_TILES=2' 'staged, 3 tiles/CTA, 3 stages:CLIK_TMA=1,CLIK_TMA_TILES=3,CLIK_STAGES=3' 'staged, 4 tiles/CTA, 2 stages:CLIK_TMA=1,CLIK_TMA_TILES=4' 'staged, 4 tiles/CTA, 4 stages:CLIK_TMA=1,CLIK_TMA_TILES=4,CLIK_STAGES=4' 'staged, 8 tiles/CTA, 2 stages:CLIK_TMA=1,CLIK_TMA_TILES=8' 'staged, 1 tile/CTA:CLIK_TMA=1,CLIK_TMA_TILES=1' 'staged, persistent balanced grid:CLIK_TMA=1'
echo "== ur5_track (2^20), 1 stream plain order"
TUNE_STEPS=200 python tools/tune.py ur5_track 1048576 'plain kernel:CLIK_BENCH_STREAMS=1' 'staged, 2 tiles/CTA:CLIK_TMA=1,CLIK_TMA_TILES=2,CLIK_BENCH_STREAMS=1' 'staged, 4 tiles/CTA, 2 stages:CLIK_TMA=1,CLIK_TMA_TILES=4,CLIK_BENCH_STREAMS=1'
echo "== ur5_moe2016_pinv (2^20), 2 streams"
TUNE_STEPS=100 python tools/tune.py ur5_moe2016_pinv 1048576 'plain kernel (default):' 'staged, 2 tiles/CTA:CLIK_TMA=1,CLIK_TMA_TILES=2' 'staged, 4 tiles/CTA, 2 stages:CLIK_TMA=1,CLIK_TMA_TILES=4'
} > gpurun_out/r2_ab12.txt 2>&1
cat gpurun_out/r2_ab12.txt | cut -c1-120
