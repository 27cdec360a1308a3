#!/bin/bash
# independent batches alternating over several CUDA streams (bench --streams) x programmatic dependent launch (CLIK_PDL)
mkdir -p gpurun_out
{
echo "== ur5_track (2^20)"
TUNE_STEPS=200 python tools/tune.py ur5_track 1048576 '1 stream, PDL 2 (default):' '1 stream, PDL 1:CLIK_PDL=1' '2 streams, PDL 1:CLIK_BENCH_STREAMS=2,CLIK_PDL=1' '2 streams, PDL 2:CLIK_BENCH_STREAMS=2' '3 streams, PDL 1:CLIK_BENCH_STREAMS=3,CLIK_PDL=1'
echo "== iiwa_multitask (2^20)"
TUNE_STEPS=100 python tools/tune.py iiwa_multitask 1048576 '1 stream, PDL 2 (default):' '2 streams, PDL 1:CLIK_BENCH_STREAMS=2,CLIK_PDL=1' '2 streams, PDL 2:CLIK_BENCH_STREAMS=2'
echo "== ur5_qp (2^18)"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 '1 stream, PDL 2 (default):' '1 stream, PDL 1:CLIK_PDL=1' '2 streams, PDL 1:CLIK_BENCH_STREAMS=2,CLIK_PDL=1' '2 streams, PDL 2:CLIK_BENCH_STREAMS=2' '3 streams, PDL 1:CLIK_BENCH_STREAMS=3,CLIK_PDL=1' '3 streams, PDL 2:CLIK_BENCH_STREAMS=3' '4 streams, PDL 2:CLIK_BENCH_STREAMS=4'
echo "== ur5_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 '1 stream, PDL 2 (default):' '2 streams, PDL 1:CLIK_BENCH_STREAMS=2,CLIK_PDL=1' '2 streams, PDL 2:CLIK_BENCH_STREAMS=2' '3 streams, PDL 2:CLIK_BENCH_STREAMS=3'
echo "== ur5_moe2016_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 '1 stream, PDL 2 (default):' '2 streams, PDL 1:CLIK_BENCH_STREAMS=2,CLIK_PDL=1' '2 streams, PDL 2:CLIK_BENCH_STREAMS=2' '3 streams, PDL 2:CLIK_BENCH_STREAMS=3'
echo "== ur5_moe2016_qp (2^23)"
TUNE_STEPS=20 python tools/tune.py ur5_moe2016_qp 8388608 '1 stream, PDL 2 (default):' '2 streams, PDL 2:CLIK_BENCH_STREAMS=2'
} > gpurun_out/r2_ab9.txt 2>&1
cat gpurun_out/r2_ab9.txt | cut -c1-110
