#!/bin/bash
# two streams: can the QP tail pass interleave with the other stream's fast pass when its CTAs fit the hole a fast CTA leaves?
mkdir -p gpurun_out
{
echo "== ur5_moe2016_qp (2^23), 2 streams"
TUNE_STEPS=20 python tools/tune.py ur5_moe2016_qp 8388608 'default (tail 255 regs):' 'tail capped to 3 CTAs/SM:CLIK_QP_MINBLOCKS=3' 'tail capped to 4 CTAs/SM:CLIK_QP_MINBLOCKS=4'
echo "== ur5_moe2016_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'default (tail 255 regs):' 'tail capped to 3 CTAs/SM:CLIK_QP_MINBLOCKS=3' 'tail capped to 4 CTAs/SM:CLIK_QP_MINBLOCKS=4'
echo "== ur5_qp (2^18), 2 streams"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'default (tail 255 regs):' 'tail capped to 3 CTAs/SM:CLIK_QP_MINBLOCKS=3' 'tail capped to 4 CTAs/SM:CLIK_QP_MINBLOCKS=4'
echo "== ur5_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'default (tail 255 regs):' 'tail capped to 3 CTAs/SM:CLIK_QP_MINBLOCKS=3' 'tail capped to 4 CTAs/SM:CLIK_QP_MINBLOCKS=4'
} > gpurun_out/r2_ab10.txt 2>&1
cat gpurun_out/r2_ab10.txt | cut -c1-110
