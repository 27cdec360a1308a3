#!/bin/bash
mkdir -p gpurun_out
{
echo "== ur5_track (2^20): occupancy / block-size variants after the sincos clean-up"
python tools/tune.py ur5_track 1048576 'default (128 thr, 66 regs, 7 CTAs/SM):' 'cap 64 regs (8 CTAs/SM):CLIK_MINBLOCKS=8' 'block 64, cap 16:CLIK_BLOCK=64,CLIK_MINBLOCKS=16' 'block 64:CLIK_BLOCK=64' 'block 256, cap 4:CLIK_BLOCK=256,CLIK_MINBLOCKS=4' 'block 256:CLIK_BLOCK=256' 'cap 72 regs (7 CTAs):CLIK_MINBLOCKS=7' 'cap 6 CTAs:CLIK_MINBLOCKS=6'
echo "== ur5_moe2016_pinv (2^20)"
python tools/tune.py ur5_moe2016_pinv 1048576 'default:' 'cap 6 CTAs/SM:CLIK_MINBLOCKS=6' 'cap 7:CLIK_MINBLOCKS=7' 'cap 5:CLIK_MINBLOCKS=5' 'block 64:CLIK_BLOCK=64' 'block 256:CLIK_BLOCK=256'
echo "== ur5_qp (2^18)"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'default:' 'cap 4:CLIK_QP_MINBLOCKS=4' 'block 64:CLIK_BLOCK=64' 'block 64 cap 8:CLIK_BLOCK=64,CLIK_QP_MINBLOCKS=8' 'block 256:CLIK_BLOCK=256'
echo "== ur5_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'default:' 'block 64:CLIK_BLOCK=64'
echo "== ur5_moe2016_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'default:' 'cap 3:CLIK_QP_MINBLOCKS=3' 'block 64:CLIK_BLOCK=64'
} > gpurun_out/r2_ab3.txt 2>&1
cat gpurun_out/r2_ab3.txt
