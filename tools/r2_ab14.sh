#!/bin/bash
# QP fast kernel occupancy after the flip rule / structural zeros: register caps and block sizes
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18), 2 streams"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'default (cap 3 CTAs/SM, 168 regs):' 'fast cap 4 CTAs/SM (128 regs):CLIK_QP_FAST_MINBLOCKS=4' 'fast uncapped:CLIK_QP_FAST_MINBLOCKS=0' 'block 64:CLIK_BLOCK=64' 'block 64, fast cap 8:CLIK_BLOCK=64,CLIK_QP_FAST_MINBLOCKS=8' 'block 96, cap 5:CLIK_BLOCK=96,CLIK_QP_FAST_MINBLOCKS=5'
echo "== ur5_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'default (cap 3 CTAs/SM, 168 regs):' 'fast cap 4 CTAs/SM (128 regs):CLIK_QP_FAST_MINBLOCKS=4' 'fast uncapped:CLIK_QP_FAST_MINBLOCKS=0'
echo "== ur5_moe2016_qp (2^20), 2 streams"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'default (cap 3 CTAs/SM, 168 regs):' 'fast cap 4 CTAs/SM (128 regs):CLIK_QP_FAST_MINBLOCKS=4' 'fast uncapped:CLIK_QP_FAST_MINBLOCKS=0'
echo "== iiwa_multitask (2^20), 2 streams"
TUNE_STEPS=100 python tools/tune.py iiwa_multitask 1048576 'default (cap 4 CTAs/SM, 128 regs):' 'cap 5 (96 regs):CLIK_MINBLOCKS=5' 'cap 3 (168 regs):CLIK_MINBLOCKS=3' 'block 64, cap 8:CLIK_BLOCK=64,CLIK_MINBLOCKS=8' 'block 64, cap 10:CLIK_BLOCK=64,CLIK_MINBLOCKS=10'
} > gpurun_out/r2_ab14.txt 2>&1
cat gpurun_out/r2_ab14.txt | cut -c1-150
