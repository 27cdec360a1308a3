#!/bin/bash
# Round-2 A/B session: thread-per-instance vs sub-warp mapping, register caps, mode-tail routes.
mkdir -p gpurun_out
{
echo "== iiwa_multitask (2^20)"; TUNE_STEPS=60 python tools/tune.py iiwa_multitask 1048576 'default (unit-set shortcut, cap 4 CTAs/SM):' 'sub-warp mapping, whole step:CLIK_PINV_GROUP=1' 'cap 3 CTAs/SM:CLIK_MINBLOCKS=3' 'cap 5 CTAs/SM:CLIK_MINBLOCKS=5' 'no cap:CLIK_MINBLOCKS=1' 'block 64 cap 8:CLIK_BLOCK=64,CLIK_MINBLOCKS=8'
echo "== iiwa_multitask, closed-form shortcut off (29 static modes + run-time tail)"; TUNE_STEPS=30 python tools/tune.py iiwa_multitask 1048576 'fast + group tail (default):CLIK_UNIT_SETS=0' 'one kernel, in-thread tail:CLIK_UNIT_SETS=0,CLIK_PINV_SPLIT=0' 'all modes run-time, fast + group:CLIK_UNIT_SETS=0,CLIK_NSTATIC=1' 'all modes run-time, one kernel:CLIK_UNIT_SETS=0,CLIK_NSTATIC=1,CLIK_PINV_SPLIT=0' 'all modes run-time, sub-warp whole step:CLIK_UNIT_SETS=0,CLIK_NSTATIC=1,CLIK_PINV_GROUP=1'
echo "== iiwa_multitask_stress (2^20)"; TUNE_STEPS=40 python tools/tune.py iiwa_multitask_stress 1048576 'default:' 'sub-warp mapping, whole step:CLIK_PINV_GROUP=1' 'cap 3 CTAs/SM:CLIK_MINBLOCKS=3' 'no cap:CLIK_MINBLOCKS=1'
echo "== ur5_track (2^20)"; python tools/tune.py ur5_track 1048576 'default:' 'sub-warp mapping, whole step:CLIK_PINV_GROUP=1'
echo "== ur5_moe2016_pinv (2^20)"; python tools/tune.py ur5_moe2016_pinv 1048576 'default:' 'sub-warp mapping, whole step:CLIK_PINV_GROUP=1'
} > gpurun_out/r2_ab.txt 2>&1
cat gpurun_out/r2_ab.txt
