#!/bin/bash
# QP later passes: persistent span grids (middle pass = remaining prediction passes, tail = iteration); CTAs = resident x waves
mkdir -p gpurun_out
{
echo "== ur5_qp (2^18)"
TUNE_STEPS=100 python tools/tune.py ur5_qp 262144 'waves 1 (default):' 'waves 2:CLIK_QP_TAIL_WAVES=2' 'waves 4:CLIK_QP_TAIL_WAVES=4' 'waves 1, stream order:CLIK_PDL=1'
echo "== ur5_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_qp 1048576 'waves 1 (default):' 'waves 2:CLIK_QP_TAIL_WAVES=2' 'waves 4:CLIK_QP_TAIL_WAVES=4' 'waves 8:CLIK_QP_TAIL_WAVES=8' 'waves 1, stream order:CLIK_PDL=1'
echo "== ur5_moe2016_qp (2^20)"
TUNE_STEPS=60 python tools/tune.py ur5_moe2016_qp 1048576 'waves 1 (default):' 'waves 2:CLIK_QP_TAIL_WAVES=2' 'waves 4:CLIK_QP_TAIL_WAVES=4' 'waves 1, stream order:CLIK_PDL=1'
echo "== ur5_moe2016_qp (2^23)"
TUNE_STEPS=20 python tools/tune.py ur5_moe2016_qp 8388608 'waves 1 (default):' 'waves 4:CLIK_QP_TAIL_WAVES=4' 'waves 1, stream order:CLIK_PDL=1'
} > gpurun_out/r2_ab8.txt 2>&1
cat gpurun_out/r2_ab8.txt | cut -c1-110
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:clik_qp -s 9 -c 9 --csv --log-file gpurun_out/r2_qp_launches2.csv python bench.py --secondary-only ur5_qp > /dev/null 2>&1
grep clik_qp gpurun_out/r2_qp_launches2.csv | awk -F'","' '{print $5, $(NF)}' | head -9
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:clik_qp -s 9 -c 6 --csv --log-file gpurun_out/r2_qp_launches3.csv python bench.py --secondary-only ur5_moe2016_qp > /dev/null 2>&1
grep clik_qp gpurun_out/r2_qp_launches3.csv | awk -F'","' '{print $5, $(NF)}' | head -6
