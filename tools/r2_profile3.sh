#!/bin/bash
# full captures of the remaining step kernels with the final build (Moe-2016 SRMTP at 2^23, QP tail passes)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:clik_pinv_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_moe_pinv \
    python bench.py --secondary-only ur5_moe2016_pinv > /dev/null 2> gpurun_out/r2_prof_moe_pinv.err
ncu --set full --clock-control none --import-source on -k regex:clik_qp_fast_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_moe_qp_fast \
    python bench.py --secondary-only ur5_moe2016_qp > /dev/null 2> gpurun_out/r2_prof_moe_qp_fast.err
ncu --set full --clock-control none --import-source on -k regex:clik_qp_tail_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_qp_tail_final \
    python bench.py --secondary-only ur5_qp > /dev/null 2> gpurun_out/r2_prof_qp_tail_final.err
ls -la gpurun_out/*.ncu-rep | tail -5
