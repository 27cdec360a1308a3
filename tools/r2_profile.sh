#!/bin/bash
# Round-2 profiling session (run on the GPU box through gpurun): metric passes for the roofline counts of
# every BASELINE scenario + full captures of the kernels VERDICT r1 named.  Outputs under gpurun_out/.
set -u
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum
mkdir -p gpurun_out
ncu --metrics $M --clock-control none -k regex:clik_pinv_kernel -s 6 -c 4 --csv --log-file gpurun_out/cnt_ur5_track.csv \
    python bench.py --no-secondary --no-cpu-baseline --steps 6 --warmup 3 --e2e-steps 1 > /dev/null 2> gpurun_out/cnt_ur5_track.err
for s in iiwa_multitask ur5_qp ur5_moe2016_pinv ur5_moe2016_qp; do
  ncu --metrics $M --clock-control none -k regex:clik_ -s 6 -c 4 --csv --log-file gpurun_out/cnt_$s.csv \
      python bench.py --secondary-only $s > /dev/null 2> gpurun_out/cnt_$s.err
done
if [ "${FULL:-1}" = "1" ]; then
  ncu --set full --clock-control none --import-source on -k regex:clik_pinv_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_iiwa \
      python bench.py --secondary-only iiwa_multitask > /dev/null 2> gpurun_out/r2_prof_iiwa.err
  ncu --set full --clock-control none --import-source on -k regex:clik_qp_tail_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_qp_tail \
      python bench.py --secondary-only ur5_qp > /dev/null 2> gpurun_out/r2_prof_qp_tail.err
  ncu --set full --clock-control none --import-source on -k regex:clik_qp_fast_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_qp_fast \
      python bench.py --secondary-only ur5_qp > /dev/null 2> gpurun_out/r2_prof_qp_fast.err
  ncu --set full --clock-control none --import-source on -k regex:clik_pinv_kernel -s 6 -c 1 -f -o gpurun_out/r2_prof_ur5_track \
      python bench.py --no-secondary --no-cpu-baseline --steps 6 --warmup 3 --e2e-steps 1 > /dev/null 2> gpurun_out/r2_prof_ur5_track.err
fi
ls -la gpurun_out | tail -20
