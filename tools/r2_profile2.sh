#!/bin/bash
# Round-2 closing profile session: roofline counts for every scenario with the final kernels, DRAM traffic of a
# range of launches (write-back included), full captures of the headline and the QP fast kernel.
set -u
mkdir -p gpurun_out
FULL=0 bash tools/r2_profile.sh > /dev/null 2>&1
for s in ur5_track:1048576:8 ur5_moe2016_pinv:8388608:3; do
  IFS=: read name batch k <<< "$s"
  ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/range_$name.csv python tools/ncu_range_traffic.py $name $batch $k > gpurun_out/range_$name.json 2> gpurun_out/range_$name.err
  tail -4 gpurun_out/range_$name.csv; cat gpurun_out/range_$name.json
done
ncu --set full --clock-control none --import-source on -k regex:clik_qp_fast_kernel -s 4 -c 1 -f -o gpurun_out/r2_prof_qp_fast_final \
    python bench.py --secondary-only ur5_qp > /dev/null 2> gpurun_out/r2_prof_qp_fast_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 6 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-secondary > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out | tail -12
