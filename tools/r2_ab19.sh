#!/bin/bash
# QP end to end on pinned host arrays: zero-copy kernels (fast pass + tail over PCIe) vs the staged copy pipeline
mkdir -p gpurun_out
run() {
  name="${1%%:*}"; kv="${1#*:}"; sc="$2"; b="$3"
  env $(echo $kv | tr ',' ' ') python bench.py --scenario $sc --batch $b --steps 6 --warmup 3 --e2e-steps 30 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-50s e2e %.4e steps/s  (h2d %d B, d2h %d B per instance)' % ('$name', d['e2e']['value'], d['e2e']['h2d_bytes_per_step']//$b, d['e2e']['d2h_bytes_per_step']//$b))"
}
{
for cfg in "ur5_qp 262144" "ur5_qp 1048576" "ur5_moe2016_qp 1048576"; do
  set -- $cfg; echo "== $1 $2"
  for v in "zero-copy kernels (default):" "staged pipeline:CLIK_ZERO_COPY=0" "staged pipeline, chunk 2^16:CLIK_ZERO_COPY=0,CLIK_HOST_CHUNK=65536" "zero-copy, single kernel (no split):CLIK_QP_SPLIT=0"; do run "$v" $1 $2; done
done
} > gpurun_out/r2_ab19.txt 2>&1
cat gpurun_out/r2_ab19.txt
