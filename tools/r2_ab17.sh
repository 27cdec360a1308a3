#!/bin/bash
# e2e (pinned host arrays, zero-copy): plain kernel loads vs the staged kernel's bulk async copies from mapped host memory
mkdir -p gpurun_out
{
for v in "plain kernel over PCIe (default):" "staged kernel over PCIe, 2 stages:CLIK_ZC_STAGED=1,CLIK_BENCH_STAGED=on" "staged, 4 stages:CLIK_ZC_STAGED=1,CLIK_BENCH_STAGED=on,CLIK_STAGES=4" "plain, unroll 2:CLIK_UNROLL=2" "plain, block 256:CLIK_BLOCK=256"; do
  name="${v%%:*}"; kv="${v#*:}"
  env $(echo $kv | tr ',' ' ') python bench.py --steps 20 --warmup 3 --e2e-steps 40 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-45s e2e %.4e steps/s   device-resident %.4e' % ('$name', d['e2e']['value'], d['value']))"
done
} > gpurun_out/r2_ab17.txt 2>&1
cat gpurun_out/r2_ab17.txt
