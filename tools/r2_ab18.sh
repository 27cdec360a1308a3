#!/bin/bash
# e2e (pinned host arrays, zero-copy kernel over PCIe): CTAs in flight
mkdir -p gpurun_out
run() {
  name="${1%%:*}"; kv="${1#*:}"; sc="${2:-ur5_track}"; b="${3:-1048576}"
  env $(echo $kv | tr ',' ' ') python bench.py --scenario $sc --batch $b --steps 6 --warmup 3 --e2e-steps 40 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-50s e2e %.4e steps/s' % ('$name', d['e2e']['value']))"
}
{
echo "== ur5_track 2^20"
for v in "no cap (default):" "1 CTA/SM:CLIK_ZC_CTAS_PER_SM=1" "2 CTAs/SM:CLIK_ZC_CTAS_PER_SM=2" "3 CTAs/SM:CLIK_ZC_CTAS_PER_SM=3" "4 CTAs/SM:CLIK_ZC_CTAS_PER_SM=4" "6 CTAs/SM:CLIK_ZC_CTAS_PER_SM=6" "8 CTAs/SM (persistent):CLIK_ZC_CTAS_PER_SM=8" "block 256:CLIK_BLOCK=256" "block 256, 1 CTA/SM:CLIK_BLOCK=256,CLIK_ZC_CTAS_PER_SM=1" "block 256, 2 CTAs/SM:CLIK_BLOCK=256,CLIK_ZC_CTAS_PER_SM=2" "block 512:CLIK_BLOCK=512" "block 64, 4 CTAs/SM:CLIK_BLOCK=64,CLIK_ZC_CTAS_PER_SM=4"; do run "$v"; done
echo "== ur5_qp 2^18"
for v in "no cap (default):" "2 CTAs/SM:CLIK_ZC_CTAS_PER_SM=2" "1 CTA/SM:CLIK_ZC_CTAS_PER_SM=1"; do run "$v" ur5_qp 262144; done
echo "== ur5_moe2016_pinv 2^20"
for v in "no cap (default):" "2 CTAs/SM:CLIK_ZC_CTAS_PER_SM=2" "4 CTAs/SM:CLIK_ZC_CTAS_PER_SM=4"; do run "$v" ur5_moe2016_pinv 1048576; done
} > gpurun_out/r2_ab18.txt 2>&1
cat gpurun_out/r2_ab18.txt
