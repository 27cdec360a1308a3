#!/usr/bin/env python
"""Run bench.py under several kernel-build / launch variants (environment knobs of
casclik_b200/codegen/emit.py and csrc/clik_abi.cu) and print one line per variant.
Usage (on the GPU box):  python tools/tune.py [scenario] > gpurun_out/tune.txt"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scenario = sys.argv[1] if len(sys.argv) > 1 else "ur5_track"
VARIANTS = [
    ("plain unroll=1", {"CLIK_TMA": "0"}),
    ("plain unroll=2 (168 regs, interleaved)", {"CLIK_TMA": "0", "CLIK_UNROLL": "2"}),
    ("plain unroll=2 minblocks=4 (<=128)", {"CLIK_TMA": "0", "CLIK_UNROLL": "2", "CLIK_MINBLOCKS": "4"}),
    ("plain unroll=2 minblocks=5 (<=96)", {"CLIK_TMA": "0", "CLIK_UNROLL": "2", "CLIK_MINBLOCKS": "5"}),
    ("plain unroll=2 minblocks=6 (<=80)", {"CLIK_TMA": "0", "CLIK_UNROLL": "2", "CLIK_MINBLOCKS": "6"}),
    ("plain unroll=2 block=64 minblocks=10", {"CLIK_TMA": "0", "CLIK_UNROLL": "2", "CLIK_BLOCK": "64", "CLIK_MINBLOCKS": "10"}),
    ("plain unroll=4 minblocks=5", {"CLIK_TMA": "0", "CLIK_UNROLL": "4", "CLIK_MINBLOCKS": "5"}),
    ("plain unroll=1 minblocks=7 (<=72)", {"CLIK_TMA": "0", "CLIK_MINBLOCKS": "7"}),
]
for name, env in VARIANTS:
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--scenario", scenario,
                        "--batch", os.environ.get("TUNE_BATCH", "1048576"), "--steps", "400", "--warmup", "10", "--e2e-steps", "1", "--no-cpu-baseline"],
                       env=e, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print("%-45s %.4e steps/s  %.4f ms/step  hbm %.3f fp64 %.3f  launch %s" % (
            name, d["value"], d["ms_per_step"], d["roofline_detail"]["hbm"]["frac"],
            d["roofline_detail"].get("fp64", {}).get("frac", float("nan")), json.dumps(d["config"]["launch"].get(d["config"]["launch"].get("used","plain"), d["config"]["launch"]))),
            flush=True)
    except Exception as exc:
        print("%-45s FAILED %s\n%s" % (name, exc, p.stderr[-1500:]), flush=True)
