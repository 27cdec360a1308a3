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
    ("plain", {}),
    ("prefetch 512 CTAs ahead", {"CLIK_PREFETCH_CTAS": "512"}),
    ("prefetch 1036 CTAs ahead", {"CLIK_PREFETCH_CTAS": "1036"}),
    ("prefetch 2072 CTAs ahead", {"CLIK_PREFETCH_CTAS": "2072"}),
    ("prefetch 4144 CTAs ahead", {"CLIK_PREFETCH_CTAS": "4144"}),
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
