#!/usr/bin/env python
"""Batch-size sweep of bench.py (device-resident value only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scenario = sys.argv[1] if len(sys.argv) > 1 else "ur5_track"
extra = dict(kv.split("=") for kv in sys.argv[2:])
for lg in (16, 18, 20, 21, 22, 23, 24):
    e = dict(os.environ)
    e.update(extra)
    steps = max(20, min(400, (1 << 28) >> lg))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--scenario", scenario,
                        "--batch", str(1 << lg), "--sets", "3" if lg >= 23 else "5",
                        "--steps", str(steps), "--warmup", "5", "--e2e-steps", "1", "--no-cpu-baseline"],
                       env=e, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print("2^%d  %.4e steps/s  %.4f ms/step  hbm %.3f" % (lg, d["value"], d["ms_per_step"],
              d["roofline_detail"]["hbm"]["frac"]), flush=True)
    except Exception as exc:
        print("2^%d FAILED %s %s" % (lg, exc, p.stderr[-800:]), flush=True)
