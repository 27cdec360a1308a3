#!/usr/bin/env python
"""Occupancy sweep of the QP step kernel: register cap (CLIK_QP_MINBLOCKS) x block size (CLIK_BLOCK).
Usage (on the GPU box):  python tools/tune_qp.py [scenario] > gpurun_out/tune_qp.txt"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scenario = sys.argv[1] if len(sys.argv) > 1 else "ur5_qp"
VARIANTS = [("block %s, min blocks/SM %s" % (b, m), {"CLIK_BLOCK": b, "CLIK_QP_MINBLOCKS": m})
            for b, m in (("128", "0"), ("128", "3"), ("128", "4"), ("128", "5"), ("64", "0"), ("64", "6"),
                         ("64", "8"), ("64", "10"), ("32", "12"), ("32", "16"), ("256", "2"))]
for name, env in VARIANTS:
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--scenario", scenario, "--batch",
                        os.environ.get("TUNE_BATCH", "262144"), "--steps", "60", "--warmup", "5", "--e2e-steps", "1",
                        "--no-cpu-baseline"], env=e, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print("%-32s %.4e steps/s  %.4f ms/step  launch %s" % (name, d["value"], d["ms_per_step"],
                                                               json.dumps(d["config"]["launch"])), flush=True)
    except Exception as exc:
        print("%-32s FAILED %s\n%s" % (name, exc, p.stderr[-800:]), flush=True)
