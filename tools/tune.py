#!/usr/bin/env python
"""Run bench.py under several kernel-build / launch variants (environment knobs of
casclik_b200/codegen/emit.py and csrc/clik_abi.cu) and print one line per variant.
Usage (on the GPU box):
    python tools/tune.py <scenario> <batch> 'name:K=V,K=V' 'name2:...' > gpurun_out/tune.txt
A variant with no knobs ('plain:') is the default build."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scenario = sys.argv[1] if len(sys.argv) > 1 else "ur5_track"
batch = sys.argv[2] if len(sys.argv) > 2 else "1048576"
variants = []
for spec in sys.argv[3:] or ["plain:"]:
    name, _, kv = spec.partition(":")
    variants.append((name, dict(p.split("=", 1) for p in kv.split(",") if p)))
for name, env in variants:
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--scenario", scenario,
                        "--batch", batch, "--steps", os.environ.get("TUNE_STEPS", "200"), "--warmup", "10",
                        "--e2e-steps", "1", "--no-cpu-baseline", "--no-secondary"],
                       env=e, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        launch = d["config"]["launch"]
        print("%-40s %.4e steps/s  %.4f ms/step  hbm %.3f  launch %s" % (
            name, d["value"], d["ms_per_step"], d["roofline_detail"]["hbm"]["frac"],
            json.dumps(launch.get(launch.get("used", "plain"), launch) if isinstance(launch, dict) else launch)),
            flush=True)
    except Exception as exc:
        print("%-40s FAILED %s\n%s" % (name, exc, p.stderr[-1500:]), flush=True)
