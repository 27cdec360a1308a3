#!/usr/bin/env python
"""One process, one host batch, k GPUs: solve_batch(host arrays, devices=k) — the product-level sharded
solve (BASELINE.json configs[4]: global batch 2^23 cut over 1/2/4/8 B200, final host gather) — timed end to
end (pinned host buffers in, results in host memory) and checked bit-for-bit against the 1-GPU result.

    python tools/multi_gpu_solve.py [--log2 23] [--scenarios ur5_moe2016_pinv,ur5_moe2016_qp,ur5_track]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from casclik_b200 import scenarios  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, default=23)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--scenarios", default="ur5_moe2016_pinv,ur5_moe2016_qp,ur5_track")
args = ap.parse_args()
N = 1 << args.log2
ndev = torch.cuda.device_count()
rows = []
for name in args.scenarios.split(","):
    sc = scenarios.get(name)
    ctrl = sc.make_controller()
    ctrl.setup_solver()
    inp = sc.sample(N, seed=5)
    pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    t, q, x, y = pin(inp["t"]), pin(inp["q"]), pin(inp.get("x")), pin(inp.get("y"))
    ref = None
    k = 1
    while k <= ndev:
        first = ctrl.solve_batch(t, q, x, y, devices=k)
        out = tuple(None if a is None else pin(np.empty_like(a)) for a in first)
        ctrl.solve_batch(t, q, x, y, out=out, devices=k)            # warm (handles, streams)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            ctrl.solve_batch(t, q, x, y, out=out, devices=k)
        dt = (time.perf_counter() - t0) / args.reps
        if ref is None:
            ref = tuple(None if a is None else a.copy() for a in out)
        same = all(a is None or np.array_equal(a, b) for a, b in zip(out, ref))
        rows.append({"scenario": name, "gpus": k, "global_batch": N, "e2e_steps_per_s": N / dt,
                     "ms_per_solve": 1e3 * dt, "bit_identical_to_1gpu": bool(same)})
        print("%-18s %d GPU(s): %.3e steps/s  (%.2f ms per 2^%d solve)  identical to 1 GPU: %s"
              % (name, k, N / dt, 1e3 * dt, args.log2, same), flush=True)
        k *= 2
print(json.dumps(rows))
