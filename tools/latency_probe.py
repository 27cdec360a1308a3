#!/usr/bin/env python
"""Single-instance solve() latency through the reference-style API (the notebooks' %%timeit cells)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from casclik_b200 import scenarios

for name in ("ur5_track", "ur5_moe2016_pinv", "ur5_qp"):
    sc = scenarios.get(name)
    ctrl = sc.make_controller()
    ctrl.setup_problem_functions(); ctrl.setup_solver()
    inp = sc.sample(64, seed=1)
    def one(i):
        kw = {}
        if inp["y"] is not None:
            kw["input_var"] = inp["y"][:, i]
        return ctrl.solve(float(inp["t"][i]), inp["q"][:, i], **kw)
    for i in range(8):
        one(i)
    t0 = time.perf_counter()
    n = 400
    for k in range(n):
        one(k % 64)
    dt = (time.perf_counter() - t0) / n
    print("%-20s solve(): %.1f us per call (reference notebooks: 40-323 us)" % (name, dt * 1e6))
