#!/usr/bin/env python
"""Single-instance solve() latency: through the reference-style Python API (the notebooks' %%timeit cells,
BASELINE.md §1: 40-323 us per solve on their CPU) and of the C ABI call underneath (clik_*_solve_one)."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from casclik_b200 import scenarios, runtime  # noqa: E402

for name in ("ur5_track", "ur5_moe2016_pinv", "iiwa_multitask", "ur5_qp", "ur5_moe2016_qp"):
    sc = scenarios.get(name)
    ctrl = sc.make_controller()
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    inp = sc.sample(64, seed=1)

    def one(i):
        kw = {}
        if inp["y"] is not None:
            kw["input_var"] = inp["y"][:, i]
        return ctrl.solve(float(inp["t"][i]), inp["q"][:, i], **kw)

    for i in range(8):
        one(i)
    n = 2000
    t0 = time.perf_counter()
    for k in range(n):
        one(k % 64)
    dt_api = (time.perf_counter() - t0) / n
    # the ABI call alone, arguments prepared once
    lib, skill = runtime.load_library(), ctrl._skill()
    q = np.ascontiguousarray(inp["q"][:, 0])
    y = None if inp["y"] is None else np.ascontiguousarray(inp["y"][:, 0])
    vp = ctypes.c_void_p
    yp = vp(y.ctypes.data) if y is not None else None
    if sc.controller == "qp":
        sol, st, act = np.empty(ctrl._qn), np.zeros(1, np.int32), np.zeros(2, np.uint32)
        args = (skill.handle, ctypes.c_double(0.0), vp(q.ctypes.data), None, yp, None, vp(sol.ctypes.data),
                vp(st.ctypes.data), vp(act.ctypes.data), 0)
        fn = lib.clik_qp_solve_one
    else:
        qd, md = np.empty(q.size), np.zeros(1, np.int32)
        args = (skill.handle, ctypes.c_double(0.0), vp(q.ctypes.data), None, yp, vp(qd.ctypes.data), None,
                vp(md.ctypes.data))
        fn = lib.clik_pinv_solve_one
    for _ in range(20):
        runtime.check(fn(*args))
    t0 = time.perf_counter()
    for _ in range(n):
        fn(*args)
    dt_abi = (time.perf_counter() - t0) / n
    print("%-20s solve(): %.1f us per call through the Python API, %.1f us in the C ABI call "
          "(reference notebooks: 40-323 us)" % (name, dt_api * 1e6, dt_abi * 1e6), flush=True)
