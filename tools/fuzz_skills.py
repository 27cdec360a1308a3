#!/usr/bin/env python
"""Differential fuzzing without a GPU: random skills (tests/fuzz_skills.py) through the kernels' own source
compiled for the host (tests/test_kernel_code_on_host.py) against the oracle.
Usage:  python tools/fuzz_skills.py FIRST_SEED LAST_SEED"""
import ctypes
import os
import pathlib
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import casclik_b200 as cc  # noqa: E402
import test_kernel_code_on_host as H  # noqa: E402
from fuzz_skills import make_skill  # noqa: E402
from oracle_bridge import oracle_pinv  # noqa: E402

bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    spec, opts, inp = make_skill(seed)
    ctrl = cc.PseudoInverseController(spec, options=dict(opts))
    with tempfile.TemporaryDirectory() as d:
        lib = H._host_library(ctrl, pathlib.Path(d))
        t, q, x, y = H._inputs(inp)
        nq, N = q.shape
        nx = ctrl._nx
        if nx and x is None:
            x = np.zeros((nx, N))
            inp = dict(inp, x=x)
        qdot, xdot = np.full((nq, N), np.nan), (np.full((nx, N), np.nan) if nx else None)
        mode = np.full(N, -9, dtype=np.int32)
        lib.clik_pinv_kernel(ctypes.c_longlong(N), H._p(t), ctypes.c_int(1), H._p(q), H._p(x),
                             H._p(y if ctrl._ny else None), H._p(qdot), H._p(xdot), H._p(mode))
    got = qdot if xdot is None else np.vstack([qdot, xdot])
    ref_v, ref_mode = oracle_pinv(spec, inp, dict(opts))
    mismatch = float((mode != ref_mode).mean())
    same = mode == ref_mode
    err = np.linalg.norm(got[:, same] - ref_v[:, same], axis=0) / np.maximum(np.linalg.norm(ref_v[:, same], axis=0), 1e-12)
    flag = "" if (mismatch == 0.0 and err.max() < 1e-7) else "   <<<<<"
    bad += bool(flag)
    print("seed %4d  nq %d  %-30s modes %3d  mode mismatch %.4f  max rel err %.2e%s" % (
        seed, nq, ",".join(type(c).__name__[:3] + str(c.expression.size()[0]) for c in spec.constraints),
        ctrl.n_modes, mismatch, err.max(), flag), flush=True)
print("suspicious:", bad)
sys.exit(1 if bad else 0)
