#!/usr/bin/env python
"""Does running the kernel directly on pinned (mapped) host memory beat the staged host pipeline?"""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from casclik_b200 import scenarios, runtime

sc = scenarios.get("ur5_track")
ctrl = sc.make_controller(); ctrl.setup_solver()
N = 1 << 20
inp = sc.sample(N, seed=3)
pin = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).pin_memory() for k in ("t", "q", "y")}
qd = torch.empty((6, N), dtype=torch.float64).pin_memory()
md = torch.empty((N,), dtype=torch.int32).pin_memory()
lib = runtime.load_library()
sk = ctrl._skill()
p = lambda t_: ctypes.c_void_p(t_.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def zero_copy():
    runtime.check(lib.clik_pinv_step(sk.handle, N, p(pin["t"]), 1, p(pin["q"]), None, p(pin["y"]), p(qd), None, p(md), stream))
    torch.cuda.synchronize()

def staged():
    ctrl.solve_batch(pin["t"].numpy(), pin["q"].numpy(), None, pin["y"].numpy(), out=(qd.numpy(), None, md.numpy()))

for name, f in (("staged pipeline", staged), ("zero-copy kernel on pinned host memory", zero_copy)):
    for _ in range(3):
        f()
    t0 = time.perf_counter()
    for _ in range(20):
        f()
    dt = (time.perf_counter() - t0) / 20
    print("%-42s %.3f ms  %.3e steps/s" % (name, dt * 1e3, N / dt))
ref = qd.clone(); staged(); print("same bits:", bool(torch.equal(ref, qd)))
