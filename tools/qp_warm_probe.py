#!/usr/bin/env python
"""Cold vs warm-started QP step and rollout throughput (UR5 9x15 structured solver)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from casclik_b200 import scenarios

sc = scenarios.get("ur5_qp")
ctrl = sc.make_controller()
ctrl.setup_problem_functions(); ctrl.setup_solver()
N = 1 << 20
inp = sc.sample(N, seed=0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
t, q, y = up(inp["t"]), up(inp["q"]), up(inp["y"])
sol, st, act = ctrl.solve_batch(t, q, None, y)
torch.cuda.synchronize()


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = (torch.empty_like(sol), torch.empty_like(st), torch.empty_like(act))
ms_c = timeit(lambda: ctrl.solve_batch(t, q, None, y, out=out))
ms_w = timeit(lambda: ctrl.solve_batch(t, q, None, y, out=out, warm_active=act))
ms_x = timeit(lambda: ctrl.solve_batch(t, q, None, y, out=out, warmstart=sol))
print("step cold  %.3f ms  %.3e steps/s" % (ms_c, N / ms_c * 1e3))
print("step warm (active set of the solution) %.3f ms  %.3e steps/s" % (ms_w, N / ms_w * 1e3))
print("step warm (x0 = solution) %.3f ms  %.3e steps/s" % (ms_x, N / ms_x * 1e3))
K = 100
n2 = 1 << 18
q2, y2 = q[:, :n2].contiguous(), y[:, :n2].contiguous()


def rollout_qp():
    qq = q2.clone()
    return ctrl.rollout_batch(0.0, qq, K, 0.008, input_var=y2, max_speed=np.pi / 5)


r = rollout_qp()
ms = timeit(rollout_qp, reps=5)
print("QP rollout %d steps x %d instances: %.2f ms  %.3e controller-steps/s, failed %d" % (
    K, n2, ms, K * n2 / ms * 1e3, int(r["n_failed"].sum())))
pc = scenarios.get("ur5_track").make_controller()
pc.setup_solver()


def rollout_pinv():
    qq = q.clone()
    return pc.rollout_batch(0.0, qq, 200, 0.008, input_var=y, max_speed=np.pi / 5)


rollout_pinv()
ms = timeit(rollout_pinv, reps=5)
print("pinv rollout 200 steps x %d instances: %.2f ms  %.3e controller-steps/s" % (N, ms, 200 * N / ms * 1e3))
