#!/usr/bin/env python
"""Moe-2016 example 2 of the reference notebooks (examples/notebooks/ur5_moe2016_example2.ipynb),
for N robots at once: set-based SRMTP pseudo-inverse controller, then the reactive QP controller,
simulated closed loop on the GPU.

    python examples/ur5_moe2016_batch.py [N] [steps]

Only the imports differ from the notebook's skill definition; the simulation cell is replaced by
`rollout_batch` (clip + explicit Euler on the device).
"""
import sys
import time

import numpy as np
import torch

import casclik_b200 as cc
from casclik_b200 import cs
from casclik_b200.fk import converter

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dt, max_speed = 0.008, np.pi / 5

# ---- robot and skill (notebook cells 2-8) --------------------------------------------------------
fk_dict = converter.from_denavit_hartenberg(
    joint_angles=["s"] * 6, link_lengths=[0., -0.425, -0.392, 0., 0., 0.],
    link_offsets=[0.089, 0., 0., 0.109, 0.095, 0.082],
    link_twists=[np.pi / 2, 0., 0., np.pi / 2, -np.pi / 2, 0.])
t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", 6), cs.MX.sym("dq", 6)
p_fk = cs.Function("p_fk", [t, q], [fk_dict["T_fk"](q)[:3, 3]])
omega = 0.1
path_des = cs.vertcat(0.5 * cs.sin(omega * t) * cs.sin(omega * t) + 0.2,
                      0.5 * cs.cos(omega * t) + 0.25 * cs.sin(omega * t),
                      0.5 * cs.sin(omega * t) * cs.cos(omega * t) + 0.1)
box = {"x": (0.1, 0.6, 8), "y": (-0.5, 0.4, 7), "z": (-0.3, 0.25, 9)}
constraints = [cc.SetConstraint(label="colav_" + k, expression=p_fk(t, q)[i], set_min=lo, set_max=hi,
                                priority=prio, constraint_type="hard", gain=5e2)
               for i, (k, (lo, hi, prio)) in enumerate(box.items())]
constraints.append(cc.EqualityConstraint(label="move_point2", expression=p_fk(t, q) - path_des, priority=10,
                                         constraint_type="soft", gain=0.15))
skill = cc.SkillSpecification(label="box_move", time_var=t, robot_var=q, robot_vel_var=dq,
                              constraints=constraints)
skill.print_constraints()

# ---- N initial states around the notebook's UR5_home ---------------------------------------------------
home = np.array([-50., -160., -110., -90., -90., 0.]) * np.pi / 180.0
rng = np.random.default_rng(0)
q0 = torch.from_numpy(home[:, None] + 0.2 * rng.standard_normal((6, N))).cuda()

for name, cls in (("pinv", cc.PseudoInverseController), ("qp", cc.ReactiveQPController)):
    ctrl = cls(skill_spec=skill)
    t0 = time.time()
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    print("%s: setup %.2f s" % (name, time.time() - t0))
    state = q0.clone()
    torch.cuda.synchronize()
    t0 = time.time()
    out = ctrl.rollout_batch(0.0, state, steps, dt, max_speed=max_speed)
    torch.cuda.synchronize()
    wall = time.time() - t0
    pos = np.array([p_fk(steps * dt, state[:, i].cpu().numpy()).toarray()[:, 0] for i in range(4)])
    print("%s: %d robots x %d steps in %.3f s (%.2e controller-steps/s), failed steps: %d"
          % (name, N, steps, wall, N * steps / wall, int(out["n_failed"].sum())))
    print("   end-effector of the first robots after %.1f s:\n%s" % (steps * dt, np.round(pos, 4)))
    if name == "pinv":
        print("   final modes (first 16):", out["mode"][:16].cpu().numpy())
