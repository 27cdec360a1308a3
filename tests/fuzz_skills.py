"""Random skills for differential testing of the whole stack (expression layer -> lowering -> emitted
CUDA -> kernel headers) against the oracle: random smooth expressions of q / t / virtual / input
variables in 1-4 constraints of every class, random priorities, gains and options."""
import numpy as np

import casclik_b200 as cc
from casclik_b200 import cs


def rand_expr(rng, syms, depth=0):
    """random smooth scalar expression of the given symbols"""
    pick = rng.integers(0, 6)
    a = syms[rng.integers(len(syms))]
    b = syms[rng.integers(len(syms))]
    c = float(rng.uniform(-1, 1))
    if pick == 0: return a + c * b
    if pick == 1: return cs.sin(a) + c * b
    if pick == 2: return a * b + c
    if pick == 3: return cs.cos(a + b) * c + a
    if pick == 4: return a - c * cs.sin(b) * a
    return c * a + 0.5 * b * b

def make_skill(seed):
    rng = np.random.default_rng(seed)
    nq = int(rng.integers(2, 6))
    t, q = cs.MX.sym("t"), cs.MX.sym("q", nq)
    has_x, has_y = rng.random() < 0.35, rng.random() < 0.4
    x = cs.MX.sym("x") if has_x else None
    y = cs.MX.sym("y", 2) if has_y else None
    syms = [q[i] for i in range(nq)] + ([x] if has_x else []) + ([y[0], y[1]] if has_y else [])
    syms_t = syms + [t]
    cons = []
    n_cons = int(rng.integers(1, 5))
    n_sets = 0
    for k in range(n_cons):
        kind = rng.choice(["eq", "eq", "set", "set", "veleq"])
        prio = int(rng.integers(0, 6))
        if kind == "eq":
            rows = int(rng.integers(1, 4))
            e = cs.vertcat(*[rand_expr(rng, syms_t) for _ in range(rows)])
            gain = float(rng.uniform(0.5, 3.0)) if rng.random() < 0.7 else np.diag(rng.uniform(0.5, 2.0, rows))
            cons.append(cc.EqualityConstraint("eq%d" % k, e, gain=gain, priority=prio))
        elif kind == "set" and n_sets < 4:
            n_sets += 1
            e = syms[rng.integers(nq)] if rng.random() < 0.5 else rand_expr(rng, syms_t)
            lo = float(rng.uniform(-0.6, 0.0)); hi = lo + float(rng.uniform(0.2, 0.8))
            cons.append(cc.SetConstraint("set%d" % k, e, set_min=lo, set_max=hi, gain=float(rng.uniform(0.5, 3)), priority=prio))
        else:
            e = rand_expr(rng, syms)
            cons.append(cc.VelocityEqualityConstraint("vel%d" % k, e, target=float(rng.uniform(-0.3, 0.3)), priority=prio))
    if not any(isinstance(c, (cc.EqualityConstraint, cc.VelocityEqualityConstraint)) for c in cons):
        cons.append(cc.EqualityConstraint("eq_last", rand_expr(rng, syms_t), priority=9))
    kw = dict(virtual_var=x, input_var=y)
    spec = cc.SkillSpecification("fuzz%d" % seed, t, q, constraints=cons, **{k: v for k, v in kw.items() if v is not None})
    opts = {}
    if rng.random() < 0.5: opts["damping_factor"] = 1e-4
    if rng.random() < 0.3: opts["feedforward"] = False
    N = 400
    inp = {"t": rng.uniform(0, 3, N), "q": rng.uniform(-0.9, 0.9, (nq, N))}
    if has_x: inp["x"] = rng.uniform(-0.9, 0.9, (1, N))
    if has_y: inp["y"] = rng.uniform(-0.9, 0.9, (2, N))
    return spec, opts, inp


def make_qp_skill(seed):
    """Random skill for the QP controller: every constraint class, hard and soft, random weights."""
    rng = np.random.default_rng(10_000 + seed)
    nq = int(rng.integers(2, 6))
    t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", nq), cs.MX.sym("dq", nq)
    has_x, has_y = rng.random() < 0.35, rng.random() < 0.4
    x, dx = (cs.MX.sym("x"), cs.MX.sym("dx")) if has_x else (None, None)
    y = cs.MX.sym("y", 2) if has_y else None
    syms = [q[i] for i in range(nq)] + ([x] if has_x else []) + ([y[0], y[1]] if has_y else [])
    syms_t = syms + [t]
    cons = []
    for k in range(int(rng.integers(1, 5))):
        kind = rng.choice(["eq", "set", "veleq", "velset", "limits"])
        soft = "soft" if rng.random() < 0.6 else "hard"
        sw = float(rng.uniform(0.5, 3.0))
        if kind == "eq":
            rows = int(rng.integers(1, 3))
            e = cs.vertcat(*[rand_expr(rng, syms_t) for _ in range(rows)])
            cons.append(cc.EqualityConstraint("eq%d" % k, e, gain=float(rng.uniform(0.5, 3.0)), constraint_type="soft",
                                              slack_weight=sw))
        elif kind == "set":
            e = rand_expr(rng, syms_t)
            lo = float(rng.uniform(-0.6, 0.0))
            cons.append(cc.SetConstraint("set%d" % k, e, set_min=lo, set_max=lo + float(rng.uniform(0.3, 0.9)),
                                         gain=float(rng.uniform(0.5, 2.0)), constraint_type=soft, slack_weight=sw))
        elif kind == "veleq":
            cons.append(cc.VelocityEqualityConstraint("vel%d" % k, rand_expr(rng, syms), target=float(rng.uniform(-0.3, 0.3)),
                                                      constraint_type="soft", slack_weight=sw))
        elif kind == "velset":
            cons.append(cc.VelocitySetConstraint("spd%d" % k, q, set_min=-float(rng.uniform(0.3, 1.0)) * np.ones(nq),
                                                 set_max=float(rng.uniform(0.3, 1.0)) * np.ones(nq)))
        else:
            cons.append(cc.SetConstraint("lim%d" % k, q, set_min=-np.ones(nq), set_max=np.ones(nq),
                                         gain=float(rng.uniform(0.5, 2.0))))
    kw = {"robot_vel_var": dq}
    if has_x:
        kw.update(virtual_var=x, virtual_vel_var=dx)
    if has_y:
        kw["input_var"] = y
    spec = cc.SkillSpecification("qpfuzz%d" % seed, t, q, constraints=cons, **kw)
    weights = {}
    if rng.random() < 0.5:
        weights["robot_var_weights"] = [float(v) for v in rng.uniform(0.5, 2.0, nq)]
    N = 200
    inp = {"t": rng.uniform(0, 3, N), "q": rng.uniform(-0.9, 0.9, (nq, N))}
    if has_x:
        inp["x"] = rng.uniform(-0.9, 0.9, (1, N))
    if has_y:
        inp["y"] = rng.uniform(-0.9, 0.9, (2, N))
    return spec, weights, inp


def make_dense_qp_skill(seed):
    """More than 6 dense rows (or more than 12 variables): the generic in-kernel QP solver."""
    rng = np.random.default_rng(20_000 + seed)
    nq = int(rng.integers(3, 6))
    t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", nq), cs.MX.sym("dq", nq)
    syms_t = [q[i] for i in range(nq)] + [t]
    cons = []
    for k in range(4):
        rows = int(rng.integers(2, 4))
        e = cs.vertcat(*[rand_expr(rng, syms_t) for _ in range(rows)])
        if k < 2:
            cons.append(cc.EqualityConstraint("eq%d" % k, e, gain=float(rng.uniform(0.5, 3.0)), constraint_type="soft",
                                              slack_weight=float(rng.uniform(0.5, 3.0))))
        else:
            lo = rng.uniform(-0.6, 0.0, rows)
            cons.append(cc.SetConstraint("set%d" % k, e, set_min=lo, set_max=lo + rng.uniform(0.3, 0.9, rows),
                                         gain=float(rng.uniform(0.5, 2.0)), constraint_type="soft"))
    cons.append(cc.VelocitySetConstraint("spd", q, set_min=-np.ones(nq), set_max=np.ones(nq)))
    spec = cc.SkillSpecification("qpdense%d" % seed, t, q, robot_vel_var=dq, constraints=cons)
    N = 120
    return spec, {}, {"t": rng.uniform(0, 3, N), "q": rng.uniform(-0.9, 0.9, (nq, N))}


def make_option_skill(seed):
    """Skills for the two experimental pinv options: vector-valued sets with multidim_sets, or a final
    SetConstraint with converge_final_set_to_max."""
    rng = np.random.default_rng(30_000 + seed)
    nq = int(rng.integers(3, 6))
    t, q = cs.MX.sym("t"), cs.MX.sym("q", nq)
    syms_t = [q[i] for i in range(nq)] + [t]
    multidim = seed % 2 == 0
    cons = [cc.EqualityConstraint("task", cs.vertcat(*[rand_expr(rng, syms_t) for _ in range(int(rng.integers(1, 3)))]),
                                  gain=float(rng.uniform(0.5, 2.0)), priority=5)]
    if multidim:
        for k in range(int(rng.integers(1, 3))):
            rows = int(rng.integers(2, 4))
            e = cs.vertcat(*[rand_expr(rng, syms_t) for _ in range(rows)])
            lo = rng.uniform(-0.6, 0.0, rows)
            cons.append(cc.SetConstraint("box%d" % k, e, set_min=lo, set_max=lo + rng.uniform(0.3, 0.9, rows),
                                         gain=float(rng.uniform(0.5, 3.0)), priority=int(rng.integers(0, 5))))
        opts = {"multidim_sets": True}
    else:
        cons.append(cc.SetConstraint("mid", rand_expr(rng, syms_t), set_min=-0.4, set_max=0.3, priority=2))
        lo = float(rng.uniform(-0.6, 0.0))
        cons.append(cc.SetConstraint("final", rand_expr(rng, syms_t), set_min=lo, set_max=lo + float(rng.uniform(0.3, 0.9)),
                                     gain=float(rng.uniform(0.5, 3.0)), priority=9))
        opts = {"converge_final_set_to_max": True}
    if rng.random() < 0.5:
        opts["damping_factor"] = 1e-4
    spec = cc.SkillSpecification("optfuzz%d" % seed, t, q, constraints=cons)
    N = 400
    return spec, opts, {"t": rng.uniform(0, 3, N), "q": rng.uniform(-0.9, 0.9, (nq, N))}
