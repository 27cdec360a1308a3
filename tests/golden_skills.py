"""Catalogue of the skills behind tests/golden/controller_vectors.json.

Every entry is built from a namespace `ns` that provides the constraint / specification classes
and `cs` (the CasADi-like symbolic module).  The generator (tests/golden/make_controller_vectors.py)
passes the REFERENCE's classes, so the unmodified reference controllers produce the vectors; the
tests pass casclik_b200's classes and must reproduce them.  The seeded inputs are part of the
fixture, so nothing here has to be re-sampled identically.
"""
import math

import numpy as np

N_CANDIDATES = 400   # instances the generator runs; it keeps a subset that covers the modes seen


class Namespace(object):
    def __init__(self, cs, mod):
        self.cs = cs
        for n in ("EqualityConstraint", "SetConstraint", "VelocityEqualityConstraint",
                  "VelocitySetConstraint", "SkillSpecification", "PseudoInverseController",
                  "ReactiveQPController"):
            setattr(self, n, getattr(mod, n))


def _fk():
    from casclik_b200 import fk
    return fk


def _ur5_q(rng, N):
    return rng.uniform(0.25 * math.pi, 0.75 * math.pi, size=(6, N))


def ur5_track(ns, rng, N):
    cs = ns.cs
    d = _fk().ur5()
    t, q, y = cs.MX.sym("t"), cs.MX.sym("q", 6), cs.MX.sym("y", 3)
    p = d["T_fk"](q)[:3, 3]
    c = ns.EqualityConstraint(label="track_point", expression=p - y, gain=1.0)
    spec = ns.SkillSpecification(label="ur5_track", time_var=t, robot_var=q, input_var=y, constraints=[c])
    inp = {"t": np.zeros(N), "q": _ur5_q(rng, N), "y": rng.uniform(-0.5, 0.5, (3, N))}
    return spec, inp


def _moe_parts(ns):
    cs = ns.cs
    d = _fk().from_denavit_hartenberg(
        joint_angles=["s"] * 6, link_lengths=[0., -0.425, -0.392, 0., 0., 0.],
        link_offsets=[0.089, 0., 0., 0.109, 0.095, 0.082],
        link_twists=[math.pi / 2, 0., 0., math.pi / 2, -math.pi / 2, 0.])
    t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", 6), cs.MX.sym("dq", 6)
    p = d["T_fk"](q)[:3, 3]
    w = 0.1
    path = cs.vertcat(0.5 * cs.sin(w * t) * cs.sin(w * t) + 0.2, 0.5 * cs.cos(w * t) + 0.25 * cs.sin(w * t),
                      0.5 * cs.sin(w * t) * cs.cos(w * t) + 0.1)
    return t, q, dq, p, path


def ur5_moe2016(ns, rng, N):
    t, q, dq, p, path = _moe_parts(ns)
    cx = ns.SetConstraint(label="colav_x", expression=p[0], set_min=0.1, set_max=0.6, priority=8, gain=5e2)
    cy = ns.SetConstraint(label="colav_y", expression=p[1], set_min=-0.5, set_max=0.4, priority=7, gain=5e2)
    cz = ns.SetConstraint(label="colav_z", expression=p[2], set_min=-0.3, set_max=0.25, priority=9, gain=5e2)
    cp = ns.EqualityConstraint(label="move_point2", expression=p - path, priority=10,
                               constraint_type="soft", gain=0.15)
    spec = ns.SkillSpecification(label="box_move", time_var=t, robot_var=q, robot_vel_var=dq,
                                 constraints=[cx, cy, cz, cp])
    return spec, {"t": rng.uniform(0.0, 80.0, N), "q": _ur5_q(rng, N)}


def ur5_moe2016_multidim(ns, rng, N):
    t, q, dq, p, path = _moe_parts(ns)
    box = ns.SetConstraint(label="colav_box", expression=p, set_min=np.array([0.1, -0.5, -0.3]),
                           set_max=np.array([0.6, 0.4, 0.25]), priority=7, gain=5e2)
    cp = ns.EqualityConstraint(label="move_point2", expression=p - path, priority=10,
                               constraint_type="soft", gain=0.15)
    spec = ns.SkillSpecification(label="box_move_multidim", time_var=t, robot_var=q, robot_vel_var=dq,
                                 constraints=[box, cp])
    return spec, {"t": rng.uniform(0.0, 80.0, N), "q": _ur5_q(rng, N)}


def iiwa_multitask(ns, rng, N):
    cs = ns.cs
    d = _fk().iiwa14()
    n = 7
    t, q, y = cs.MX.sym("t"), cs.MX.sym("q", n), cs.MX.sym("y", 12)
    lower, upper = np.array(d["lower"]), np.array(d["upper"])
    T = d["T_fk"](q)
    R, p = T[:3, :3], T[:3, 3]
    R_des, p_des = cs.reshape(y[:9], 3, 3), y[9:]
    lims = [ns.SetConstraint(label="limit_q_%d" % i, expression=q[i], set_min=float(lower[i]),
                             set_max=float(upper[i]), priority=i) for i in range(n)]
    pose = cs.vertcat(p - p_des, cs.norm_fro(cs.mtimes(R_des.T, R) - np.eye(3)))
    pose_c = ns.EqualityConstraint(label="pose", expression=pose, gain=1.0, priority=n)
    spec = ns.SkillSpecification(label="iiwa_multitask", time_var=t, robot_var=q, input_var=y,
                                 constraints=lims + [pose_c])
    width = (upper - lower)[:, None]
    qs = rng.uniform(0.0, 1.0, (n, N)) * (1.1 * width) + (lower[:, None] - 0.05 * width)
    # desired pose: a rotation about a random axis + a reachable point
    ys = np.zeros((12, N))
    for i in range(N):
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = rng.uniform(-math.pi, math.pi)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        Rm = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * K @ K
        ys[:9, i] = Rm.reshape(-1, order="F")
        ys[9:, i] = rng.uniform(-0.5, 0.5, 3) + np.array([0.0, 0.0, 0.6])
    return spec, {"t": np.zeros(N), "q": qs, "y": ys}


def cart_path(ns, rng, N):
    """Cart path-following skill of the notebooks (virtual path variable) + a VelocityEquality."""
    cs = ns.cs
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    x, dx = cs.MX.sym("x"), cs.MX.sym("dx")
    up = ns.EqualityConstraint("move_up_path_cnstr", 300 - x, gain=1.0, priority=1)
    lim = ns.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0, priority=1)
    dist = ns.EqualityConstraint("min_dist_cnstr", 0.4 * cs.sin(0.3 * x) - p, gain=1.0,
                                 constraint_type="soft", priority=3)
    vel = ns.VelocityEqualityConstraint("drift", p + 0.1 * x, target=0.05, priority=4)
    spec = ns.SkillSpecification("path", t, p, robot_vel_var=dp, virtual_var=x, virtual_vel_var=dx,
                                 constraints=[up, dist, lim, vel])
    return spec, {"t": np.zeros(N), "q": rng.uniform(-0.2, 1.2, (1, N)), "x": rng.uniform(0, 20, (1, N))}


def kitchen_sink(ns, rng, N):
    cs = ns.cs
    t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", 4), cs.MX.sym("dq", 4)
    x, dx, y = cs.MX.sym("x"), cs.MX.sym("dx"), cs.MX.sym("y", 2)
    e1 = cs.vertcat(cs.sin(q[0]) + q[1] * cs.cos(0.3 * t) - y[0], q[2] * q[3] - y[1] + 0.1 * x)
    c1 = ns.EqualityConstraint("mat_gain", e1, gain=np.array([[2.0, 0.3], [0.0, 1.5]]), priority=2)
    c2 = ns.SetConstraint("expr_bounds", q[1] + 0.2 * cs.sin(t), gain=3.0,
                          set_min=cs.MX(-0.4) + 0.0 * y[0] - 0.1 * cs.cos(x), set_max=cs.MX(0.5) + 0.05 * y[1],
                          priority=1)
    c3 = ns.VelocityEqualityConstraint("vel_target", q[0] + 0.5 * q[3], target=0.2 * cs.sin(t) + 0.1 * y[0],
                                       priority=3)
    # (a list gain passes the reference's size check but its controllers then call cs.mtimes(list, e),
    # which CasADi rejects for a 2-row expression: the diagonal is spelled out as a matrix here)
    c4 = ns.EqualityConstraint("diag_gain", cs.vertcat(q[2] - 0.3, x - t), gain=np.diag([0.7, 1.3]), priority=4)
    c5 = ns.SetConstraint("plain", q[3], set_min=-0.2, set_max=0.3, priority=0)
    c6 = ns.VelocitySetConstraint("ignored_by_pinv", q, set_min=-1.0 * np.ones(4), set_max=np.ones(4))
    c7 = ns.EqualityConstraint("expr_gain", q[0] - q[1], gain=cs.MX(1.0) + q[2] * q[2], priority=5)
    spec = ns.SkillSpecification("sink", t, q, robot_vel_var=dq, virtual_var=x, virtual_vel_var=dx,
                                 input_var=y, constraints=[c1, c2, c3, c4, c5, c6, c7])
    inp = {"t": rng.uniform(0, 10, N), "q": rng.uniform(-0.6, 0.6, (4, N)), "x": rng.uniform(-1, 1, (1, N)),
           "y": rng.uniform(-0.5, 0.5, (2, N))}
    return spec, inp


def conv_last(ns, rng, N):
    cs = ns.cs
    t, q = cs.MX.sym("t"), cs.MX.sym("q", 4)
    reach = ns.EqualityConstraint("reach", cs.vertcat(cs.sin(q[0]) + q[1] - 0.4 * cs.cos(0.2 * t), q[2] * q[3] - 0.1),
                                  gain=1.5, priority=1)
    lim = ns.SetConstraint("lim", q[1], set_min=-0.3, set_max=0.35, priority=2)
    last = ns.SetConstraint("final_set", q[0] + 0.5 * q[3], gain=2.0, set_min=-0.2, set_max=0.25, priority=3)
    spec = ns.SkillSpecification("conv", t, q, constraints=[last, reach, lim])
    return spec, {"t": rng.uniform(0, 10, N), "q": rng.uniform(-0.6, 0.6, (4, N))}


def ur5_qp(ns, rng, N):
    cs = ns.cs
    d = _fk().ur5()
    t, q, dq, y = cs.MX.sym("t"), cs.MX.sym("q", 6), cs.MX.sym("dq", 6), cs.MX.sym("y", 3)
    p = d["T_fk"](q)[:3, 3]
    max_speed = math.pi / 5
    c_pos = ns.EqualityConstraint(label="Minimize_point_error", expression=y - p, gain=50.,
                                  constraint_type="soft")
    c_lim = ns.SetConstraint(label="Joint_Limits", expression=q, set_min=np.array(d["lower"]),
                             set_max=np.array(d["upper"]))
    c_spd = ns.VelocitySetConstraint(label="Joint_speed_limits", expression=q,
                                     set_min=-cs.vertcat([max_speed] * 6), set_max=cs.vertcat([max_speed] * 6))
    spec = ns.SkillSpecification(label="ur5_qp", time_var=t, robot_var=q, robot_vel_var=dq, input_var=y,
                                 constraints=[c_pos, c_lim, c_spd])
    qs, ys = _ur5_q(rng, N), rng.uniform(-0.5, 0.5, (3, N))
    # the first instances start a few millimetres from their target: the regime the notebook's closed
    # loop runs in (K = 50, dt = 8 ms).  Far targets saturate every speed limit and the weakly coupled
    # wrist joints go bang-bang, where a 1e-12 perturbation moves a switching time by a whole step —
    # fine for single steps, useless as a closed-loop known answer.
    p_fk = cs.Function("p_fk", [q], [p])
    for i in range(4):
        ys[:, i] = np.asarray(p_fk(qs[:, i]).toarray()).reshape(-1) + rng.uniform(-0.004, 0.004, 3)
    return spec, {"t": np.zeros(N), "q": qs, "y": ys}


def cart_path_qp(ns, rng, N):
    cs = ns.cs
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    x, dx = cs.MX.sym("x"), cs.MX.sym("dx")
    up = ns.EqualityConstraint("move_up_path_cnstr", 300 - x, gain=1.0, constraint_type="soft", priority=1)
    slow = ns.VelocitySetConstraint("slow_path_cnstr", x, set_min=-0.5, set_max=0.5)
    dist = ns.EqualityConstraint("min_dist_cnstr", 0.4 * cs.sin(0.3 * x) - p, gain=1.0,
                                 constraint_type="soft", priority=1)
    lim = ns.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0)
    spd = ns.VelocitySetConstraint("speed_limit_cnstr", p, set_min=-0.275, set_max=0.275)
    spec = ns.SkillSpecification("path_trajectory_skill", t, p, robot_vel_var=dp, virtual_var=x,
                                 virtual_vel_var=dx, constraints=[up, slow, dist, lim, spd])
    return spec, {"t": np.zeros(N), "q": rng.uniform(0.0, 1.0, (1, N)), "x": rng.uniform(0, 10, (1, N))}


def kitchen_sink_qp(ns, rng, N):
    cs = ns.cs
    t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", 3), cs.MX.sym("dq", 3)
    x, dx, y = cs.MX.sym("x"), cs.MX.sym("dx"), cs.MX.sym("y", 2)
    c1 = ns.EqualityConstraint("soft_eq", cs.vertcat(cs.sin(q[0]) + q[1] - y[0], q[2] * q[0] - y[1] + 0.1 * x),
                               gain=np.array([[2.0, 0.3], [0.0, 1.5]]), constraint_type="soft", slack_weight=3.0)
    # (set_min is an MX too: the reference's size check calls set_min.is_symbolic() while looking at
    # an MX set_max, constraints.py:247, and so rejects a float set_min next to an expression set_max)
    c2 = ns.SetConstraint("soft_set", q[1] + 0.2 * cs.sin(t), gain=3.0, set_min=cs.MX(-0.4),
                          set_max=cs.MX(0.5) + 0.05 * y[1], constraint_type="soft")
    c3 = ns.VelocityEqualityConstraint("hard_veleq", q[0] + 0.5 * q[2] + x, target=0.2 * cs.sin(t))
    c4 = ns.VelocitySetConstraint("speed", q, set_min=-0.8 * np.ones(3), set_max=0.8 * np.ones(3))
    c5 = ns.VelocitySetConstraint("vspeed", x, set_min=-0.5, set_max=0.5)
    spec = ns.SkillSpecification("qp_sink", t, q, robot_vel_var=dq, virtual_var=x, virtual_vel_var=dx,
                                 input_var=y, constraints=[c1, c2, c3, c4, c5])
    inp = {"t": rng.uniform(0, 10, N), "q": rng.uniform(-0.6, 0.6, (3, N)), "x": rng.uniform(-1, 1, (1, N)),
           "y": rng.uniform(-0.5, 0.5, (2, N))}
    return spec, inp


def _cart_points(rng, N):
    p = rng.uniform(-0.3, 1.4, N)
    p[:6] = [0.25, 1.2, 0.0, 1.0, 0.6, -0.1]          # the known-answer points of SURVEY §8c first
    return {"t": np.zeros(N), "q": p.reshape(1, N)}


def cart_pinv(ns, rng, N, target=0.75):
    """cart_on_track notebook, pinv flavour (SURVEY §8c KATs P1-P3)."""
    cs = ns.cs
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    lim = ns.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0, priority=1)
    eq = ns.EqualityConstraint("min_dist_cnstr", target - p, gain=1.0, priority=2)
    spec = ns.SkillSpecification("cart", t, p, robot_vel_var=dp, constraints=[eq, lim])
    return spec, _cart_points(rng, N)


def cart_pinv_far(ns, rng, N):
    return cart_pinv(ns, rng, N, target=1.5)


def cart_qp(ns, rng, N):
    """cart_on_track notebook, QP flavour (SURVEY §8c KATs Q1-Q2)."""
    cs = ns.cs
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    eq = ns.EqualityConstraint("min_dist_cnstr", 0.75 - p, gain=1.0, constraint_type="soft", priority=1)
    lim = ns.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0)
    spd = ns.VelocitySetConstraint("speed_limit_cnstr", p, gain=10.0, set_min=-0.275, set_max=0.275)
    spec = ns.SkillSpecification("cart_qp", t, p, robot_vel_var=dp, constraints=[eq, lim, spd])
    inp = _cart_points(rng, N)
    # further than 0.275 outside [0, 1] the hard limit row and the hard speed row contradict each other
    # (the reference raises there; casclik_b200 reports status 2 — tests/test_gpu_qp.py)
    inp["q"] = np.clip(inp["q"], -0.25, 1.25)
    return spec, inp


# name -> (builder, controller, constructor keyword arguments)
CASES = {
    "pinv/ur5_track": (ur5_track, "pinv", {}),
    "pinv/ur5_track_standard_pinv": (ur5_track, "pinv", {"options": {"pinv_method": "standard"}}),
    "pinv/ur5_track_lambda_1e-26": (ur5_track, "pinv", {"options": {"damping_factor": 1e-26}}),
    "pinv/ur5_moe2016": (ur5_moe2016, "pinv", {}),
    "pinv/ur5_moe2016_no_feedforward": (ur5_moe2016, "pinv", {"options": {"feedforward": False}}),
    "pinv/ur5_moe2016_scalar_sets_multidim": (ur5_moe2016, "pinv", {"options": {"multidim_sets": True}}),
    "pinv/ur5_moe2016_multidim": (ur5_moe2016_multidim, "pinv", {"options": {"multidim_sets": True}}),
    "pinv/iiwa_multitask": (iiwa_multitask, "pinv", {}),
    "pinv/cart_path": (cart_path, "pinv", {}),
    "pinv/kitchen_sink": (kitchen_sink, "pinv", {}),
    "pinv/kitchen_sink_damping_1e-4": (kitchen_sink, "pinv", {"options": {"damping_factor": 1e-4}}),
    "pinv/conv_last": (conv_last, "pinv", {"options": {"converge_final_set_to_max": True}}),
    "pinv/cart_kat": (cart_pinv, "pinv", {}),
    "pinv/cart_kat_far_target": (cart_pinv_far, "pinv", {}),
    "qp/cart_kat": (cart_qp, "qp", {"robot_var_weights": [1.0]}),
    "qp/ur5_qp": (ur5_qp, "qp", {}),
    "qp/ur5_moe2016": (ur5_moe2016, "qp", {}),
    "qp/cart_path": (cart_path_qp, "qp", {"robot_var_weights": [1.0]}),
    "qp/kitchen_sink": (kitchen_sink_qp, "qp", {"robot_var_weights": [1.0, 2.0, 0.5],
                                                "virtual_var_weights": [4.0]}),
}


def seed_of(name):
    import zlib
    return zlib.crc32(name.encode()) % (1 << 31)       # stable when cases are added


def build(ns, name, inputs=None, N=N_CANDIDATES):
    """-> (spec, inputs, controller kind, constructor kwargs).  `inputs` overrides the sampled ones."""
    builder, kind, kwargs = CASES[name]
    spec, inp = builder(ns, np.random.default_rng(seed_of(name)), N)
    return spec, (inputs if inputs is not None else inp), kind, {k: (dict(v) if isinstance(v, dict) else v)
                                                                 for k, v in kwargs.items()}
