"""Scripted uses of the class API (constructors, validation, bookkeeping) behind
tests/golden/api_behaviour.json.  tests/golden/make_api_behaviour.py runs every case against the
REFERENCE's classes (over the stand-in casadi module) and records either the returned summary or
the exception type; tests/test_golden_api.py replays them against casclik_b200's classes.
Each case takes the namespace `ns` of golden_skills.Namespace and returns something JSON-able."""
import numpy as np


def _syms(ns, nq=3):
    cs = ns.cs
    return cs.MX.sym("t"), cs.MX.sym("q", nq), cs.MX.sym("dq", nq)


def _eq(ns, **kw):
    t, q, dq = _syms(ns)
    c = ns.EqualityConstraint("c", q[:2] - 0.1, **kw)
    return [c.constraint_type, c.priority, float(c.slack_weight), list(c.size())]


def _set(ns, rows=1, **kw):
    t, q, dq = _syms(ns)
    c = ns.SetConstraint("s", q[:rows], **kw)
    lo, hi = np.asarray(c.set_min, dtype=float).reshape(-1), np.asarray(c.set_max, dtype=float).reshape(-1)
    return [lo.tolist(), hi.tolist(), c.priority, c.constraint_type]


def _spec_summary(spec):
    return {"order": [c.label for c in spec.constraints], "n_slack": spec.n_slack_var,
            "slack": None if spec.slack_var is None else list(spec.slack_var.size()),
            "n_robot": spec.n_robot_var, "n_virtual": spec.n_virtual_var, "n_input": spec.n_input_var,
            "has_virtual": bool(spec._has_virtual), "has_input": bool(spec._has_input),
            "count": spec.count_constraints(),
            "robot_vel": list(spec.robot_vel_var.size()),
            "virtual_vel": None if spec.virtual_vel_var is None else list(spec.virtual_vel_var.size())}


def _spec(ns, **over):
    cs = ns.cs
    t, q, dq = _syms(ns)
    x, dx, y = cs.MX.sym("x"), cs.MX.sym("dx"), cs.MX.sym("y", 2)
    c1 = ns.EqualityConstraint("late", q[0] - y[0], priority=5, constraint_type="soft")
    c2 = ns.SetConstraint("first", q[1], set_min=-1.0, set_max=1.0, priority=0)
    c3 = ns.VelocityEqualityConstraint("mid_a", q[2] + x, target=0.1, priority=2)
    c4 = ns.VelocitySetConstraint("mid_b", q, set_min=-np.ones(3), set_max=np.ones(3), priority=2,
                                  constraint_type="soft")
    kw = dict(label="spec", time_var=t, robot_var=q, robot_vel_var=dq, virtual_var=x, virtual_vel_var=dx,
              input_var=y, constraints=[c1, c2, c3, c4])
    for k, v in over.items():                      # plain values, or factories that need the namespace
        kw[k] = v(ns, None) if callable(v) else v
    return ns.SkillSpecification(**kw)


def _pinv(ns, options=None, sets=2, rows=1):
    t, q, dq = _syms(ns, 4)
    cons = [ns.SetConstraint("s%d" % i, q[i:i + rows], set_min=-np.ones(rows) if rows > 1 else -1.0,
                             set_max=np.ones(rows) if rows > 1 else 1.0, priority=i) for i in range(sets)]
    cons.append(ns.EqualityConstraint("task", q[0] * q[1] + q[3] - 0.2, priority=9))
    spec = ns.SkillSpecification("p", t, q, constraints=cons)
    ctrl = ns.PseudoInverseController(spec, options=options)
    try:
        ctrl.setup_problem_functions(load=False)    # casclik_b200: build, do not load onto a device
    except TypeError:
        ctrl.setup_problem_functions()              # the reference's signature
    return {"n_modes": int(ctrl.n_modes), "n_sets": int(ctrl.n_set_constraints), "n_state": int(ctrl.n_state_var),
            "map": [[int(b) for b in row] for row in ctrl.activation_map],
            "options": {k: ctrl.options[k] for k in ("feedforward", "multidim_sets", "converge_final_set_to_max",
                                                     "pinv_method", "damping_factor")}}


def _qp(ns, **kw):
    cs = ns.cs
    t, q, dq = _syms(ns)
    x, dx = cs.MX.sym("x"), cs.MX.sym("dx")
    c1 = ns.EqualityConstraint("soft", q[:2] - 0.1 * x, constraint_type="soft", slack_weight=2.5)
    c2 = ns.VelocitySetConstraint("spd", q, set_min=-np.ones(3), set_max=np.ones(3))
    c3 = ns.SetConstraint("soft_set", q[2], set_min=-0.5, set_max=0.5, constraint_type="soft")
    spec = ns.SkillSpecification("q", t, q, robot_vel_var=dq, virtual_var=x, virtual_vel_var=dx,
                                 constraints=[c1, c2, c3])
    ctrl = ns.ReactiveQPController(spec, **kw)
    H = ctrl.get_cost_expr()
    Hn = np.asarray(cs.Function("H", [t, q, x], [H])(0.0, np.zeros(3), 0.0).toarray(), dtype=float)
    return {"H_diag": np.diag(Hn).tolist(), "H_offdiag_nnz": int(np.count_nonzero(Hn - np.diag(np.diag(Hn)))),
            "mu": float(ctrl.weight_shifter), "solver": ctrl.options["solver_name"]}


CASES = {
    # --- constraint classes --------------------------------------------------------------------------
    "eq/defaults": lambda ns: _eq(ns),
    "eq/soft_priority": lambda ns: _eq(ns, constraint_type="soft", priority=7, slack_weight=3.0),
    "eq/gain_matrix_ok": lambda ns: _eq(ns, gain=np.eye(2)),
    "eq/gain_matrix_wrong_shape": lambda ns: _eq(ns, gain=np.eye(3)),
    "eq/gain_list_ok": lambda ns: _eq(ns, gain=[1.0, 2.0]),
    "eq/gain_list_wrong_length": lambda ns: _eq(ns, gain=[1.0, 2.0, 3.0]),
    "eq/gain_list_of_strings": lambda ns: _eq(ns, gain=["a", "b"]),
    "eq/gain_string": lambda ns: _eq(ns, gain="high"),
    "eq/gain_dm_scalar": lambda ns: _eq(ns, gain=ns.cs.DM(2.0)),
    "eq/gain_mx_matrix": lambda ns: _eq(ns, gain=ns.cs.MX(np.eye(2))),
    "eq/gain_mx_wrong_shape": lambda ns: _eq(ns, gain=ns.cs.MX(np.eye(3))),
    "eq/expression_two_columns": lambda ns: ns.EqualityConstraint("c", ns.cs.MX.sym("m", 2, 2)).priority,
    "set/defaults_are_1e10": lambda ns: _set(ns, rows=2),
    "set/float_bounds_scalar": lambda ns: _set(ns, set_min=-1.0, set_max=2.0),
    "set/float_bounds_on_vector": lambda ns: _set(ns, rows=2, set_min=-1.0, set_max=2.0),
    "set/array_bounds": lambda ns: _set(ns, rows=2, set_min=np.array([-1.0, -2.0]), set_max=np.array([1.0, 2.0])),
    "set/array_bounds_wrong_length": lambda ns: _set(ns, rows=2, set_min=np.array([-1.0, -2.0, -3.0]),
                                                     set_max=np.array([1.0, 2.0])),
    "set/column_array_bounds": lambda ns: _set(ns, rows=2, set_min=np.array([[-1.0], [-2.0]]),
                                               set_max=np.array([[1.0], [2.0]])),
    "set/dm_bounds": lambda ns: _set(ns, rows=2, set_min=ns.cs.DM([-1.0, -2.0]), set_max=ns.cs.DM([1.0, 2.0])),
    "set/string_bound": lambda ns: _set(ns, set_min="low", set_max=1.0),
    "set/pure_symbol_bound": lambda ns: _set(ns, set_min=ns.cs.MX.sym("b"), set_max=ns.cs.MX.sym("c")),
    "veleq/defaults": lambda ns: (lambda c: [c.target, c.priority, c.constraint_type])(
        ns.VelocityEqualityConstraint("v", _syms(ns)[1][0])),
    "velset/defaults": lambda ns: (lambda c: [float(c.set_min), float(c.set_max), c.priority])(
        ns.VelocitySetConstraint("v", _syms(ns)[1][0])),
    # --- SkillSpecification -----------------------------------------------------------------------------
    "spec/bookkeeping": lambda ns: _spec_summary(_spec(ns)),
    "spec/time_var_not_mx": lambda ns: _spec_summary(_spec(ns, time_var=0.0)),
    "spec/time_var_list": lambda ns: _spec_summary(_spec(ns, time_var=[0.0])),
    "spec/robot_vel_wrong_size": lambda ns: _spec_summary(_spec(ns, robot_vel_var=lambda ns, L: ns.cs.MX.sym("w", 2))),
    "spec/robot_vel_not_mx": lambda ns: _spec_summary(_spec(ns, robot_vel_var=np.zeros(3))),
    "spec/robot_vel_default": lambda ns: _spec_summary(_spec(ns, robot_vel_var=None)),
    "spec/virtual_vel_wrong_size": lambda ns: _spec_summary(_spec(ns, virtual_vel_var=lambda ns, L: ns.cs.MX.sym("w", 2))),
    "spec/virtual_vel_default": lambda ns: _spec_summary(_spec(ns, virtual_vel_var=None)),
    "spec/no_virtual_no_input": lambda ns: _spec_summary(_spec(ns, virtual_var=None, virtual_vel_var=None,
                                                                input_var=None)),
    "spec/unused_input": lambda ns: _spec_summary(_spec(ns, input_var=lambda ns, L: ns.cs.MX.sym("unused", 4))),
    "spec/no_constraints": lambda ns: _spec_summary(_spec(ns, constraints=[])),
    # --- controllers ------------------------------------------------------------------------------------
    "pinv/defaults_two_sets": lambda ns: _pinv(ns),
    "pinv/three_sets_map": lambda ns: _pinv(ns, sets=3),
    "pinv/options_merge": lambda ns: _pinv(ns, options={"damping_factor": 1e-3, "feedforward": False}),
    "pinv/multirow_set_needs_option": lambda ns: _pinv(ns, sets=1, rows=2),
    "pinv/multirow_set_with_option": lambda ns: _pinv(ns, options={"multidim_sets": True}, sets=1, rows=2),
    "qp/default_weights": lambda ns: _qp(ns),
    "qp/list_weights": lambda ns: _qp(ns, robot_var_weights=[1.0, 2.0, 3.0], virtual_var_weights=[4.0],
                                      slack_var_weights=[1.0, 2.0, 3.0]),
    "qp/array_weights": lambda ns: _qp(ns, robot_var_weights=np.array([0.5, 0.5, 2.0])),
    "qp/dm_weights": lambda ns: _qp(ns, robot_var_weights=ns.cs.DM([1.0, 2.0, 3.0])),
    "qp/robot_weights_wrong_length": lambda ns: _qp(ns, robot_var_weights=[1.0, 2.0]),
    "qp/virtual_weights_wrong_length": lambda ns: _qp(ns, virtual_var_weights=[1.0, 2.0]),
    "qp/slack_weights_wrong_length": lambda ns: _qp(ns, slack_var_weights=[1.0]),
    "qp/options_solver_name": lambda ns: _qp(ns, options={"solver_name": "ooqp"}),
}


def run_case(ns, name):
    """-> {"ok": summary} or {"raises": exception class name}."""
    try:
        out = CASES[name](ns)
    except Exception as exc:                       # noqa: BLE001 — the exception type IS the recorded behaviour
        return {"raises": type(exc).__name__}
    import json
    return {"ok": json.loads(json.dumps(out, default=lambda o: float(o) if isinstance(o, (np.floating, float)) else str(o)))}
