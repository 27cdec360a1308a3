"""CASCLIK API mirror: constraint classes, SkillSpecification, controller construction —
names, defaults, sorting, counting, error behaviour of the reference."""
import io
import contextlib

import numpy as np
import pytest

import casclik_b200 as cc
from casclik_b200 import cs


def _cart_skill():
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    eq = cc.EqualityConstraint(label="min_dist_cnstr", expression=0.75 - p, gain=1.0,
                               constraint_type="soft", priority=1)
    lim = cc.SetConstraint(label="cart_limit_cnstr", expression=p, gain=1.0, set_min=0.0, set_max=1.0)
    spd = cc.VelocitySetConstraint(label="speed_limit_cnstr", expression=p, gain=10.0,
                                   set_min=-0.275, set_max=0.275)
    return cc.SkillSpecification(label="move_to_point_skill", time_var=t, robot_var=p,
                                 robot_vel_var=dp, constraints=[eq, lim, spd]), (eq, lim, spd)


def test_exports_match_reference_package():
    for name in ("EqualityConstraint", "SetConstraint", "VelocityEqualityConstraint",
                 "VelocitySetConstraint", "SkillSpecification", "PseudoInverseController",
                 "ReactiveQPController"):
        assert hasattr(cc, name)
    assert cc.ReactiveQPController.weight_shifter == 0.001
    assert cc.PseudoInverseController.controller_type == "PseudoInverseController"


def test_print_constraints_text_matches_notebook_output():
    spec, _ = _cart_skill()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        spec.print_constraints()
    assert buf.getvalue() == (
        "SkillSpecification: move_to_point_skill\n#0: min_dist_cnstr\n#1: cart_limit_cnstr\n"
        "#2: speed_limit_cnstr\nHas virtual var: False\nHas input var: False\nN constraints: 3\n"
        "N equality:\n\tPos:1\tVel:0\nN set:\n\tPos:1\tVel:1\n")
    assert spec.count_constraints() == {"all": 3, "equality": 1, "velocity_equality": 0, "set": 1,
                                        "velocity_set": 1, "hard": 2, "soft": 1}


def test_constraints_are_stably_sorted_by_priority_and_slack_is_counted():
    t, q = cs.MX.sym("t"), cs.MX.sym("q", 3)
    a = cc.EqualityConstraint("a", q[:2], priority=5, constraint_type="soft")
    b = cc.SetConstraint("b", q[0], set_min=-1.0, set_max=1.0, priority=1)
    c = cc.EqualityConstraint("c", q[2], priority=5, constraint_type="soft")
    d = cc.SetConstraint("d", q[1], set_min=-1.0, set_max=1.0, priority=1)
    spec = cc.SkillSpecification("s", t, q, constraints=[a, b, c, d])
    assert [x.label for x in spec.constraints] == ["b", "d", "a", "c"]
    assert spec.n_slack_var == 3 and spec.slack_var.shape == (3, 1)
    assert spec.robot_vel_var.shape == (3, 1) and spec.n_robot_var == 3
    # attributes may be mutated and the list re-assigned (notebooks do this)
    a.priority = 0
    spec.constraints = [a, b, c, d]
    assert [x.label for x in spec.constraints] == ["a", "b", "d", "c"]
    spec2 = cc.SkillSpecification("hard", t, q, constraints=[b, d])
    assert spec2.slack_var is None and spec2.n_slack_var == 0


def test_has_virtual_and_has_input_are_structural():
    t, q, x, y = cs.MX.sym("t"), cs.MX.sym("q", 2), cs.MX.sym("x"), cs.MX.sym("y", 2)
    c1 = cc.EqualityConstraint("c1", q - y)
    spec = cc.SkillSpecification("s", t, q, virtual_var=x, input_var=y, constraints=[c1])
    assert spec._has_input and not spec._has_virtual and spec.n_virtual_var == 1
    c2 = cc.VelocityEqualityConstraint("c2", q[0], target=cs.sin(x))
    spec.constraints = [c1, c2]
    assert spec._has_virtual
    c3 = cc.SetConstraint("c3", q[0], set_min=-1.0, set_max=1.0)
    spec3 = cc.SkillSpecification("s3", t, q, input_var=y, constraints=[c3])
    assert not spec3._has_input and spec3.n_input_var == 2


def test_constraint_defaults_and_size_checks():
    q = cs.MX.sym("q", 3)
    s = cc.SetConstraint("s", q)
    assert np.all(s.set_min == -1e10) and np.all(s.set_max == 1e10) and s.priority == 1
    assert s.constraint_type == "hard" and s.slack_weight == 1.0 and s.gain == 1.0
    v = cc.VelocitySetConstraint("v", q)
    assert v.set_min == -1e10 and v.set_max == 1e10
    assert cc.VelocityEqualityConstraint("ve", q).target == 0.0
    cc.EqualityConstraint("ok", q, gain=np.eye(3))
    cc.EqualityConstraint("ok2", q, gain=cs.DM.eye(3))
    cc.EqualityConstraint("ok3", q, gain=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        cc.EqualityConstraint("bad", q, gain=np.eye(2))
    with pytest.raises(TypeError):
        cc.EqualityConstraint("bad", q, gain="fast")
    with pytest.raises(ValueError):
        cc.SetConstraint("bad", q, set_min=0.0, set_max=1.0)       # scalar bound, 3 rows
    with pytest.raises(TypeError):
        cc.SetConstraint("bad", q[0], set_min="low", set_max=1.0)
    with pytest.raises(TypeError):
        cc.SkillSpecification("s", cs.MX.sym("t"), q, robot_vel_var=cs.DM.zeros(3))
    with pytest.raises(ValueError):
        cc.SkillSpecification("s", cs.MX.sym("t"), q, robot_vel_var=cs.MX.sym("dq", 2))
    merged = cc.EqualityConstraint("a", q[:2], gain=2.0) + cc.EqualityConstraint("b", q[2], gain=3.0)
    assert merged.label == "a+b" and merged.expression.shape == (3, 1)
    with pytest.raises(TypeError):
        cc.EqualityConstraint("a", q[0], priority=1) + cc.EqualityConstraint("b", q[1], priority=2)


def test_pinv_controller_attributes_and_mode_table():
    t, q = cs.MX.sym("t"), cs.MX.sym("q", 3)
    sets = [cc.SetConstraint("lim%d" % i, q[i], set_min=-1.0, set_max=1.0, priority=i) for i in range(3)]
    eq = cc.EqualityConstraint("task", q - 0.5, priority=3)
    spec = cc.SkillSpecification("s", t, q, constraints=[eq] + sets)
    ctrl = cc.PseudoInverseController(skill_spec=spec)
    assert ctrl.n_set_constraints == 3 and ctrl.n_modes == 8 and ctrl.n_state_var == 3
    assert ctrl.activation_map == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1],
                                   [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]]
    assert ctrl.options == {"feedforward": True, "multidim_sets": False,
                            "converge_final_set_to_max": False, "pinv_method": "damped",
                            "damping_factor": 1e-7,
                            "function_opts": {"jit": True, "print_time": False,
                                              "jit_options": {"flags": "-O2"}}}
    modes = ctrl.get_problem_expressions()
    assert len(modes) == 8 and modes[4]["active_set_names"] == ["lim0", "lim1"]
    assert ctrl.solve_initial_problem(0.0, [0, 0, 0]) == (None, None)
    # options dict is kept by reference and read at setup time (Appendix A19)
    opts = {"damping_factor": 1e-3}
    c2 = cc.PseudoInverseController(spec, options=opts)
    assert c2.options is opts and opts["pinv_method"] == "damped"
    with pytest.raises(RuntimeError):
        c2.solve(0.0, [0.0, 0.0, 0.0])          # not set up
    multi = cc.SkillSpecification("m", t, q, constraints=[cc.SetConstraint("box", q), eq])
    with pytest.raises(NotImplementedError):
        cc.PseudoInverseController(multi).setup_problem_functions(load=False)
    cc.PseudoInverseController(multi, options={"multidim_sets": True}).setup_problem_functions(load=False)
    # converge_final_set_to_max: only takes effect when the final constraint is a set; it needs a task above
    cc.PseudoInverseController(multi, options={"multidim_sets": True, "converge_final_set_to_max": True}
                               ).setup_problem_functions(load=False)
    tail_set = cc.SkillSpecification("ts", t, q, constraints=[
        cc.SetConstraint("lim", q[0], set_min=-1.0, set_max=1.0, priority=9)])
    with pytest.raises(ValueError):
        cc.PseudoInverseController(tail_set, options={"converge_final_set_to_max": True}).setup_problem_functions(load=False)


def test_qp_controller_weights_and_options():
    spec, _ = _cart_skill()
    ctrl = cc.ReactiveQPController(skill_spec=spec, robot_var_weights=[1.0])
    assert ctrl.options["solver_name"] == "qpoases" and ctrl.options["solver_opts"]["printLevel"] == "none"
    assert ctrl.robot_var_weights.shape == (1, 1) and ctrl.slack_var_weights.shape == (1, 1)
    with pytest.raises(ValueError):
        cc.ReactiveQPController(spec, robot_var_weights=[1.0, 2.0])
    with pytest.raises(ValueError):
        cc.ReactiveQPController(spec, slack_var_weights=np.ones(3))
    H = ctrl.get_cost_expr()
    A, lb, ub = ctrl.get_constraints_expr()
    assert np.allclose(H.toarray(), np.diag([0.001, 1.001]))
    assert A.shape == (3, 2) and lb.shape == (3, 1) and ub.shape == (3, 1)
    f = cs.Function("f", [spec.time_var, spec.robot_var], [A, lb, ub])
    An, lbn, ubn = (o.toarray() for o in f(0.0, 0.0))
    assert np.array_equal(An, [[-1, -1], [1, 0], [1, 0]])
    assert np.allclose(lbn[:, 0], [-0.75, 0.0, -0.275]) and np.allclose(ubn[:, 0], [-0.75, 1.0, 0.275])
    ctrl.setup_problem_functions(load=False)
    assert np.allclose(ctrl.H_func(0.0, 0.3).toarray(), np.diag([0.001, 1.001]))
    assert np.allclose(ctrl.Blb_func(0.0, 0.3).toarray()[:, 0], [-0.45, -0.3, -0.275])
