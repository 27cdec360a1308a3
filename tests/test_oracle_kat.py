"""Pins the oracle: reference-produced mode tables, hand-derived known answers of the in-tree
formulas, FK known answers from the notebooks."""
import json
import os

import numpy as np
import pytest

from oracle_bridge import orc, oracle_pinv, oracle_qp_problem
import casclik_b200 as cc
from casclik_b200 import cs

HERE = os.path.dirname(os.path.abspath(__file__))
LAM = 1e-7


def test_activation_maps_match_reference_output():
    golden = json.load(open(os.path.join(HERE, "golden", "activation_maps.json")))
    for k, table in golden.items():
        assert orc.activation_map(int(k)) == table
    assert golden["3"] == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1],
                           [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]]
    assert len(golden["7"]) == 128


def _cart_pinv(p_des):
    t, p = cs.MX.sym("t"), cs.MX.sym("p")
    lim = cc.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0, priority=1)
    eq = cc.EqualityConstraint("min_dist_cnstr", p_des - p, gain=1.0, priority=2)
    return cc.SkillSpecification("cart", t, p, constraints=[eq, lim])


def _run_cart(spec, p0):
    v, mode = oracle_pinv(spec, {"t": np.zeros(1), "q": np.array([[p0]])})
    return float(v[0, 0]), int(mode[0])


def test_kat_p1_mode0_double_application():
    v, mode = _run_cart(_cart_pinv(0.75), 0.25)
    assert mode == 0
    expect = 0.5 * (1 + 2 * LAM) / (1 + LAM) ** 2
    assert abs(v - expect) < 1e-15
    assert abs(v - 0.499999999999995) < 1e-15
    assert abs(v - 0.5 / (1 + LAM)) > 1e-9        # the textbook damped answer is NOT what the reference computes


def test_kat_p2_outside_but_returning():
    v, mode = _run_cart(_cart_pinv(0.75), 1.2)
    assert mode == 0
    assert abs(v - (-0.45) * (1 + 2 * LAM) / (1 + LAM) ** 2) < 1e-15


def test_kat_p3_set_activates_and_leaks():
    v, mode = _run_cart(_cart_pinv(1.5), 1.2)
    assert mode == 1
    expect = LAM / (1 + LAM) * 0.3 / (1 + LAM)
    assert abs(v - expect) < 1e-12 and abs(v - 2.99999940e-08) < 1e-15


def test_in_tangent_cone_thresholds():
    f = orc.in_tangent_cone
    one = np.ones(1)
    # inside / on the boundary (within 1e-12): always admissible, whatever the direction
    assert f(0.5 * one, -one, 0 * one, one)[0]
    assert f(0.0 * one, -one, 0 * one, one)[0]
    assert f((1 + 5e-13) * one, one, 0 * one, one)[0]
    # outside: strict sign test on the derivative
    assert not f(1.1 * one, 0.0 * one, 0 * one, one)[0]
    assert f(1.1 * one, -1e-9 * one, 0 * one, one)[0]
    assert not f(-0.1 * one, 0.0 * one, 0 * one, one)[0]
    assert f(-0.1 * one, 1e-9 * one, 0 * one, one)[0]


def _cart_qp():
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    eq = cc.EqualityConstraint("min_dist_cnstr", 0.75 - p, gain=1.0, constraint_type="soft", priority=1)
    lim = cc.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0)
    spd = cc.VelocitySetConstraint("speed_limit_cnstr", p, gain=10.0, set_min=-0.275, set_max=0.275)
    return cc.SkillSpecification("cart_qp", t, p, robot_vel_var=dp, constraints=[eq, lim, spd])


def test_kat_q_matrices():
    h, A, lb, ub = oracle_qp_problem(_cart_qp(), {"t": np.zeros(1), "q": np.array([[0.0]])})
    assert np.allclose(h, [0.001, 1.001], rtol=0, atol=1e-18)
    assert np.array_equal(A[0], [[-1, -1], [1, 0], [1, 0]])
    assert np.allclose(lb[0], [-0.75, 0.0, -0.275]) and np.allclose(ub[0], [-0.75, 1.0, 0.275])


def test_kat_q1_speed_limit_active():
    h, A, lb, ub = oracle_qp_problem(_cart_qp(), {"t": np.zeros(1), "q": np.array([[0.0]])})
    x, lam, st = orc.solve_qp(h, A, lb, ub)
    assert st[0] == 0
    assert abs(x[0, 0] - 0.275) < 1e-12 and abs(x[0, 1] - 0.475) < 1e-12
    k = orc.kkt_residuals(h, A[0], lb[0], ub[0], x[0], lam[0])
    assert abs(k["objective"] - 0.112963125) < 1e-12
    assert max(k["primal"], k["stationarity"], k["sign"]) < 1e-12
    assert lam[0, 2] > 0 and lam[0, 1] == 0        # speed row at its upper bound


def test_kat_q2_no_inequality_active():
    h, A, lb, ub = oracle_qp_problem(_cart_qp(), {"t": np.zeros(1), "q": np.array([[0.6]])})
    x, lam, st = orc.solve_qp(h, A, lb, ub)
    assert st[0] == 0
    assert abs(x[0, 0] - 0.14985029940119762) < 1e-12
    assert abs(x[0, 1] - 1.4970059880239917e-4) < 1e-12
    assert lam[0, 1] == 0 and lam[0, 2] == 0


def test_qp_oracle_random_kkt_and_scipy():
    rng = np.random.default_rng(3)
    from scipy.optimize import minimize
    for trial in range(40):
        n, m = int(rng.integers(2, 7)), int(rng.integers(1, 10))
        h = rng.uniform(0.001, 2.0, n)
        A = rng.normal(size=(m, n))
        x_feas = rng.normal(size=n)
        r = A @ x_feas
        lb = r - rng.uniform(0, 1, m)
        ub = r + rng.uniform(0, 1, m)
        eqrows = rng.random(m) < 0.2
        lb[eqrows] = r[eqrows]
        ub[eqrows] = r[eqrows]
        x, lam, st = orc.solve_qp_single(h, A, lb, ub)
        assert st == 0
        k = orc.kkt_residuals(h, A, lb, ub, x, lam)
        assert k["primal"] < 1e-9 and k["stationarity"] < 1e-9 and k["sign"] < 1e-9
        if trial < 8:
            cons = [{"type": "ineq", "fun": lambda z, A=A, lb=lb: A @ z - lb},
                    {"type": "ineq", "fun": lambda z, A=A, ub=ub: ub - A @ z}]
            ref = minimize(lambda z: 0.5 * z @ (h * z), x_feas, jac=lambda z: h * z,
                           constraints=cons, method="SLSQP", options={"ftol": 1e-14, "maxiter": 500})
            assert abs(ref.fun - k["objective"]) < 1e-6 * (1 + abs(ref.fun))


def test_qp_oracle_detects_infeasible():
    h = np.ones(2)
    A = np.array([[1.0, 0.0], [1.0, 0.0]])
    x, lam, st = orc.solve_qp_single(h, A, np.array([1.0, -3.0]), np.array([2.0, -2.0]))
    assert st == 2


def test_damped_pinv_branches():
    rng = np.random.default_rng(0)
    J = rng.normal(size=(2, 3, 6))
    P = orc.damped_pinv(J)
    assert P.shape == (2, 6, 3)
    ref = np.stack([j.T @ np.linalg.inv(j @ j.T + LAM * np.eye(3)) for j in J])
    assert np.allclose(P, ref, rtol=1e-10)
    Jt = rng.normal(size=(2, 8, 6))
    Pt = orc.damped_pinv(Jt)
    ref = np.stack([np.linalg.inv(j.T @ j + LAM * np.eye(6)) @ j.T for j in Jt])
    assert np.allclose(Pt, ref, rtol=1e-8)
    assert np.allclose(orc.damped_pinv(J, "standard"), np.linalg.pinv(J), rtol=1e-9)


def test_first_equality_is_applied_twice():
    """Appendix A1: v = P des + (I - P J) P des, not the textbook P des."""
    rng = np.random.default_rng(1)
    J = rng.normal(size=(5, 3, 6))
    e = rng.normal(size=(5, 3))
    blk = orc.Block(orc.EQ, e, J, np.zeros((5, 3)), 1.0)
    v, mode = orc.pinv_step([blk], 6)
    P = orc.damped_pinv(J)
    w = np.einsum("nij,nj->ni", P, -e)
    expect = w + np.einsum("nij,nj->ni", np.eye(6) - P @ J, w)
    assert np.allclose(v, expect, rtol=1e-13, atol=1e-15)
    assert np.all(mode == 0)
    # a first VelocityEqualityConstraint is not doubled
    blk2 = orc.Block(orc.VELEQ, e, J, np.zeros((5, 3)), 1.0, target=e)
    v2, _ = orc.pinv_step([blk2], 6)
    assert np.allclose(v2, np.einsum("nij,nj->ni", P, e), rtol=1e-13, atol=1e-15)
