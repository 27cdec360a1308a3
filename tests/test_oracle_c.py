"""The C restatement (oracle/clik_oracle.c, used as the timed CPU baseline) agrees with the NumPy
oracle, and the oracle's independent geometric FK agrees with the expression graph."""
import os
import sys

import numpy as np
import pytest

from oracle_bridge import orc, oracle_pinv, close
from casclik_b200 import scenarios, fk

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import c_port  # noqa: E402


def test_c_port_matches_numpy_oracle_ur5_track():
    sc = scenarios.get("ur5_track")
    inp = sc.sample(2000, seed=11)
    ref_v, _ = oracle_pinv(sc.spec, inp)
    chain = orc.load_chain(fk.UR5_URDF, "base_link", "tool0")
    got, used = c_port.pinv_track(chain, inp["q"], inp["y"], gain=1.0, lam=1e-7, threads=2)
    assert used == 2
    assert close(got, ref_v, 1e-9, 1e-12).all(), np.abs(got - ref_v).max()


def test_independent_fk_matches_expression_graph():
    sc = scenarios.get("ur5_track")
    inp = sc.sample(500, seed=4)
    chain = orc.load_chain(fk.UR5_URDF, "base_link", "tool0")
    p, J = orc.position_jacobian(chain, inp["q"].T.copy())
    from oracle_bridge import blocks_from_skill
    blocks, n = blocks_from_skill(sc.spec, inp["t"], inp["q"], None, inp["y"])
    assert n == 6
    assert np.abs(blocks[0].e - (p - inp["y"].T)).max() < 1e-14
    assert np.abs(blocks[0].J - J).max() < 1e-14      # AD Jacobian == geometric Jacobian
    assert np.all(blocks[0].Jt == 0.0)


@pytest.mark.parametrize("name", ["ur5_track", "ur5_moe2016_pinv", "iiwa_multitask"])
def test_generic_c_port_reproduces_the_numpy_oracle_pinv(name):
    """Generic part of clik_oracle.c (generated expression C + literal per-mode algebra + mode search,
    pseudo_inverse.py:259-451, :512-556): the timed CPU baseline of every pinv scenario."""
    from casclik_b200 import scenarios
    from oracle_bridge import oracle_pinv, close
    sc = scenarios.get(name)
    port = c_port.PinvPort(sc.spec)
    inp = sc.sample(1500, seed=4)
    v, mode, used = port.solve(inp, threads=2)
    ref_v, ref_mode = oracle_pinv(sc.spec, inp)
    assert used == 2 and np.array_equal(mode, ref_mode)
    assert len(np.unique(ref_mode)) >= (1 if name == "ur5_track" else 4)
    assert close(v, ref_v, 1e-9, 1e-12).all(), np.abs(v - ref_v).max()


@pytest.mark.parametrize("name", ["ur5_qp", "ur5_moe2016_qp"])
def test_generic_c_port_reproduces_the_numpy_oracle_qp(name):
    """clik_ref_qp_instance: line-by-line port of clik_oracle.py:solve_qp_single (reactive_qp.py:461-528
    with the conic solver restated as a dual active-set method)."""
    from casclik_b200 import scenarios
    from oracle_bridge import oracle_qp_problem, orc
    sc = scenarios.get(name)
    port = c_port.QpPort(sc.spec)
    inp = sc.sample(400, seed=4)
    sol, status, active, _ = port.solve(inp, threads=2)
    h, A, lb, ub = oracle_qp_problem(sc.spec, inp)
    xo, lamo, sto = orc.solve_qp(h, A, lb, ub)
    assert np.array_equal(status, sto) and np.all(sto == 0)
    assert np.abs(sol.T - xo).max() < 1e-8
    m = A.shape[1]
    up = np.array([sum(1 << r for r in range(m) if l[r] > 0) for l in lamo], dtype=np.uint32)
    lo = np.array([sum(1 << r for r in range(m) if l[r] < 0) for l in lamo], dtype=np.uint32)
    assert np.array_equal(active[0], up) and np.array_equal(active[1], lo)


def test_generic_c_port_qp_on_fuzzed_problems_incl_infeasible():
    from fuzz_skills import make_qp_skill
    from oracle_bridge import oracle_qp_problem, orc
    for seed in range(4):
        spec, weights, inp = make_qp_skill(seed)
        port = c_port.QpPort(spec, **weights)
        inp = {k: v[..., :120] for k, v in inp.items()}
        sol, status, active, _ = port.solve(inp, threads=1)
        w = {"w_rob": weights["robot_var_weights"]} if "robot_var_weights" in weights else {}
        if port.nx and inp.get("x") is None:
            inp = dict(inp, x=np.zeros((port.nx, 120)))
        h, A, lb, ub = oracle_qp_problem(spec, inp, **w)
        xo, lamo, sto = orc.solve_qp(h, A, lb, ub)
        assert np.array_equal(status, sto), seed
        ok = sto == 0
        assert np.abs(sol.T[ok] - xo[ok]).max() < 1e-7 * (1 + np.abs(xo[ok]).max())
