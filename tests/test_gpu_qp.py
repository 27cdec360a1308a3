"""Parity of the CUDA reactive-QP path (through the C ABI) with the oracle.

Tolerances (BASELINE.json:north_star): objective and constraint residual within 1e-6; the active
set must match the oracle's bit for bit.  On top of that every solution is certified with the
solver-independent KKT residuals (the QP is strictly convex, so KKT <=> the unique optimum that
qpOASES converges to in the reference)."""
import numpy as np
import pytest

from oracle_bridge import orc, oracle_qp_problem
import casclik_b200 as cc
from casclik_b200 import cs, scenarios, runtime

pytestmark = pytest.mark.gpu
TOL = 1e-6


def _torch():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    return torch


def _solve_device(ctrl, inp, warm=None):
    torch = _torch()
    up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    sol, status, active = ctrl.solve_batch(up(inp["t"]), up(inp["q"]), up(inp.get("x")),
                                           up(inp.get("y")), warmstart=up(warm))
    torch.cuda.synchronize()
    return sol.cpu().numpy(), status.cpu().numpy(), active.cpu().numpy().astype(np.uint32)


def _masks(lam):
    m = lam.shape[1]
    up = np.zeros(lam.shape[0], dtype=np.uint32)
    lo = np.zeros(lam.shape[0], dtype=np.uint32)
    for r in range(min(m, 32)):
        up |= (lam[:, r] > 0).astype(np.uint32) << np.uint32(r)
        lo |= (lam[:, r] < 0).astype(np.uint32) << np.uint32(r)
    return up, lo


def _check_against_oracle(spec, ctrl, inp, n_check=None):
    h, A, lb, ub = oracle_qp_problem(spec, inp)
    sol, status, active = _solve_device(ctrl, inp)
    N = A.shape[0]
    idx = np.arange(N) if n_check is None else np.arange(0, N, max(1, N // n_check))
    xo, lamo, sto = orc.solve_qp(h, A[idx], lb[idx], ub[idx])
    assert np.all(sto == 0)
    assert np.all(status[idx] == runtime.QP_SOLVED)
    xs = sol[:, idx].T
    obj = 0.5 * np.sum(h * xs * xs, axis=1)
    obj_o = 0.5 * np.sum(h * xo * xo, axis=1)
    assert np.all(np.abs(obj - obj_o) <= TOL * (1 + np.abs(obj_o))), np.abs(obj - obj_o).max()
    r = np.einsum("nij,nj->ni", A[idx], xs)
    scale = 1 + np.abs(r)
    assert np.all(lb[idx] - r <= TOL * scale) and np.all(r - ub[idx] <= TOL * scale)
    assert np.abs(xs - xo).max() <= 1e-7 * (1 + np.abs(xo).max())
    up_o, lo_o = _masks(lamo)
    assert np.array_equal(active[0, idx], up_o), "upper active-set flags differ"
    assert np.array_equal(active[1, idx], lo_o), "lower active-set flags differ"
    # solver-independent certificate on every checked instance
    for k, i in enumerate(idx[:256]):
        kk = orc.kkt_residuals(h, A[i], lb[i], ub[i], xs[k])
        assert kk["primal"] < 1e-8 and kk["stationarity"] < 1e-8 and kk["sign"] < 1e-8, kk
    return sol, status, active


def test_ur5_qp_parity_and_kkt():
    sc = scenarios.get("ur5_qp")
    ctrl = sc.make_controller()
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    inp = sc.sample(2048, seed=0)
    sol, status, active = _check_against_oracle(sc.spec, ctrl, inp)
    assert (active[0] | active[1]).any(), "some speed limits must be active with K = 50"
    # host ABI == device ABI, bit for bit
    sh, sth, ah = ctrl.solve_batch(inp["t"], inp["q"], None, inp["y"])
    assert np.array_equal(sh, sol) and np.array_equal(sth, status)
    assert np.array_equal(ah.astype(np.uint32), active)
    torch = _torch()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    sz, stz, az = ctrl.solve_batch(pin(inp["t"]), pin(inp["q"]), None, pin(inp["y"]))   # outputs pageable: staged
    out = (pin(np.zeros_like(sol)), pin(np.zeros_like(status)), pin(np.zeros((2, sol.shape[1]), dtype=np.int32)))
    sz2, stz2, az2 = ctrl.solve_batch(pin(inp["t"]), pin(inp["q"]), None, pin(inp["y"]), out=out)   # zero copy
    assert np.array_equal(sz, sol) and np.array_equal(sz2, sol) and np.array_equal(stz2, status)
    assert np.array_equal(az2.astype(np.uint32), active)
    # warm starts (primal guess as the reference's x0=, or the previous working set) only shorten
    # the iteration: same minimiser (to rounding: the iterates differ), same flags
    sw, stw, aw = _solve_device(ctrl, inp, warm=sol)
    assert np.all(stw == 0) and np.abs(sw - sol).max() < 1e-11 and np.array_equal(aw, active)
    torch = _torch()
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    s2, st2, a2 = ctrl.solve_batch(up(inp["t"]), up(inp["q"]), None, up(inp["y"]),
                                   warm_active=up(active.astype(np.int32)))
    assert bool((st2 == 0).all()) and np.abs(s2.cpu().numpy() - sol).max() < 1e-11
    assert np.array_equal(a2.cpu().numpy().astype(np.uint32), active)
    # a deliberately wrong guess (everything at its upper bound) still gives the right answer
    bad = np.zeros_like(active, dtype=np.int32)
    bad[0, :] = 0x7fff
    s3, st3, a3 = ctrl.solve_batch(up(inp["t"]), up(inp["q"]), None, up(inp["y"]), warm_active=up(bad))
    assert bool((st3 == 0).all()) and np.abs(s3.cpu().numpy() - sol).max() < 1e-11
    assert np.array_equal(a3.cpu().numpy().astype(np.uint32), active)


def test_moe2016_qp_time_varying_parity():
    sc = scenarios.get("ur5_moe2016_qp")
    ctrl = sc.make_controller()
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    _check_against_oracle(sc.spec, ctrl, sc.sample(1024, seed=2))


def _cart():
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    eq = cc.EqualityConstraint("min_dist_cnstr", 0.75 - p, gain=1.0, constraint_type="soft", priority=1)
    lim = cc.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0)
    spd = cc.VelocitySetConstraint("speed_limit_cnstr", p, gain=10.0, set_min=-0.275, set_max=0.275)
    return cc.SkillSpecification("cart_qp", t, p, robot_vel_var=dp, constraints=[eq, lim, spd])


def test_cart_known_answers_through_single_instance_api():
    ctrl = cc.ReactiveQPController(skill_spec=_cart(), robot_var_weights=[1.0])
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    rv, vv, sl = ctrl.solve(time_var=0.0, robot_var=0.0)                    # KAT Q1
    assert vv is None
    assert abs(float(rv) - 0.275) < 1e-12 and abs(float(sl) - 0.475) < 1e-12
    assert abs(ctrl.res["cost"] - 0.112963125) < 1e-12
    assert ctrl.res["active_upper"] == 0b101 and ctrl.res["active_lower"] == 0     # eq row + speed row, both held from above
    rv, _, sl = ctrl.solve(0.0, 0.6, warmstart_slack_var=[0.475])            # KAT Q2
    assert abs(float(rv) - 0.14985029940119762) < 1e-12
    assert abs(float(sl) - 1.4970059880239917e-4) < 1e-12
    # initial-value problem (robot velocity fixed at 0): slack = 0.75 - p
    ctrl.setup_initial_problem_solver()
    virt0, slack0 = ctrl.solve_initial_problem(0.0, 0.25)
    assert virt0 is None and abs(float(slack0) - 0.5) < 1e-12


def test_virtual_variable_qp():
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    x, dx = cs.MX.sym("x"), cs.MX.sym("dx")
    up = cc.EqualityConstraint("move_up_path_cnstr", 300 - x, gain=1.0, constraint_type="soft", priority=1)
    slow = cc.VelocitySetConstraint("slow_path_cnstr", x, set_min=-0.5, set_max=0.5)
    dist = cc.EqualityConstraint("min_dist_cnstr", 0.4 * cs.sin(0.3 * x) - p, gain=1.0,
                                 constraint_type="soft", priority=1)
    lim = cc.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0)
    spd = cc.VelocitySetConstraint("speed_limit_cnstr", p, set_min=-0.275, set_max=0.275)
    spec = cc.SkillSpecification("path_trajectory_skill", t, p, robot_vel_var=dp, virtual_var=x,
                                 virtual_vel_var=dx, constraints=[up, slow, dist, lim, spd])
    ctrl = cc.ReactiveQPController(spec, robot_var_weights=[1.0])
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    rng = np.random.default_rng(1)
    N = 300
    inp = {"t": np.zeros(N), "q": rng.uniform(0.0, 1.0, (1, N)), "x": rng.uniform(0, 10, (1, N))}
    _check_against_oracle(spec, ctrl, inp)
    rv, vv, sl = ctrl.solve(0.0, 0.3, virtual_var=1.0)
    assert vv is not None and sl.shape == (2, 1)
    assert abs(float(vv) - 0.5) < 1e-9      # path variable saturates its speed limit


def test_infeasible_and_iteration_cap_are_status_flags(monkeypatch):
    t, p = cs.MX.sym("t"), cs.MX.sym("p")
    a = cc.VelocitySetConstraint("a", p, set_min=1.0, set_max=2.0)
    b = cc.VelocitySetConstraint("b", p, set_min=-3.0, set_max=-2.0)
    spec = cc.SkillSpecification("infeasible", t, p, constraints=[a, b])
    ctrl = cc.ReactiveQPController(spec)
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    sol, status, _ = ctrl.solve_batch(np.zeros(3), np.zeros((1, 3)))
    assert np.all(status == runtime.QP_INFEASIBLE)
    with pytest.raises(RuntimeError):
        ctrl.solve(0.0, 0.0)
    # iteration cap: with the crash start switched off a cold solve of the UR5 problem needs ~10
    # iterations, so a cap of 1 must be reported (status 1), never a wrong "solved"
    sc = scenarios.get("ur5_qp")
    monkeypatch.setenv("CLIK_QP_CRASH", "0")
    c2 = sc.make_controller()
    c2.setup_problem_functions()
    inp = sc.sample(64, seed=0)
    _, st, _ = c2.solve_batch(inp["t"], inp["q"], None, inp["y"], max_iter=1)
    assert np.all((st == runtime.QP_MAXITER) | (st == runtime.QP_SOLVED)) and (st == runtime.QP_MAXITER).any()
    monkeypatch.delenv("CLIK_QP_CRASH")
    c3 = sc.make_controller()
    c3.setup_problem_functions()
    sol3, st3, _ = c3.solve_batch(inp["t"], inp["q"], None, inp["y"], max_iter=1)
    assert np.all((st3 == runtime.QP_MAXITER) | (st3 == runtime.QP_SOLVED))
    full, stf, _ = c3.solve_batch(inp["t"], inp["q"], None, inp["y"])
    ok = st3 == runtime.QP_SOLVED
    assert np.all(stf == runtime.QP_SOLVED) and ok.mean() > 0.9            # the guess is usually exact ...
    assert np.array_equal(np.asarray(sol3)[:, ok], np.asarray(full)[:, ok])  # ... and a reported "solved" is the answer


def test_conic_object_dense_random_problems():
    """cs.conic-style call on numeric matrices (clik_qp_dense) vs oracle + KKT, incl. +-inf and
    +-1e10 default bounds (SetConstraint defaults, reference constraints.py:199-206)."""
    solver = cs.conic("solver", "qpoases", {}, {})
    rng = np.random.default_rng(5)
    N, n, m = 200, 7, 12
    h = rng.uniform(0.001, 2.0, (N, n))
    A = rng.normal(size=(N, m, n))
    r = np.einsum("nij,nj->ni", A, rng.normal(size=(N, n)))
    lb, ub = r - rng.uniform(0, 1, (N, m)), r + rng.uniform(0, 1, (N, m))
    lb[:, 0], ub[:, 1] = -np.inf, np.inf
    lb[:, 2], ub[:, 2] = -1e10, 1e10
    lb[:, 3] = ub[:, 3] = r[:, 3]
    x, status, active = solver.solve_dense_batch(h, A, lb, ub)
    assert np.all(status == 0)
    for i in range(N):
        xo, lamo, sto = orc.solve_qp_single(h[i], A[i], lb[i], ub[i])
        assert sto == 0 and np.abs(x[i] - xo).max() < 1e-8 * (1 + np.abs(xo).max())
        kk = orc.kkt_residuals(h[i], A[i], lb[i], ub[i], x[i])
        assert kk["primal"] < 1e-8 and kk["stationarity"] < 1e-8 and kk["sign"] < 1e-8
    res = solver(h=np.diag(h[0]), a=A[0], lba=lb[0], uba=ub[0])
    assert np.abs(res["x"].toarray()[:, 0] - x[0]).max() == 0.0


def test_qp_rollout_on_device_matches_stepwise_loop():
    torch = _torch()
    sc = scenarios.get("ur5_moe2016_qp")
    ctrl = sc.make_controller()
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    N, K, dt, vmax = 128, 25, 0.008, np.pi / 5
    inp = sc.sample(N, seed=4)
    t0 = torch.from_numpy(inp["t"]).cuda()
    q_roll = torch.from_numpy(inp["q"]).cuda()
    q_loop = q_roll.clone()
    out = ctrl.rollout_batch(t0, q_roll, K, dt, max_speed=vmax)
    act = None
    for k in range(K):
        # the rollout kernel warm-starts each step from the previous working set: do the same
        sol, status, act = ctrl.solve_batch(t0 + dt * k, q_loop, warm_active=act)
        assert bool((status == 0).all())
        v = torch.clamp(sol[:6], -vmax, vmax)
        q_loop = q_loop + v * dt
    torch.cuda.synchronize()
    assert torch.equal(q_roll, q_loop)
    assert torch.equal(out["sol"][:6], v) and int(out["n_failed"].sum()) == 0


def test_kitchen_sink_qp_parity():
    """Expression-valued weights and bounds, matrix gains, hard + soft rows of every constraint
    class, virtual and input variables (4 dense rows: structured solver with the crash start)."""
    t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", 3), cs.MX.sym("dq", 3)
    x, dx, y = cs.MX.sym("x"), cs.MX.sym("dx"), cs.MX.sym("y", 2)
    c1 = cc.EqualityConstraint("soft_eq", cs.vertcat(cs.sin(q[0]) + q[1] - y[0], q[2] * q[0] - y[1] + 0.1 * x),
                               gain=np.array([[2.0, 0.3], [0.0, 1.5]]), constraint_type="soft", slack_weight=3.0)
    c2 = cc.SetConstraint("soft_set", q[1] + 0.2 * cs.sin(t), gain=3.0, set_min=-0.4, set_max=cs.MX(0.5) + 0.05 * y[1],
                          constraint_type="soft")
    c3 = cc.VelocityEqualityConstraint("hard_veleq", q[0] + 0.5 * q[2] + x, target=0.2 * cs.sin(t))
    c4 = cc.VelocitySetConstraint("speed", q, set_min=-0.8 * np.ones(3), set_max=0.8 * np.ones(3))
    c5 = cc.VelocitySetConstraint("vspeed", x, set_min=-0.5, set_max=0.5)
    spec = cc.SkillSpecification("qp_sink", t, q, robot_vel_var=dq, virtual_var=x, virtual_vel_var=dx,
                                 input_var=y, constraints=[c1, c2, c3, c4, c5])
    ctrl = cc.ReactiveQPController(spec, robot_var_weights=[1.0, 2.0, 0.5], virtual_var_weights=[4.0])
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    rng = np.random.default_rng(3)
    N = 1500
    inp = {"t": rng.uniform(0, 10, N), "q": rng.uniform(-0.6, 0.6, (3, N)), "x": rng.uniform(-1, 1, (1, N)),
           "y": rng.uniform(-0.5, 0.5, (2, N))}
    h, A, lb, ub = oracle_qp_problem(spec, inp, w_rob=[1.0, 2.0, 0.5], w_virt=[4.0])
    assert np.allclose(h, [0.001, 0.002, 0.0005, 0.004, 3.001, 3.001, 1.001])
    sol, status, active = _solve_device(ctrl, inp)
    assert np.all(status == 0)
    for i in range(0, N, 7):
        xo, lamo, sto = orc.solve_qp_single(h, A[i], lb[i], ub[i])
        assert sto == 0 and np.abs(sol[:, i] - xo).max() < 1e-7 * (1 + np.abs(xo).max())
        kk = orc.kkt_residuals(h, A[i], lb[i], ub[i], sol[:, i])
        assert kk["primal"] < 1e-8 and kk["stationarity"] < 1e-8 and kk["sign"] < 1e-8
        up_o, lo_o = _masks(lamo[None, :])
        assert active[0, i] == up_o[0] and active[1, i] == lo_o[0]


def test_two_launch_split_equals_the_single_kernel(monkeypatch):
    """clik_qp_step with a status array runs the working-set prediction and the full solver as two
    launches (handing over through status[]); CLIK_QP_SPLIT=0 keeps everything in one kernel.  Same
    bits, for cold, x0-warm and active-set-warm calls, on ragged sizes that leave partial tiles."""
    torch = _torch()
    sc = scenarios.get("ur5_qp")
    N = 5000 + 37
    inp = sc.sample(N, seed=9)
    dev = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v)).cuda()) for k, v in inp.items()}
    outs = {}
    for split in ("1", "0"):
        monkeypatch.setenv("CLIK_QP_SPLIT", split)
        ctrl = sc.make_controller()
        ctrl.setup_problem_functions()
        ctrl.setup_solver()
        assert bool(ctrl.kernel_meta.get("qp_split")) == (split == "1")
        cold = ctrl.solve_batch(dev["t"], dev["q"], None, dev["y"])
        warm_x = ctrl.solve_batch(dev["t"], dev["q"], None, dev["y"], warmstart=cold[0])
        warm_a = ctrl.solve_batch(dev["t"], dev["q"], None, dev["y"], warm_active=cold[2])
        outs[split] = [tuple(t.clone() for t in r) for r in (cold, warm_x, warm_a)]
    for a, b in zip(outs["1"], outs["0"]):
        for ta, tb in zip(a, b):
            assert torch.equal(ta, tb)
    assert int(outs["1"][0][1].abs().sum()) == 0          # every status final (0), none left pending


@pytest.mark.parametrize("name", ["ur5_qp", "ur5_moe2016_qp"])
def test_capped_tail_kernel_gives_the_bits_of_the_uncapped_one(name, monkeypatch):
    """The tail pass exists twice — natural register allocation, and capped to 4 CTAs per SM for batches with
    more tail tiles than resident tail CTAs (clik_abi.cu picks by batch size; CLIK_QP_TAIL_PICK forces one).
    Same source, different register allocation: identical results, and the default pick is one of them."""
    torch = _torch()
    sc = scenarios.get(name)
    N = 40_000 + 11
    inp = sc.sample(N, seed=12)
    dev = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v)).cuda()) for k, v in inp.items()}
    outs = {}
    for pick in ("0", "1", None):
        if pick is None:
            monkeypatch.delenv("CLIK_QP_TAIL_PICK")
        else:
            monkeypatch.setenv("CLIK_QP_TAIL_PICK", pick)
        ctrl = sc.make_controller()
        ctrl.setup_solver()
        assert ctrl._skill(0).launch_info(7)["regs"] <= 128 < ctrl._skill(0).launch_info(4)["regs"]
        outs[pick] = [t.clone() for t in ctrl.solve_batch(dev["t"], dev["q"], dev.get("x"), dev["y"])]
    for a, b, c in zip(outs["0"], outs["1"], outs[None]):
        assert torch.equal(a, b) and torch.equal(a, c)
    assert int(outs["0"][1].abs().sum()) == 0


def test_non_finite_data_is_reported_not_returned_as_solved():
    """NaN inputs / a zero weight used to come back with status 0 and NaN velocities."""
    torch = _torch()
    sc = scenarios.get("ur5_qp")
    ctrl = sc.make_controller()
    ctrl.setup_solver()
    inp = sc.sample(64, seed=3)
    inp["q"][2, 5] = np.nan
    inp["y"][0, 9] = np.inf
    sol, status, _ = _solve_device(ctrl, inp)
    assert status[5] == runtime.QP_INVALID and status[9] != runtime.QP_SOLVED
    ok = np.ones(64, dtype=bool)
    ok[[5, 9]] = False
    assert np.all(status[ok] == 0) and np.isfinite(sol[:, ok]).all()
    with pytest.raises(RuntimeError):
        ctrl.solve(0.0, inp["q"][:, 5], input_var=inp["y"][:, 5])
    with pytest.raises(ValueError):
        bad = cc.ReactiveQPController(sc.spec, robot_var_weights=[1.0, 1.0, 0.0, 1.0, 1.0, 1.0])
        bad.setup_problem_functions(load=False)
    # dense conic entry: zero weight -> status 4, never a silent NaN; and the second capacity tier (24 x 40)
    from casclik_b200.controllers.qp_solver import ConicSolver
    solver = ConicSolver("solver", "qpoases", {}, {})
    rng = np.random.default_rng(0)
    h = np.ones((3, 4)); h[1, 2] = 0.0
    A = rng.normal(size=(3, 5, 4)); lb = -np.ones((3, 5)); ub = np.ones((3, 5))
    x, st, _ = solver.solve_dense_batch(h, A, lb, ub)
    assert st[0] == 0 and st[2] == 0 and st[1] == runtime.QP_INVALID
    nx, m, N = 24, 40, 6
    h2 = rng.uniform(0.5, 2.0, (N, nx)); A2 = rng.normal(size=(N, m, nx))
    c = rng.normal(size=(N, m)); lb2, ub2 = c - 0.3, c + 0.3
    lb2[:, 20:], ub2[:, 20:] = -1e10, 1e10                       # 20 binding-ish rows, 20 free
    x2, st2, _ = solver.solve_dense_batch(h2, A2, lb2, ub2)
    assert np.all(st2 == 0)
    for i in range(N):
        xo, lam, so = orc.solve_qp_single(h2[i], A2[i], lb2[i], ub2[i], max_iter=2000)
        assert so == 0 and np.abs(x2[i] - xo).max() < 1e-7 * (1 + np.abs(xo).max())
