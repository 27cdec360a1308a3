"""Parity of the CUDA pseudo-inverse path (through the C ABI) against the oracle, on seeded
synthetic inputs of the BASELINE.json configs.  Tolerance for joint-velocity commands is the one
BASELINE.json:north_star states: |a - b| <= 1e-12 + 1e-9*|b| element-wise; mode flags bit-exact."""
import numpy as np
import pytest

from oracle_bridge import oracle_pinv, close
import casclik_b200 as cc
from casclik_b200 import cs, scenarios

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-9, 1e-12


def _torch():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    return torch


def _run_device(ctrl, inp):
    torch = _torch()
    dev = torch.device("cuda", 0)
    up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    qd, xd, mode = ctrl.solve_batch(up(inp["t"]), up(inp["q"]), up(inp.get("x")), up(inp.get("y")))
    torch.cuda.synchronize()
    return qd.cpu().numpy(), (None if xd is None else xd.cpu().numpy()), mode.cpu().numpy()


def _setup(name):
    sc = scenarios.get(name)
    ctrl = sc.make_controller()
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    return sc, ctrl


def _assert_parity_multitask(v, spec, inp, opts, what, rtol=RTOL, atol=ATOL):
    """Multi-task chains stack [J; J; ...] (Appendix A1/A3): the damped Gram matrix of the stack has
    condition ~sigma^2/lam ~ 1e7, and two backward-stable float64 evaluations of the reference's own
    formula (LU vs Cholesky vs QR) differ by up to ~1.5e-8 element-wise / 7e-10 norm-wise (SURVEY
    §7.2 item 3).  Criterion: element-wise north-star tolerance for >= 99.9 % of the entries,
    norm-wise tolerance for every instance, and every out-of-tolerance instance must be as close
    to an extended-precision evaluation of the reference formulas as the float64 oracle is (x10)."""
    ref64, _ = oracle_pinv(spec, inp, dict(opts))
    ok = close(v, ref64, rtol, atol)
    if ok.all():
        return
    assert ok.mean() >= 0.999, _report(v, ref64, what)
    nerr = np.linalg.norm(v - ref64, axis=0)
    nref = np.linalg.norm(ref64, axis=0)
    assert (nerr <= atol + rtol * nref).all(), "%s: norm-wise %.3e" % (what, (nerr / (nref + 1e-300)).max())
    bad = np.nonzero(~ok.all(axis=0))[0]
    sub = {k: (a[..., bad] if a is not None else None) for k, a in inp.items()}
    refld, _ = oracle_pinv(spec, sub, dict(opts), dtype=np.longdouble)
    refld = refld.astype(np.float64)
    e_gpu = np.abs(v[:, bad] - refld).max(axis=0)
    e_o64 = np.abs(ref64[:, bad] - refld).max(axis=0)
    assert (e_gpu <= atol + 10.0 * np.maximum(e_o64, 1e-11 * np.abs(refld).max(axis=0))).all(), \
        "%s: gpu %.3e vs fp64 oracle %.3e from the extended-precision referee" % (what, e_gpu.max(), e_o64.max())


def _report(v, ref, what):
    err = np.abs(v - ref)
    bad = ~close(v, ref, RTOL, ATOL)
    rel = err / (np.abs(ref) + 1e-300)
    return "%s: max abs err %.3e, max rel err %.3e, %d / %d entries out of tolerance" % (
        what, err.max(), rel[np.abs(ref) > 1e-6].max() if np.any(np.abs(ref) > 1e-6) else 0.0,
        int(bad.sum()), bad.size)


def test_ur5_track_parity_device_and_host_paths():
    sc, ctrl = _setup("ur5_track")
    inp = sc.sample(4096, seed=0)
    ref_v, ref_mode = oracle_pinv(sc.spec, inp)
    v, _, mode = _run_device(ctrl, inp)
    assert np.array_equal(mode, ref_mode) and np.all(mode == 0)
    assert close(v, ref_v, RTOL, ATOL).all(), _report(v, ref_v, "ur5_track device")
    # host-buffer ABI (pipelined H2D / kernel / D2H) gives the same bits as the device ABI
    vh, _, mh = ctrl.solve_batch(inp["t"], inp["q"], None, inp["y"])
    assert np.array_equal(vh, v) and np.array_equal(mh, mode)
    # page-locked host buffers: the kernel runs directly on the mapped host memory (zero copy)
    torch = _torch()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    out = (pin(np.zeros_like(v)), None, pin(np.zeros_like(mode)))
    vz, _, mz = ctrl.solve_batch(pin(inp["t"]), pin(inp["q"]), None, pin(inp["y"]), out=out)
    assert np.array_equal(vz, v) and np.array_equal(mz, mode)


def test_ur5_track_single_instance_api_matches_reference_conventions():
    sc, ctrl = _setup("ur5_track")
    inp = sc.sample(3, seed=5)
    ref_v, _ = oracle_pinv(sc.spec, inp)
    for i in range(3):
        res = ctrl.solve(0.0, inp["q"][:, i], input_var=inp["y"][:, i])
        assert len(res) == 3 and res[1] is None and res[2] is None
        got = res[0].toarray()[:, 0]
        assert got.shape == (6,)
        assert close(got, ref_v[:, i], RTOL, ATOL).all()
        assert ctrl.current_mode == 0


def test_ragged_and_empty_batches():
    sc, ctrl = _setup("ur5_track")
    torch = _torch()
    for N in (1, 31, 129, 1000):
        inp = sc.sample(N, seed=N)
        ref_v, _ = oracle_pinv(sc.spec, inp)
        v, _, mode = _run_device(ctrl, inp)
        assert v.shape == (6, N) and close(v, ref_v, RTOL, ATOL).all()
    q0 = torch.empty((6, 0), dtype=torch.float64, device="cuda")
    y0 = torch.empty((3, 0), dtype=torch.float64, device="cuda")
    qd, _, mode = ctrl.solve_batch(0.0, q0, None, y0)
    assert qd.shape == (6, 0) and mode.shape == (0,)


def test_scalar_time_broadcast_equals_per_instance_time():
    sc, ctrl = _setup("ur5_moe2016_pinv")
    torch = _torch()
    inp = sc.sample(512, seed=2)
    q = torch.from_numpy(inp["q"]).cuda()
    a = ctrl.solve_batch(17.5, q)
    b = ctrl.solve_batch(torch.full((512,), 17.5, dtype=torch.float64, device="cuda"), q)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])


@pytest.mark.parametrize("name", ["ur5_moe2016_pinv", "iiwa_multitask"])
def test_set_based_modes_parity(name):
    sc, ctrl = _setup(name)
    inp = sc.sample(4096, seed=1)
    ref_v, ref_mode = oracle_pinv(sc.spec, inp)
    v, _, mode = _run_device(ctrl, inp)
    assert len(np.unique(ref_mode)) > 3, "inputs must exercise several modes"
    assert np.array_equal(mode, ref_mode), "mode flags differ in %d instances" % int((mode != ref_mode).sum())
    assert close(v, ref_v, RTOL, ATOL).all(), _report(v, ref_v, name)


def test_rank_deficient_stress_skill_is_as_accurate_as_the_fp64_reference_formula():
    """iiwa 9-row three-point pose task: J (9x7) has rank 6, so J'J + 1e-7 I has condition number
    ~1e8 and the reference's own float64 formula is only reproducible to ~1e-6 (two backward-
    stable solvers already differ by that much).  Criterion: mode flags
    bit-exact; the CUDA result is as close to an extended-precision evaluation of the
    reference's formulas (same float64 e/J/des inputs) as the float64 oracle is, up to a factor
    10, on top of the north-star tolerance."""
    sc, ctrl = _setup("iiwa_multitask_stress")
    inp = sc.sample(1024, seed=1)
    ref64, mode64 = oracle_pinv(sc.spec, inp)
    refld, modeld = oracle_pinv(sc.spec, inp, dtype=np.longdouble)
    v, _, mode = _run_device(ctrl, inp)
    assert np.array_equal(mode, mode64) and np.array_equal(mode64, modeld)
    refld64 = refld.astype(np.float64)
    scale = np.linalg.norm(refld64, axis=0)
    e_gpu = np.linalg.norm(v - refld64, axis=0) / scale
    e_o64 = np.linalg.norm(ref64 - refld64, axis=0) / scale
    # rounding noise amplified by ~1e8 is random per instance, so compare the error distributions
    for qt in (0.5, 0.9, 0.99):
        assert np.quantile(e_gpu, qt) <= RTOL + 10.0 * np.quantile(e_o64, qt), \
            "quantile %.2f: gpu %.3e vs fp64 oracle %.3e" % (qt, np.quantile(e_gpu, qt), np.quantile(e_o64, qt))
    assert e_gpu.max() <= RTOL + 30.0 * e_o64.max(), "max: gpu %.3e vs fp64 oracle %.3e" % (e_gpu.max(), e_o64.max())
    assert e_gpu.max() < 1e-6            # absolute sanity: far below the 2.9e-5 effect of the A1 quirk
    # and the float64 oracle really is that far from the exact value (the test is not vacuous)
    assert e_o64.max() > 1e-9


def test_multidim_sets_option_parity():
    """options["multidim_sets"] (experimental in the reference, pseudo_inverse.py:192-257, :289-298):
    vector-valued SetConstraint, row-masked reduced Jacobian, vector in-tangent-cone test."""
    sc, ctrl = _setup("ur5_moe2016_multidim")
    assert ctrl.n_modes == 2
    inp = sc.sample(4096, seed=6)
    ref_v, ref_mode = oracle_pinv(sc.spec, inp, {"multidim_sets": True})
    v, _, mode = _run_device(ctrl, inp)
    assert set(np.unique(ref_mode)) >= {0, 1}
    assert np.array_equal(mode, ref_mode), "mode flags differ in %d instances" % int((mode != ref_mode).sum())
    assert close(v, ref_v, RTOL, ATOL).all(), _report(v, ref_v, "multidim")
    # scalar sets under multidim_sets also use the masked reduced Jacobian
    sc2 = scenarios.get("ur5_moe2016_pinv")
    c2 = cc.PseudoInverseController(sc2.spec, options={"multidim_sets": True})
    c2.setup_solver()
    inp2 = sc2.sample(2048, seed=8)
    ref2, mode2 = oracle_pinv(sc2.spec, inp2, {"multidim_sets": True})
    v2, _, m2 = _run_device(c2, inp2)
    assert np.array_equal(m2, mode2) and close(v2, ref2, RTOL, ATOL).all(), _report(v2, ref2, "scalar sets, multidim")
    ref_plain, _ = oracle_pinv(sc2.spec, inp2)
    assert np.abs(ref2 - ref_plain).max() > 1e-6          # the option does change the answer


def test_converge_final_set_to_max_option_parity():
    """options["converge_final_set_to_max"] (pseudo_inverse.py:337-356): an active FINAL set is also
    driven to set_max through the null space of the constraints above it."""
    t, q = cs.MX.sym("t"), cs.MX.sym("q", 4)
    reach = cc.EqualityConstraint("reach", cs.vertcat(cs.sin(q[0]) + q[1] - 0.4 * cs.cos(0.2 * t), q[2] * q[3] - 0.1),
                                  gain=1.5, priority=1)
    lim = cc.SetConstraint("lim", q[1], set_min=-0.3, set_max=0.35, priority=2)
    last = cc.SetConstraint("final_set", q[0] + 0.5 * q[3], gain=2.0, set_min=-0.2, set_max=0.25, priority=3)
    spec = cc.SkillSpecification("conv", t, q, constraints=[last, reach, lim])
    opts = {"converge_final_set_to_max": True}
    ctrl = cc.PseudoInverseController(spec, options=dict(opts))
    ctrl.setup_solver()
    rng = np.random.default_rng(5)
    N = 3000
    inp = {"t": rng.uniform(0, 10, N), "q": rng.uniform(-0.6, 0.6, (4, N))}
    ref_v, ref_mode = oracle_pinv(spec, inp, dict(opts))
    v, _, mode = _run_device(ctrl, inp)
    assert set(np.unique(ref_mode)) >= {0, 1, 2, 3}
    assert np.array_equal(mode, ref_mode)
    _assert_parity_multitask(v, spec, inp, opts, "converge_final_set_to_max")
    plain, _ = oracle_pinv(spec, inp)
    assert np.abs(plain - ref_v).max() > 1e-3          # the option changes the command


def test_no_admissible_mode_returns_zero_and_minus_one():
    # p must stay in [0, 1] but the only task pushes it further out and the set itself cannot
    # produce motion (A5): with p = 2 and target 3 every mode is rejected or ...
    t, p = cs.MX.sym("t"), cs.MX.sym("p")
    lim1 = cc.SetConstraint("a", p, set_min=0.0, set_max=1.0, priority=1)
    lim2 = cc.SetConstraint("b", -p, set_min=-1.0, set_max=0.0, priority=2)
    eq = cc.EqualityConstraint("go", 3.0 - p, priority=3)
    spec = cc.SkillSpecification("stuck", t, p, constraints=[lim1, lim2, eq])
    ctrl = cc.PseudoInverseController(spec)
    ctrl.setup_solver()
    inp = {"t": np.zeros(4), "q": np.array([[2.0, 0.5, 2.0, -1.0]])}
    ref_v, ref_mode = oracle_pinv(spec, inp)
    v, _, mode = _run_device(ctrl, inp)
    assert np.array_equal(mode, ref_mode)
    assert close(v, ref_v, RTOL, ATOL).all()
    for i in np.nonzero(ref_mode < 0)[0]:
        assert v[0, i] == 0.0


def test_virtual_variable_and_velocity_equality():
    """Cart path-following skill of the notebooks (virtual path variable x)."""
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    x, dx = cs.MX.sym("x"), cs.MX.sym("dx")
    up = cc.EqualityConstraint("move_up_path_cnstr", 300 - x, gain=1.0, priority=1)
    lim = cc.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0, priority=1)
    dist = cc.EqualityConstraint("min_dist_cnstr", 0.4 * cs.sin(0.3 * x) - p, gain=1.0,
                                 constraint_type="soft", priority=3)
    vel = cc.VelocityEqualityConstraint("drift", p + 0.1 * x, target=0.05, priority=4)
    spec = cc.SkillSpecification("path", t, p, robot_vel_var=dp, virtual_var=x, virtual_vel_var=dx,
                                 constraints=[up, dist, lim, vel])
    ctrl = cc.PseudoInverseController(spec)
    ctrl.setup_solver()
    rng = np.random.default_rng(0)
    N = 257
    inp = {"t": np.zeros(N), "q": rng.uniform(-0.2, 1.2, (1, N)), "x": rng.uniform(0, 20, (1, N))}
    ref_v, ref_mode = oracle_pinv(spec, inp)
    v, xd, mode = _run_device(ctrl, inp)
    got = np.vstack([v, xd])
    assert np.array_equal(mode, ref_mode)
    assert close(got, ref_v, RTOL, ATOL).all(), _report(got, ref_v, "cart path")
    r = ctrl.solve(0.0, 0.3, virtual_var=2.0)
    assert r[1] is not None and r[1].shape == (1, 1)


def test_full_size_properties_ur5_track():
    """Config 2 at its BASELINE size (2^20): properties that do not need the oracle."""
    sc, ctrl = _setup("ur5_track")
    torch = _torch()
    N = 1 << 20
    inp = sc.sample(N, seed=0)
    q = torch.from_numpy(inp["q"]).cuda()
    y = torch.from_numpy(inp["y"]).cuda()
    v, _, mode = ctrl.solve_batch(0.0, q, None, y)
    torch.cuda.synchronize()
    assert bool((mode == 0).all()) and bool(torch.isfinite(v).all())
    # (1) shard equivalence: any contiguous slice run on its own gives the same bits
    for lo, hi in ((0, 1000), (12345, 70001), (N - 777, N)):
        vs, _, ms = ctrl.solve_batch(0.0, q[:, lo:hi].contiguous(), None, y[:, lo:hi].contiguous())
        assert torch.equal(vs, v[:, lo:hi])
    # (2) a first-order step along v reduces the task error for a small step (closed-loop sanity)
    from casclik_b200.sym import dag
    p_expr = sc.spec.constraints[0].expression
    nodes = p_expr.nodes()
    ids_q = [s.id for s in sc.spec.robot_var.nodes()]
    ids_y = [s.id for s in sc.spec.input_var.nodes()]

    sub = slice(0, 20000)

    def err(qq):
        vals = {i: qq[k] for k, i in enumerate(ids_q)}
        vals.update({i: inp["y"][k, sub] for k, i in enumerate(ids_y)})
        e = np.stack(dag.evaluate(nodes, vals))
        return np.sqrt((e * e).sum(axis=0))
    e0 = err(inp["q"][:, sub])
    e1 = err(inp["q"][:, sub] + 1e-3 * v[:, sub].cpu().numpy())
    assert (e1 < e0).mean() > 0.999
    # (3) oracle on a strided subsample of the full batch
    idx = np.arange(0, N, 509)
    ref_v, _ = oracle_pinv(sc.spec, {"t": inp["t"][idx], "q": inp["q"][:, idx], "y": inp["y"][:, idx]})
    got = v[:, torch.from_numpy(idx).cuda()].cpu().numpy()
    assert close(got, ref_v, RTOL, ATOL).all(), _report(got, ref_v, "ur5_track 2^20 subsample")


def test_huge_and_odd_inputs_take_the_fallback_paths():
    """Angles beyond the fast range reduction (|q| >= 1e5) go through the library sincos; an odd
    batch size cannot use 16-byte bulk copies and runs the plain kernel.  Same answers."""
    sc, ctrl = _setup("ur5_track")
    inp = sc.sample(1001, seed=9)                      # odd N: plain kernel
    inp["q"][0, ::7] += 2 * np.pi * 40000              # |q| ~ 2.5e5
    inp["q"][3, ::11] -= 2 * np.pi * 123456
    ref_v, _ = oracle_pinv(sc.spec, inp)
    v, _, _ = _run_device(ctrl, inp)
    assert close(v, ref_v, RTOL, ATOL).all(), _report(v, ref_v, "huge angles, odd N")
    inp2 = {k: (a[..., :1000] if a is not None else None) for k, a in inp.items()}
    v2, _, _ = _run_device(ctrl, inp2)
    assert np.array_equal(v2, v[:, :1000])
    # the opt-in TMA-staged persistent kernel (bulk async copies + mbarrier ring) gives the same bits
    import os
    os.environ["CLIK_TMA"] = "1"
    try:
        _, ctrl_tma = _setup("ur5_track")
        v3, _, m3 = _run_device(ctrl_tma, inp2)          # even N, aligned: TMA path
        v4, _, _ = _run_device(ctrl_tma, inp)            # odd N: falls back to the plain kernel
    finally:
        del os.environ["CLIK_TMA"]
    assert np.array_equal(v3, v2) and np.array_equal(v4, v) and np.all(m3 == 0)
    big = sc.sample(300000, seed=4)                      # many tiles per CTA: exercises the stage ring
    os.environ["CLIK_TMA"] = "1"
    try:
        _, ctrl_tma = _setup("ur5_track")
        vb, _, _ = _run_device(ctrl_tma, big)
    finally:
        del os.environ["CLIK_TMA"]
    vp, _, _ = _run_device(ctrl, big)
    assert np.array_equal(vb, vp)


def test_damping_option_is_read_at_setup_time():
    """Appendix A19: notebooks mutate ctrl.options after construction (lambda = 1e-26)."""
    sc = scenarios.get("ur5_track")
    ctrl = sc.make_controller()
    ctrl.options["damping_factor"] = 1e-26
    ctrl.setup_solver()
    inp = sc.sample(256, seed=3)
    ref_v, _ = oracle_pinv(sc.spec, inp, {"damping_factor": 1e-26})
    v, _, _ = _run_device(ctrl, inp)
    assert close(v, ref_v, 1e-8, ATOL).all(), _report(v, ref_v, "lambda=1e-26")
    ctrl2 = sc.make_controller()
    ctrl2.options["pinv_method"] = "standard"
    ctrl2.setup_solver()
    ref_s, _ = oracle_pinv(sc.spec, inp, {"pinv_method": "standard"})
    v2, _, _ = _run_device(ctrl2, inp)
    assert close(v2, ref_s, 1e-8, ATOL).all(), _report(v2, ref_s, "standard pinv")


def test_rollout_on_device_matches_stepwise_loop():
    """K closed-loop steps on the device == K calls of solve_batch with the notebook's
    clip + Euler update in between (bit for bit: same kernel arithmetic, same update rounding),
    and the whole trajectory stays within tolerance of the oracle-driven loop."""
    torch = _torch()
    sc, ctrl = _setup("ur5_moe2016_pinv")
    N, K, dt, vmax = 256, 40, 0.008, np.pi / 5
    inp = sc.sample(N, seed=3)
    t0 = torch.from_numpy(inp["t"]).cuda()
    q_dev = torch.from_numpy(inp["q"]).cuda()
    q_roll = q_dev.clone()
    out = ctrl.rollout_batch(t0, q_roll, K, dt, max_speed=vmax)
    q_loop = q_dev.clone()
    q_orc = inp["q"].copy()
    modes_equal = np.ones(N, dtype=bool)
    for k in range(K):
        tk = t0 + dt * k
        v, _, mode = ctrl.solve_batch(tk, q_loop)
        v = torch.clamp(v, -vmax, vmax)
        q_loop = q_loop + v * dt
        vo, mo = oracle_pinv(sc.spec, {"t": inp["t"] + dt * k, "q": q_orc})
        modes_equal &= (mo == mode.cpu().numpy()) | ~modes_equal
        modes_equal &= (mo == mode.cpu().numpy())
        q_orc = q_orc + np.clip(vo, -vmax, vmax) * dt
    torch.cuda.synchronize()
    assert torch.equal(q_roll, q_loop)
    assert torch.equal(out["mode"], mode) and torch.equal(out["robot_vel"], v)
    assert int(out["n_failed"].sum()) == 0
    # instances whose mode sequence agrees with the oracle's must track it closely; a mode flip
    # (an instance sitting on a switching surface to within rounding) is allowed for < 1 %
    assert modes_equal.mean() > 0.99
    err = np.abs(q_roll.cpu().numpy() - q_orc)[:, modes_equal]
    assert err.max() < 1e-9, err.max()


def test_rollout_with_virtual_variable():
    torch = _torch()
    t, p, dp = cs.MX.sym("t"), cs.MX.sym("p"), cs.MX.sym("dp")
    x, dx = cs.MX.sym("x"), cs.MX.sym("dx")
    up = cc.EqualityConstraint("move_up_path_cnstr", 300 - x, gain=1.0, priority=1)
    lim = cc.SetConstraint("cart_limit_cnstr", p, gain=1.0, set_min=0.0, set_max=1.0, priority=1)
    dist = cc.EqualityConstraint("min_dist_cnstr", 0.4 * cs.sin(0.3 * x) - p, gain=1.0, priority=3)
    spec = cc.SkillSpecification("path", t, p, robot_vel_var=dp, virtual_var=x, virtual_vel_var=dx,
                                 constraints=[up, dist, lim])
    ctrl = cc.PseudoInverseController(spec)
    ctrl.setup_solver()
    N, K, dt = 64, 100, 0.02
    p0 = torch.full((1, N), 0.0001, dtype=torch.float64, device="cuda")
    x0 = torch.linspace(0, 5, N, dtype=torch.float64, device="cuda").reshape(1, N).contiguous()
    pl, xl = p0.clone(), x0.clone()
    out = ctrl.rollout_batch(0.0, p0, K, dt, virtual_var=x0, max_speed=0.275, max_virtual_speed=0.5)
    for k in range(K):
        v, xd, _ = ctrl.solve_batch(dt * k, pl, xl)
        pl = pl + torch.clamp(v, -0.275, 0.275) * dt
        xl = xl + torch.clamp(xd, -0.5, 0.5) * dt
    assert torch.equal(p0, pl) and torch.equal(x0, xl)
    assert bool((x0 > xl.new_tensor(0.9)).all())          # the path variable advanced at its speed limit
    assert out["virtual_vel"].shape == (1, N)


def test_kitchen_sink_skill_parity():
    """Everything the constraint classes accept at once: matrix / list / expression gains,
    expression-valued set bounds and velocity targets, time-dependent expressions with and
    without feed-forward, a virtual variable, an input variable, VelocitySet ignored by pinv."""
    t, q, dq = cs.MX.sym("t"), cs.MX.sym("q", 4), cs.MX.sym("dq", 4)
    x, dx, y = cs.MX.sym("x"), cs.MX.sym("dx"), cs.MX.sym("y", 2)
    e1 = cs.vertcat(cs.sin(q[0]) + q[1] * cs.cos(0.3 * t) - y[0], q[2] * q[3] - y[1] + 0.1 * x)
    c1 = cc.EqualityConstraint("mat_gain", e1, gain=np.array([[2.0, 0.3], [0.0, 1.5]]), priority=2)
    c2 = cc.SetConstraint("expr_bounds", q[1] + 0.2 * cs.sin(t), gain=3.0,
                          set_min=cs.MX(-0.4) + 0.0 * y[0] - 0.1 * cs.cos(x), set_max=cs.MX(0.5) + 0.05 * y[1],
                          priority=1)
    c3 = cc.VelocityEqualityConstraint("vel_target", q[0] + 0.5 * q[3], target=0.2 * cs.sin(t) + 0.1 * y[0],
                                       priority=3)
    c4 = cc.EqualityConstraint("list_gain", cs.vertcat(q[2] - 0.3, x - t), gain=[0.7, 1.3], priority=4)
    c5 = cc.SetConstraint("plain", q[3], set_min=-0.2, set_max=0.3, priority=0)
    c6 = cc.VelocitySetConstraint("ignored_by_pinv", q, set_min=-1.0 * np.ones(4), set_max=np.ones(4))
    c7 = cc.EqualityConstraint("expr_gain", q[0] - q[1], gain=cs.MX(1.0) + q[2] * q[2], priority=5)
    spec = cc.SkillSpecification("sink", t, q, robot_vel_var=dq, virtual_var=x, virtual_vel_var=dx,
                                 input_var=y, constraints=[c1, c2, c3, c4, c5, c6, c7])
    assert spec._has_virtual and spec._has_input
    rng = np.random.default_rng(12)
    N = 3000
    inp = {"t": rng.uniform(0, 10, N), "q": rng.uniform(-0.6, 0.6, (4, N)), "x": rng.uniform(-1, 1, (1, N)),
           "y": rng.uniform(-0.5, 0.5, (2, N))}
    # (pinv_method="standard" is unusable here in the reference as well: with lam = 0 the doubled
    # first-equality stack [J; J] has a singular Gram matrix)
    for opts in ({}, {"feedforward": False}, {"damping_factor": 1e-4}):
        ctrl = cc.PseudoInverseController(spec, options=dict(opts))
        ctrl.setup_solver()
        assert ctrl.n_modes == 4
        ref_v, ref_mode = oracle_pinv(spec, inp, dict(opts))
        v, xd, mode = _run_device(ctrl, inp)
        got = np.vstack([v, xd])
        assert len(np.unique(ref_mode)) >= 3
        assert np.array_equal(mode, ref_mode), (opts, int((mode != ref_mode).sum()))
        _assert_parity_multitask(got, spec, inp, opts, str(opts))
