"""The CUDA path (through the C ABI) against the reference-generated fixture
tests/golden/controller_vectors.json — see test_golden_controllers.py for how it was made."""
import numpy as np
import pytest

import casclik_b200 as cc
from oracle_bridge import close, orc
from test_golden_controllers import (PINV, QP, load_case, golden_velocities, golden_qp)

pytestmark = pytest.mark.gpu


def _device(inp):
    import torch
    put = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return put(inp["t"]), put(inp["q"]), put(inp.get("x")), put(inp.get("y"))


@pytest.mark.parametrize("name", PINV)
def test_pinv_kernel_reproduces_the_reference_outputs(name):
    spec, inp, kwargs, outputs = load_case(name)
    ctrl = cc.PseudoInverseController(spec, **kwargs)
    ctrl.setup_solver()
    t, q, x, y = _device(inp)
    v, xd, mode = ctrl.solve_batch(t, q, x, y)
    got = v.cpu().numpy() if xd is None else np.vstack([v.cpu().numpy(), xd.cpu().numpy()])
    gv, gmode = golden_velocities(outputs)
    assert np.array_equal(mode.cpu().numpy(), gmode)
    ok = close(got, gv, 1e-9, 1e-12)
    if "kitchen_sink" in name or "conv_last" in name:
        # multi-task chains: cond(J J' + lam I) ~ 1e7 — norm-wise bound per instance
        err = np.linalg.norm(got - gv, axis=0) / np.maximum(np.linalg.norm(gv, axis=0), 1e-300)
        assert err.max() < 1e-9 and ok.mean() > 0.98, (err.max(), ok.mean())
    else:
        assert ok.all(), np.abs(got - gv).max()
    # and the single-instance call of the reference API, on the first instance
    r = ctrl.solve(float(inp["t"][0]), inp["q"][:, 0], None if "x" not in inp else inp["x"][:, 0],
                   None if "y" not in inp else inp["y"][:, 0])
    nq = inp["q"].shape[0]
    assert np.array_equal(np.asarray(r[0]).reshape(-1), got[:nq, 0]) and ctrl.current_mode == gmode[0]


@pytest.mark.parametrize("name", QP)
def test_qp_kernel_reproduces_the_reference_problem_and_its_minimiser(name):
    spec, inp, kwargs, outputs = load_case(name)
    ctrl = cc.ReactiveQPController(spec, **kwargs)
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    t, q, x, y = _device(inp)
    sol, status, _ = ctrl.solve_batch(t, q, x, y)
    sol = sol.cpu().numpy().T
    assert int(status.abs().sum()) == 0
    gx, gh, gA, glb, gub = golden_qp(outputs)
    for i in range(len(outputs)):
        assert np.abs(sol[i] - gx[i]).max() <= 1e-7 * (1 + np.abs(gx[i]).max())
        obj, gobj = 0.5 * (gh * sol[i] ** 2).sum(), 0.5 * (gh * gx[i] ** 2).sum()
        assert abs(obj - gobj) <= 1e-6 * (1 + abs(gobj))
        kk = orc.kkt_residuals(gh, gA[i], glb[i], gub[i], sol[i])      # against the REFERENCE-built matrices
        assert kk["primal"] < 1e-6 and kk["stationarity"] < 1e-6 and kk["sign"] < 1e-6


@pytest.mark.parametrize("name", QP)
def test_initial_value_problem_reproduces_the_reference(name):
    from test_golden_controllers import VECTORS, initial_args
    spec, inp, kwargs, outputs = load_case(name)
    ctrl = cc.ReactiveQPController(spec, **kwargs)
    ctrl.setup_initial_problem_solver()
    for i in range(len(VECTORS[name]["initial"])):
        ini, (t, q, x, dq, y) = initial_args(spec, inp, VECTORS[name], i)
        virt, slack = ctrl.solve_initial_problem(t, q, x, dq, y)
        for got, want in ((virt, ini["virtual"]), (slack, ini["slack"])):
            if want is None:
                assert got is None
            else:
                g, w = np.asarray(got.toarray()).reshape(-1), np.array(want)
                assert np.abs(g - w).max() <= 1e-7 * (1 + np.abs(w).max())


@pytest.mark.parametrize("name", QP)
def test_batched_initial_value_problem_reproduces_the_reference(name):
    """solve_initial_problem_batch: the initial QP (reactive_qp.py:300-459) lowered through the expression
    compiler into its own fused kernel, all fixture instances in one launch (device and host ABI), against
    the slack / virtual velocities the reference's solve_initial_problem produced one at a time."""
    import torch
    from test_golden_controllers import VECTORS, initial_args
    spec, inp, kwargs, outputs = load_case(name)
    ctrl = cc.ReactiveQPController(spec, **kwargs)
    ctrl.setup_initial_problem_solver()
    recs = VECTORS[name]["initial"]
    K = len(recs)
    args = [initial_args(spec, inp, VECTORS[name], i)[1] for i in range(K)]
    t = np.array([a[0] for a in args])
    q = np.stack([a[1] for a in args], axis=1)
    x = None if args[0][2] is None else np.stack([a[2] for a in args], axis=1)
    dq = np.stack([a[3] for a in args], axis=1)
    y = None if args[0][4] is None else np.stack([a[4] for a in args], axis=1)
    up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    virt, slack, status = ctrl.solve_initial_problem_batch(up(t), up(q), up(x), up(dq), up(y))
    torch.cuda.synchronize()
    assert bool((status == 0).all())
    hv, hs, hst = ctrl.solve_initial_problem_batch(t, q, x, dq, y)          # host ABI, same bits
    for i, ini in enumerate(recs):
        for got, hgot, want in ((virt, hv, ini["virtual"]), (slack, hs, ini["slack"])):
            if want is None:
                assert got is None
            else:
                g, w = got[:, i].cpu().numpy(), np.array(want)
                assert np.abs(g - w).max() <= 1e-7 * (1 + np.abs(w).max()), (name, i)
                assert np.array_equal(np.asarray(hgot)[:, i], g)
    # zero robot velocity is the default, as in the reference (:441-442)
    v0, s0, _ = ctrl.solve_initial_problem_batch(up(t), up(q), up(x), None, up(y))
    v1, s1, _ = ctrl.solve_initial_problem_batch(up(t), up(q), up(x), up(np.zeros_like(dq)), up(y))
    assert (s0 is None or torch.equal(s0, s1)) and (v0 is None or torch.equal(v0, v1))


@pytest.mark.parametrize("name", sorted(n for n in PINV + QP if "rollout" in
                                        __import__("test_golden_controllers").VECTORS[n]))
def test_device_rollout_reproduces_the_reference_simulation_loop(name):
    """The notebooks' loop around the reference's solve() (q += clip(v)*dt for `steps` steps, run by
    the generator with the reference controllers) against ONE clik_*_rollout launch."""
    import torch
    from test_golden_controllers import VECTORS
    spec, inp, kwargs, outputs = load_case(name)
    ro = VECTORS[name]["rollout"]
    cfg, n = ro["config"], ro["config"]["n"]
    qp = VECTORS[name]["controller"] == "qp"
    ctrl = (cc.ReactiveQPController if qp else cc.PseudoInverseController)(spec, **kwargs)
    if qp:
        ctrl.setup_problem_functions()
    ctrl.setup_solver()
    put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    q = put(inp["q"][:, :n])
    x = put(inp["x"][:, :n]) if "x" in inp else None
    y = put(inp["y"][:, :n]) if "y" in inp else None
    out = ctrl.rollout_batch(put(inp["t"][:n]), q, cfg["steps"], cfg["dt"], virtual_var=x, input_var=y,
                             max_speed=cfg["max_speed"], max_virtual_speed=cfg["max_virtual_speed"])
    assert int(out["n_failed"].sum()) == 0
    want = np.array(ro["q_final"]).T
    got = q.cpu().numpy()
    # a closed loop contracts errors along the task and integrates them elsewhere: 1e-7 over 80-300 steps
    assert np.abs(got - want).max() <= 1e-7 * (1 + np.abs(want).max()), np.abs(got - want).max()
    assert np.abs(got - inp["q"][:, :n]).max() > 1e-2          # the state did move
    if x is not None:
        assert np.abs(x.cpu().numpy() - np.array(ro["x_final"]).T).max() <= 1e-7
    if not qp:
        assert out["mode"].cpu().numpy().tolist() == ro["last_mode"]
