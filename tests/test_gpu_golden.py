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
