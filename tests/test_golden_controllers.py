"""The oracle against tests/golden/controller_vectors.json: outputs of the REFERENCE's own
PseudoInverseController / ReactiveQPController code (pseudo_inverse.py:259-556, reactive_qp.py:175-540),
executed unmodified over a stand-in CasADi module by tests/golden/make_controller_vectors.py.
The skills are rebuilt here with casclik_b200's classes from the same catalogue (golden_skills.py)
and evaluated on the fixture's inputs."""
import json
import os

import numpy as np
import pytest

import casclik_b200 as cc
from casclik_b200 import cs
import golden_skills as gs
from oracle_bridge import oracle_pinv, oracle_qp_problem, orc, close

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "controller_vectors.json")
with open(GOLDEN) as f:
    VECTORS = json.load(f)
NS = gs.Namespace(cs, cc)
PINV = sorted(n for n, r in VECTORS.items() if r["controller"] == "pinv")
QP = sorted(n for n, r in VECTORS.items() if r["controller"] == "qp")


def load_case(name):
    rec = VECTORS[name]
    inp = {k: np.array(v, dtype=np.float64) for k, v in rec["inputs"].items()}
    spec, inp, kind, kwargs = gs.build(NS, name, inputs=inp)
    assert kind == rec["controller"] and json.loads(json.dumps(kwargs)) == rec["kwargs"]
    return spec, inp, kwargs, rec["outputs"]


def golden_velocities(outputs):
    v = np.array([o["robot_vel"] + (o["virtual_vel"] or []) for o in outputs]).T
    return v, np.array([o["mode"] for o in outputs])


def golden_qp(outputs):
    x = np.array([o["robot_vel"] + (o["virtual_vel"] or []) + (o["slack"] or []) for o in outputs])
    return (x, np.array(outputs[0]["h"]), np.array([o["A"] for o in outputs]),
            np.array([o["lb"] for o in outputs]), np.array([o["ub"] for o in outputs]))


def qp_weights(kwargs):
    w = {}
    if "robot_var_weights" in kwargs:
        w["w_rob"] = kwargs["robot_var_weights"]
    if "virtual_var_weights" in kwargs:
        w["w_virt"] = kwargs["virtual_var_weights"]
    return w


def test_fixture_covers_the_catalogue():
    assert sorted(VECTORS) == sorted(gs.CASES)
    modes = set()
    for n in PINV:
        modes |= {(n, o["mode"]) for o in VECTORS[n]["outputs"]}
    assert len({m for n, m in modes if n == "pinv/iiwa_multitask"}) >= 12
    assert {m for n, m in modes if n == "pinv/conv_last"} >= {0, 1, 2}
    assert {m for n, m in modes if n == "pinv/cart_kat_far_target"} == {0, 1}


def test_hand_derived_known_answers_are_what_the_reference_code_returns():
    """SURVEY §8c derived P1-P3 / Q1-Q2 by hand from the reference's formulas; the fixture holds what
    the reference's code itself returns at the same points (cart skills, first instances)."""
    lam = 1e-7
    kat = VECTORS["pinv/cart_kat"]
    assert kat["inputs"]["q"][0][:2] == [0.25, 1.2]
    v0, v1 = (kat["outputs"][i]["robot_vel"][0] for i in (0, 1))
    assert kat["outputs"][0]["mode"] == 0 and abs(v0 - 0.5 * (1 + 2 * lam) / (1 + lam) ** 2) < 1e-15       # P1
    assert kat["outputs"][1]["mode"] == 0 and abs(v1 + 0.45 * (1 + 2 * lam) / (1 + lam) ** 2) < 1e-15      # P2
    far = VECTORS["pinv/cart_kat_far_target"]
    assert far["inputs"]["q"][0][1] == 1.2 and far["outputs"][1]["mode"] == 1                              # P3
    assert abs(far["outputs"][1]["robot_vel"][0] - lam / (1 + lam) * 0.3 / (1 + lam)) < 1e-15
    qp = VECTORS["qp/cart_kat"]
    assert qp["inputs"]["q"][0][2] == 0.0 and qp["inputs"]["q"][0][4] == 0.6
    q1, q2 = qp["outputs"][2], qp["outputs"][4]
    assert abs(q1["robot_vel"][0] - 0.275) < 1e-12 and abs(q1["slack"][0] - 0.475) < 1e-12                 # Q1
    assert abs(q2["robot_vel"][0] - 0.14985029940119762) < 1e-12                                           # Q2
    assert abs(q2["slack"][0] - 1.4970059880239917e-4) < 1e-12


@pytest.mark.parametrize("name", PINV)
def test_oracle_pinv_reproduces_the_reference_outputs(name):
    spec, inp, kwargs, outputs = load_case(name)
    v, mode = oracle_pinv(spec, inp, dict(kwargs.get("options", {})))
    gv, gmode = golden_velocities(outputs)
    assert np.array_equal(mode, gmode)
    assert close(v, gv, 1e-9, 1e-12).all(), np.abs(v - gv).max()


@pytest.mark.parametrize("name", QP)
def test_oracle_qp_matrices_and_solution_reproduce_the_reference(name):
    spec, inp, kwargs, outputs = load_case(name)
    h, A, lb, ub = oracle_qp_problem(spec, inp, **qp_weights(kwargs))
    gx, gh, gA, glb, gub = golden_qp(outputs)
    assert np.abs(h - gh).max() <= 1e-15
    assert np.abs(A - gA).max() <= 1e-13
    assert close(lb, glb, 1e-13, 1e-13).all() and close(ub, gub, 1e-13, 1e-13).all()
    for i in range(len(outputs)):
        x, lam, status = orc.solve_qp_single(h, A[i], lb[i], ub[i])
        assert status == 0 and np.abs(x - gx[i]).max() <= 1e-9 * (1 + np.abs(gx[i]).max())


def initial_args(spec, inp, rec, i):
    ini = rec["initial"][i]
    col = lambda k: None if k not in inp else inp[k][:, i]
    return ini, (float(inp["t"][i]), inp["q"][:, i], col("x"), np.array(ini["dq0"]), col("y"))


@pytest.mark.parametrize("name", QP)
def test_initial_value_problem_matrices_reproduce_the_reference(name):
    """reactive_qp.py:297-424: H, A, lb, ub of the slack / virtual-variable initial problem (the
    matrices are host-side expression evaluation; the solve itself is checked on the GPU)."""
    spec, inp, kwargs, outputs = load_case(name)
    ctrl = cc.ReactiveQPController(spec, **kwargs)
    ctrl.setup_initial_problem_solver()
    assert ctrl._has_initial
    ip = ctrl._initial
    for i in range(len(VECTORS[name]["initial"])):
        ini, (t, q, x, dq, y) = initial_args(spec, inp, VECTORS[name], i)
        vals = [t, q, dq] + ([x] if spec._has_virtual else []) + ([y] if spec._has_input else [])
        H, A, lb, ub = (np.asarray(ip.funcs[k](*vals).toarray()) for k in ("H", "A", "Blb", "Bub"))
        assert np.abs(np.diag(H) - np.array(ini["h"])).max() <= 1e-15 and np.count_nonzero(H - np.diag(np.diag(H))) == 0
        assert np.abs(A - np.array(ini["A"])).max() <= 1e-13
        assert close(lb.reshape(-1), np.array(ini["lb"]), 1e-13, 1e-13).all()
        assert close(ub.reshape(-1), np.array(ini["ub"]), 1e-13, 1e-13).all()
