"""The product's expression layer and the fixture's QP minimisers against INDEPENDENT implementations
(tests/golden/independent_pins.json, made by tests/golden/make_independent_pins.py with sympy / scipy and
none of casclik_b200):

  * every catalogue skill rebuilt with sympy as the symbolic engine (its own URDF / DH forward
    kinematics, sympy.diff Jacobians, 40-digit evaluation): constraint values, d e / d [q; x], d e / d t,
    expression-valued gains / bounds / targets must equal what casclik_b200.sym + AD produce (the numbers
    the oracle bridge feeds the oracle and the emitter turns into CUDA);
  * every fixture QP solved by scipy's interior point + a dense KKT polish: same minimiser as the oracle's
    active-set solver that generated the fixture (and as the reference-side values stored in it)."""
import json
import os

import numpy as np
import pytest

from casclik_b200 import cs
from casclik_b200.sym import dag
from oracle_bridge import blocks_from_skill, orc
from test_golden_controllers import VECTORS, load_case, golden_qp, QP

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "independent_pins.json")) as f:
    PINS = json.load(f)


def _pin(name):
    rec = PINS["expressions"][name]
    return PINS["expressions"][rec["same_as"]] if "same_as" in rec else rec


@pytest.mark.parametrize("name", sorted(VECTORS))
def test_expression_layer_matches_the_sympy_rebuild(name):
    spec, inp, kwargs, outputs = load_case(name)
    pin = _pin(name)
    K = pin["n_instances"]
    sub = {k: (v[..., :K] if v is not None else None) for k, v in inp.items()}
    blocks, n = blocks_from_skill(spec, sub["t"], sub["q"], sub.get("x"), sub.get("y"))
    by_label = {c["label"]: c for c in pin["constraints"]}
    assert len(blocks) == len(spec.constraints) == len(by_label)
    worst = 0.0
    for c, b in zip(spec.constraints, blocks):                 # blocks are in the spec's (priority-sorted) order
        p = by_label[c.label]
        for i in range(K):
            row = p["instances"][i]
            e, J, Jt = np.array(row["e"])[:, 0], np.array(row["J"]), np.array(row["Jt"])[:, 0]
            for got, want in ((b.e[i], e), (b.J[i], J), (b.Jt[i], Jt)):
                err = np.abs(got - want).max() / (1.0 + np.abs(want).max())
                worst = max(worst, err)
                assert err < 2e-14, (name, c.label, i, err)
            for fld in ("gain", "set_min", "set_max", "target"):
                if fld in row:
                    got = getattr(b, fld)
                    got = np.asarray(got[i] if isinstance(got, np.ndarray) and got.ndim >= 2 and got.shape[0] == K
                                     else got, dtype=float)
                    want = np.array(row[fld], dtype=float)
                    assert np.abs(got.reshape(-1) - want.reshape(-1)).max() < 2e-14, (name, c.label, fld)
    assert worst < 2e-14


def test_the_pins_cover_the_hard_expressions():
    labels = {c["label"] for n in PINS["expressions"].values() if "constraints" in n for c in n["constraints"]}
    assert {"pose", "mat_gain", "expr_bounds", "expr_gain", "move_point2", "colav_box", "track_point"} <= labels
    iiwa = _pin("pinv/iiwa_multitask")
    pose = [c for c in iiwa["constraints"] if c["label"] == "pose"][0]["instances"][0]
    assert len(pose["e"]) == 4 and len(pose["J"][0]) == 7 and abs(pose["J"][3][0]) > 1e-6    # the norm_fro row
    ks = _pin("pinv/kitchen_sink")
    assert any("set_min" in c["instances"][0] for c in ks["constraints"])
    assert any("gain" in c["instances"][0] for c in ks["constraints"])


@pytest.mark.parametrize("name", QP)
def test_fixture_qp_minimisers_match_a_second_solver(name):
    """reactive_qp.py:493 — the conic solve: the oracle's dual active-set method (which produced the
    fixture's minimisers) against scipy trust-constr + KKT polish on the reference-built matrices."""
    spec, inp, kwargs, outputs = load_case(name)
    gx, gh, gA, glb, gub = golden_qp(outputs)
    rows = PINS["scipy_qp"][name]
    assert len(rows) == len(outputs)
    for i, row in enumerate(rows):
        assert row["success"]
        xs = np.array(row["x"])
        assert np.abs(xs - gx[i]).max() <= 1e-8 * (1 + np.abs(gx[i]).max()), (name, i, np.abs(xs - gx[i]).max())
        obj_s, obj_g = 0.5 * (gh * xs * xs).sum(), 0.5 * (gh * gx[i] ** 2).sum()
        assert abs(obj_s - obj_g) <= 1e-9 * (1 + abs(obj_g))
        xo, lam, st = orc.solve_qp_single(gh, gA[i], glb[i], gub[i])
        assert st == 0 and np.abs(xo - xs).max() <= 1e-8 * (1 + np.abs(xs).max())


@pytest.mark.parametrize("name", sorted(n for n in VECTORS if VECTORS[n]["controller"] == "pinv"))
def test_lowered_program_matches_the_sympy_rebuild(name):
    """What the emitter turns into CUDA — codegen.lower.PinvProgram, whose Jacobian rows may come from the
    kinematic-chain pull-back (sym.dag: reverse mode through registered FK blocks) instead of forward AD —
    evaluated with the NumPy interpreter against the sympy pins."""
    from casclik_b200.codegen import PinvProgram
    from casclik_b200.controllers import PseudoInverseController
    spec, inp, kwargs, outputs = load_case(name)
    opts = PseudoInverseController(spec, **kwargs).options
    prog = PinvProgram(spec, opts)
    pin = _pin(name)
    K = pin["n_instances"]
    by_label = {c["label"]: c for c in pin["constraints"]}
    vals = {prog.syms.t[0].id: inp["t"][:K]}
    for arr, nodes in (("q", prog.syms.q), ("x", prog.syms.x), ("y", prog.syms.y)):
        for k, s in enumerate(nodes):
            vals[s.id] = inp[arr][k, :K]
    for b in prog.blocks:
        p = by_label[b["label"]]
        nodes = [n for r in b["J"] for n in r] + list(b["e"]) + list(b["jt"])
        got = [np.broadcast_to(np.asarray(v, dtype=float), (K,)) for v in dag.evaluate(nodes, vals)]
        m, ns = b["rows"], prog.ns
        for i in range(K):
            row = p["instances"][i]
            J = np.array([[got[r * ns + c][i] for c in range(ns)] for r in range(m)])
            e = np.array([got[m * ns + r][i] for r in range(m)])
            jt = np.array([got[m * ns + m + r][i] for r in range(m)])
            wantJ = np.array(row["J"])
            assert np.abs(J - wantJ).max() <= 1e-13 * (1 + np.abs(wantJ).max()), (name, b["label"], i)
            assert np.abs(e - np.array(row["e"])[:, 0]).max() <= 2e-14 * (1 + np.abs(e).max())
            assert np.abs(jt - np.array(row["Jt"])[:, 0]).max() <= 2e-14 * (1 + np.abs(jt).max())


def test_chain_pullback_is_chosen_for_the_orientation_row_and_shrinks_the_program():
    """iiwa pose task: the ||R_des' R - I||_F row takes the reverse / tip-frame derivation (963 -> 393 eval
    flops); position rows and the UR5 skills keep forward AD (it is the smaller graph there)."""
    from casclik_b200 import scenarios
    from casclik_b200.codegen import PinvProgram
    sc = scenarios.get("iiwa_multitask")
    ctrl = sc.make_controller()
    auto = PinvProgram(sc.spec, ctrl.options)
    with dag.ad_mode("forward"):
        fwd = PinvProgram(sc.spec, ctrl.options)
    ja = [n for b in auto.blocks for r in b["J"] for n in r]
    jf = [n for b in fwd.blocks for r in b["J"] for n in r]
    assert dag.graph_cost(ja) < 0.6 * dag.graph_cost(jf)
    pose_a, pose_f = auto.blocks[-1]["J"], fwd.blocks[-1]["J"]
    assert all(a is f for a, f in zip(pose_a[0], pose_f[0]))            # position row: same (forward) nodes
    assert any(a is not f for a, f in zip(pose_a[3], pose_f[3]))        # orientation row: different derivation
    inp = sc.sample(50, seed=2)
    vals = {auto.syms.t[0].id: inp["t"]}
    for arr, nodes in (("q", auto.syms.q), ("y", auto.syms.y)):
        for k, s in enumerate(nodes):
            vals[s.id] = inp[arr][k]
    va = np.array([np.broadcast_to(v, (50,)) for v in dag.evaluate(ja, vals)])
    vf = np.array([np.broadcast_to(v, (50,)) for v in dag.evaluate(jf, vals)])
    assert np.abs(va - vf).max() < 1e-13
    u = scenarios.get("ur5_track")
    cu = u.make_controller()
    with dag.ad_mode("forward"):
        uf = PinvProgram(u.spec, cu.options)
    ua = PinvProgram(u.spec, cu.options)
    assert all(a is f for ra, rf in zip(ua.blocks[0]["J"], uf.blocks[0]["J"]) for a, f in zip(ra, rf))
