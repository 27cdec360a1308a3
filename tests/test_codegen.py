"""Expression compiler: the emitted straight-line text computes what the graph says (checked by
compiling the C flavour with gcc and comparing with the NumPy interpreter), every BASELINE
scenario compiles to an sm_100a cubin with nvcc, and the hot kernel stays in registers."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from casclik_b200 import cs, scenarios, build
from casclik_b200.codegen import PinvProgram, QpProgram, emit_skill, emit_c_function
from casclik_b200.sym import dag


def test_emitted_c_matches_numpy_interpreter():
    sc = scenarios.get("iiwa_multitask")
    ctrl = sc.make_controller()
    prog = PinvProgram(sc.spec, ctrl.options)
    outs = []
    for b in prog.blocks:
        outs += b["e"] + b["jt"] + [n for r in b["J"] for n in r] + b.get("des", [])
    names = dict(prog.syms.names)
    src = emit_c_function("ev", names, outs, "double t, const double* q, const double* x, const double* y")
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "ev.c")
        so = os.path.join(tmp, "ev.so")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", so, c, "-lm"], check=True)
        lib = ctypes.CDLL(so)
        inp = sc.sample(16, seed=3)
        ids_q = [s.id for s in prog.syms.q]
        ids_y = [s.id for s in prog.syms.y]
        for k in range(16):
            q = np.ascontiguousarray(inp["q"][:, k])
            y = np.ascontiguousarray(inp["y"][:, k])
            out = np.zeros(len(outs))
            lib.ev(ctypes.c_double(0.0), q.ctypes.data_as(ctypes.c_void_p), None,
                   y.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
            vals = {prog.syms.t[0].id: 0.0}
            vals.update({i: q[j] for j, i in enumerate(ids_q)})
            vals.update({i: y[j] for j, i in enumerate(ids_y)})
            ref = np.array([float(v) for v in dag.evaluate(outs, vals)])
            assert np.abs(out - ref).max() <= 1e-13 * (1 + np.abs(ref).max())


def test_skill_source_structure_and_meta():
    sc = scenarios.get("ur5_moe2016_pinv")
    ctrl = sc.make_controller()
    prog = PinvProgram(sc.spec, ctrl.options)
    assert [b["label"] for b in prog.blocks] == ["colav_y", "colav_x", "colav_z", "move_point2"]
    assert prog.n_sets == 3 and prog.m == 6 and prog.ns == 6 and prog.n_in == 0
    src, meta = emit_skill(pinv=prog, label="box_move")
    assert meta["n_modes"] == 8 and meta["has_pinv"] and not meta["has_qp"]
    assert "clik_mode_tab[8] = {0, 1, 2, 4, 3, 5, 6, 7}" in src     # popcount order, set 0 = LSB
    assert "NSETS = 3" in src and 'extern "C" __global__' in src and "clik::sincos_fast" in src
    assert meta["pinv_eval"]["transcendentals"]["sin"] >= 5
    assert meta["pinv_flops_mode0"] > meta["pinv_eval"]["flops"] > 0
    # the time-varying path constraint reads t, every joint matters for the box sets
    assert meta["pinv_inputs_read"] == 6      # t + q1..q5 (the flange position does not depend on q6)


def test_qp_program_matches_reference_layout():
    sc = scenarios.get("ur5_qp")
    ctrl = sc.make_controller()
    prog = QpProgram(sc.spec, ctrl.robot_var_weights, ctrl.virtual_var_weights,
                     ctrl.slack_var_weights, ctrl.weight_shifter)
    assert prog.nx == 9 and prog.m == 15
    hv = [n.val for n in prog.h]
    assert np.allclose(hv, [0.001] * 6 + [1.001] * 3)
    A = prog.A
    for r in range(3):                                   # soft rows carry -I on their slack
        assert [n.val if n.is_const else None for n in A[r][6:]] == [-1.0 if c == r else 0.0 for c in range(3)]
    for r in range(6):                                   # joint limits and speed limits: identity rows
        for blk in (3, 9):
            assert [n.val for n in A[blk + r]] == [1.0 if c == r else 0.0 for c in range(9)]
    assert all(n.is_const and abs(n.val - math_pi_5()) < 1e-15 for n in prog.ub[9:])


def math_pi_5():
    import math
    return math.pi / 5


@pytest.mark.parametrize("name", sorted(scenarios.REGISTRY))
def test_every_scenario_compiles_for_sm_100a(name):
    ctrl = scenarios.get(name).make_controller()
    ctrl.setup_problem_functions(load=False)
    assert os.path.exists(ctrl.cubin_path) and os.path.getsize(ctrl.cubin_path) > 1000
    log = open(ctrl.cubin_path.replace(".cubin", ".log")).read()
    assert "sm_100a" in log
    kernel = "clik_qp_kernel" if name.endswith("qp") else "clik_pinv_kernel"
    assert ("Compiling entry function '%s'" % kernel) in log


def test_headline_kernel_stays_in_registers():
    ctrl = scenarios.get("ur5_track").make_controller()
    ctrl.setup_problem_functions(load=False)
    log = open(ctrl.cubin_path.replace(".cubin", ".log")).read()
    sec = log[log.index("Compiling entry function 'clik_pinv_kernel'"):]
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", sec)
    assert m and m.group(1) == "0" and m.group(2) == "0"
    regs = int(re.search(r"Used (\d+) registers", sec).group(1))
    assert regs <= 96
    assert ctrl.kernel_meta["pinv_bytes_per_step"] == 8 * 8 + 8 * 6 + 4


def test_cubin_cache_is_keyed_on_source_and_headers():
    a = build.source_hash("x")
    assert a == build.source_hash("x") and a != build.source_hash("y")
