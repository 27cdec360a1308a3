"""Test glue: evaluate a skill's constraint expressions with the NumPy DAG interpreter and hand
the numbers to the oracle (oracle/clik_oracle.py).  The CUDA code generator and kernels are not
involved: this path shares only the expression graph (AD) with the product, and tests/test_fk.py
checks that graph against the oracle's independent geometric FK/Jacobian."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clik_oracle as orc  # noqa: E402

from casclik_b200 import cs  # noqa: E402
from casclik_b200.sym import dag  # noqa: E402
from casclik_b200.codegen.lower import kind_of  # noqa: E402

KINDS = {0: orc.EQ, 1: orc.SET, 2: orc.VELEQ, 3: orc.VELSET}


def _values(spec, t, q, x, y):
    vals = {spec.time_var.nodes()[0].id: np.asarray(t, dtype=np.float64)}
    for i, s in enumerate(spec.robot_var.nodes()):
        vals[s.id] = q[i]
    if spec.virtual_var is not None:
        for i, s in enumerate(spec.virtual_var.nodes()):
            vals[s.id] = x[i] if x is not None else np.zeros_like(q[0])
    if spec.input_var is not None and y is not None:
        for i, s in enumerate(spec.input_var.nodes()):
            vals[s.id] = y[i]
    return vals


def _num(matrix, vals, N):
    """cs matrix -> (N, rows, cols) float array."""
    m = matrix if isinstance(matrix, cs.GenericMatrixCommon) else cs.DM(matrix)
    r, c = m.shape
    out = dag.evaluate(m.nodes(), vals) if m.numel() else []
    arr = np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), (N,)) for v in out], axis=1) \
        if out else np.zeros((N, 0))
    return arr.reshape(N, c, r).transpose(0, 2, 1)


def blocks_from_skill(spec, t, q, x=None, y=None):
    """-> (list of oracle Blocks in priority order, n_state).  Jacobians are plain forward-mode AD
    (dag.ad_mode("forward")): the oracle never takes the kinematic-chain pull-back the product's lowering
    may choose, so kernel-vs-oracle parity also cross-checks the two derivations."""
    with dag.ad_mode("forward"):
        return _blocks_from_skill(spec, t, q, x, y)


def _blocks_from_skill(spec, t, q, x=None, y=None):
    N = q.shape[1]
    t = np.broadcast_to(np.asarray(t, dtype=np.float64).reshape(-1), (N,)) if np.ndim(t) else \
        np.full((N,), float(t))
    vals = _values(spec, t, q, x, y)
    state = spec.robot_var if spec.virtual_var is None else cs.vertcat(spec.robot_var, spec.virtual_var)
    blocks = []
    for c in spec.constraints:
        e = c.expression
        rows = e.size()[0]
        kind = KINDS[kind_of(c)]
        E = _num(e, vals, N)[:, :, 0]
        J = _num(cs.jacobian(e, state), vals, N)
        Jt = _num(cs.jacobian(e, spec.time_var), vals, N)[:, :, 0]
        g = c.gain
        if isinstance(g, list):
            g = np.diag(np.asarray(g, dtype=float))
        elif isinstance(g, cs.GenericMatrixCommon):
            g = _num(g, vals, N)
            g = g[:, 0, 0] if g.shape[1:] == (1, 1) and False else g
            if g.shape[1:] == (1, 1):
                g = g[0, 0, 0] if np.all(g == g[0]) else g
        kw = {}
        if hasattr(c, "set_min"):
            kw["set_min"] = _num(c.set_min, vals, N)[:, :, 0] if not np.isscalar(c.set_min) else float(c.set_min)
            kw["set_max"] = _num(c.set_max, vals, N)[:, :, 0] if not np.isscalar(c.set_max) else float(c.set_max)
        if hasattr(c, "target"):
            kw["target"] = _num(c.target, vals, N)[:, :, 0] if not np.isscalar(c.target) else float(c.target)
        blocks.append(orc.Block(kind, E, J, Jt, g, soft=(c.constraint_type == "soft"),
                                slack_weight=float(c.slack_weight), **kw))
    n_state = spec.n_robot_var + (spec.n_virtual_var if spec.virtual_var is not None else 0)
    return blocks, n_state


def oracle_pinv(spec, inputs, options=None, dtype=None):
    """dtype=np.longdouble evaluates the same literal formulas on the same float64 constraint
    values (e, J, Jt) in extended precision: the referee for ill-conditioned skills."""
    blocks, n = blocks_from_skill(spec, inputs["t"], inputs["q"], inputs.get("x"), inputs.get("y"))
    if dtype is not None:
        blocks = orc.as_dtype(blocks, dtype)
    v, mode = orc.pinv_step(blocks, n, options)
    return v.T.copy(), mode


def oracle_qp_problem(spec, inputs, **weights):
    blocks, n = blocks_from_skill(spec, inputs["t"], inputs["q"], inputs.get("x"), inputs.get("y"))
    n_virt = spec.n_virtual_var if spec.virtual_var is not None else 0
    return orc.qp_matrices(blocks, spec.n_robot_var, n_virt, **weights)


def close(a, b, rtol=1e-9, atol=1e-12):
    """|a - b| <= atol + rtol*|b| element-wise (the north-star tolerance for the pinv path)."""
    return np.abs(a - b) <= atol + rtol * np.abs(b)
