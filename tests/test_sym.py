"""Expression layer: AD against finite differences, simplification, CasADi-surface semantics."""
import numpy as np
import pytest

from casclik_b200 import cs
from casclik_b200.sym import dag


def _fd(f, x0, h=1e-6):
    cols = []
    for i in range(len(x0)):
        e = np.zeros(len(x0))
        e[i] = h
        cols.append((f(x0 + e) - f(x0 - e)) / (2 * h))
    return np.stack(cols, axis=1)


def test_jacobian_matches_finite_differences_on_a_rich_expression():
    q = cs.MX.sym("q", 4)
    t = cs.MX.sym("t")
    e = cs.vertcat(cs.sin(q[0]) * cs.cos(q[1] * t) + q[2] ** 2 / (1.0 + q[3] * q[3]),
                   cs.sqrt(1.0 + q[0] * q[0]) * cs.exp(-q[1]) + cs.atan2(q[2], 2.0 + q[3]),
                   cs.norm_2(cs.vertcat(q[0], q[1], q[2])) * cs.fabs(q[3]) + cs.tan(0.1 * q[0]),
                   cs.if_else(q[0] > 0.3, q[1] * q[2], q[3] ** 3, True) + cs.log(2.0 + q[1] ** 2),
                   cs.fmax(q[0], q[1]) + cs.fmin(q[2], q[3]) + cs.asin(0.1 * q[0]) + cs.acos(0.1 * q[1]))
    f = cs.Function("f", [t, q], [e, cs.jacobian(e, q), cs.jacobian(e, t), cs.jtimes(e, q, cs.DM([1., -2., 0.5, 3.]))])
    rng = np.random.default_rng(0)
    for _ in range(5):
        x0 = rng.uniform(0.4, 1.5, 4)
        val, J, Jt, jv = (r.toarray() for r in f(0.7, x0))
        Jfd = _fd(lambda x: f(0.7, x)[0].toarray()[:, 0], x0)
        assert np.abs(J - Jfd).max() < 1e-7
        Jtfd = (f(0.7 + 1e-6, x0)[0].toarray() - f(0.7 - 1e-6, x0)[0].toarray()) / 2e-6
        assert np.abs(Jt - Jtfd).max() < 1e-7
        assert np.abs(jv[:, 0] - J @ np.array([1., -2., 0.5, 3.])).max() < 1e-12


def test_hash_consing_and_simplification():
    x, y = dag.symbol("x"), dag.symbol("y")
    assert dag.add(x, y) is dag.add(y, x)
    assert dag.mul(x, dag.ONE) is x and dag.mul(x, dag.ZERO) is dag.ZERO
    assert dag.add(x, dag.ZERO) is x and dag.sub(x, x) is dag.ZERO
    assert dag.neg(dag.neg(x)) is x
    assert dag.mul(dag.neg(x), dag.neg(y)) is dag.mul(x, y)
    assert dag.sin(dag.neg(x)) is dag.neg(dag.sin(x)) and dag.cos(dag.neg(x)) is dag.cos(x)
    assert dag.const(2.0) is dag.const(2.0) and dag.const(-0.0) is dag.ZERO
    assert dag.pow_(x, dag.const(2.0)) is dag.mul(x, x)
    assert dag.if_else(dag.ONE, x, y) is x and dag.if_else(dag.lt(x, y), x, x) is x
    # structural zeros survive differentiation: d(x*y)/dz == 0 exactly
    z = dag.symbol("z")
    assert dag.jacobian([dag.mul(x, y)], [z])[0][0] is dag.ZERO


def test_structural_nnz_drives_has_virtual_detection():
    q, v = cs.MX.sym("q", 2), cs.MX.sym("v", 2)
    e = cs.vertcat(q[0] * q[1], cs.sin(q[0]))
    assert cs.jacobian(e, v).nnz() == 0
    assert cs.jacobian(e, q).nnz() == 3
    assert cs.jacobian(e + v[1], v).nnz() == 2


def test_casadi_surface_shapes_and_indexing():
    T = cs.MX.sym("T", 4, 4)
    assert T.size() == (4, 4) and T.shape == (4, 4) and T.size1() == 4 and T.size2() == 4
    assert T[:3, 3].shape == (3, 1) and T[0, :3].shape == (1, 3) and T[:3, :3].shape == (3, 3)
    assert T[0, :3].T.shape == (3, 1)
    q = cs.MX.sym("q", 6)
    assert q[2].shape == (1, 1) and q[1:4].shape == (3, 1) and q[-1].nodes()[0] is q.nodes()[5]
    assert T[5].nodes()[0] is T[1, 1].nodes()[0]           # linear indexing is column-major
    assert cs.vertcat([1.] * 3).shape == (3, 1)             # single list argument
    assert cs.vertcat(q, q[:2]).shape == (8, 1) and cs.horzcat(q, q).shape == (6, 2)
    assert cs.MX.zeros(3).shape == (3, 1) and cs.DM.zeros((2, 5)).shape == (2, 5)
    assert cs.MX.eye(3).nnz() == 3
    assert cs.diag(cs.vertcat(1., 2.)).shape == (2, 2) and cs.diag(cs.MX.eye(3)).shape == (3, 1)
    assert cs.reshape(q, 3, 2).T.shape == (2, 3)
    assert q.is_symbolic() and not (q + 1).is_symbolic() and not cs.MX(cs.DM([1., 2.])).is_symbolic()
    assert isinstance(q, cs.GenericMatrixCommon) and isinstance(cs.DM(1.0), cs.GenericMatrixCommon)


def test_mixed_arithmetic_and_numpy_interop():
    q = cs.MX.sym("q", 3)
    a = np.array([1., 2., 3.])
    r = a - q                       # ndarray on the left must defer to the matrix type
    assert isinstance(r, cs.MX) and r.shape == (3, 1)
    r2 = 2.0 * q + q * a - q / 2
    f = cs.Function("f", [q], [r, r2, cs.mtimes(q.T, q), cs.mtimes(np.eye(3) * 2, q), cs.dot(q, a)])
    out = [o.toarray() for o in f([0.5, -1.0, 2.0])]
    x = np.array([0.5, -1.0, 2.0])
    assert np.allclose(out[0][:, 0], a - x) and np.allclose(out[1][:, 0], 2 * x + x * a - x / 2)
    assert np.allclose(out[2], x @ x) and np.allclose(out[3][:, 0], 2 * x) and np.allclose(out[4], x @ a)
    d = -cs.vertcat([0.5] * 2)
    assert isinstance(d, cs.DM) and np.allclose(d.toarray()[:, 0], [-0.5, -0.5])
    assert float(cs.DM(3.5)) == 3.5 and int(cs.DM(1.0)) == 1 and bool(cs.DM(1.0)) and not cs.DM(0.0)
    with pytest.raises(TypeError):
        bool(q[0] > 0)


def test_function_calling_conventions():
    t, q = cs.MX.sym("t"), cs.MX.sym("q", 2)
    f = cs.Function("f", [t, q], [t * q, cs.sumsqr(q)], ["t", "q"], ["a", "b"], {"jit": True})
    a, b = f(2.0, [1.0, 3.0])
    assert isinstance(a, cs.DM) and np.allclose(a.toarray()[:, 0], [2.0, 6.0]) and float(b) == 10.0
    a2, _ = f(2.0, np.array([1.0, 3.0]))
    assert np.array_equal(a2.toarray(), a.toarray())
    a3, _ = f(cs.DM(2.0), cs.DM([1.0, 3.0]))
    assert np.array_equal(a3.toarray(), a.toarray())
    r = f(t=2.0, q=[1.0, 3.0])
    assert set(r) == {"a", "b"}
    # symbolic call substitutes (how the notebooks use T_fk(q))
    z = cs.MX.sym("z", 2)
    sa, sb = f(t, 2 * z)
    assert isinstance(sa, cs.MX) and cs.jacobian(sb, z).nnz() == 2
    g = cs.Function("g", [q], [cs.sin(q[0])])
    assert isinstance(g(q), cs.MX) and g(q).shape == (1, 1)
    with pytest.raises(RuntimeError):
        cs.Function("bad", [t], [t * q[0]])            # free variable
    with pytest.raises(TypeError):
        f(1.0)


def test_small_linear_algebra():
    A = np.array([[4., 1., 0.], [1., 3., 1.], [0., 1., 2.]])
    assert np.allclose(cs.inv(A).toarray(), np.linalg.inv(A))
    assert np.allclose(cs.solve(A, np.array([1., 2., 3.])).toarray()[:, 0], np.linalg.solve(A, [1., 2., 3.]))
    assert abs(float(cs.det(A)) - np.linalg.det(A)) < 1e-12
    q = cs.MX.sym("q", 2)
    M = cs.vertcat(cs.horzcat(2 + q[0] ** 2, q[1]), cs.horzcat(q[1], 3.0))
    f = cs.Function("f", [q], [cs.inv(M), cs.pinv(cs.horzcat(M, q))])
    Mi, P = (o.toarray() for o in f([0.3, -0.7]))
    Mn = np.array([[2 + 0.09, -0.7], [-0.7, 3.0]])
    assert np.allclose(Mi, np.linalg.inv(Mn))
    assert np.allclose(P, np.linalg.pinv(np.hstack([Mn, [[0.3], [-0.7]]])))
    assert np.allclose(cs.cross([1., 0, 0], [0, 1., 0]).toarray()[:, 0], [0, 0, 1])
    assert abs(float(cs.norm_fro(np.eye(3))) - np.sqrt(3)) < 1e-15


def test_dm_prints_like_casadi():
    assert str(cs.DM(1.0192017582309205)) == "1.0192"
    assert str(cs.DM([1, 2, 3])) == "[1, 2, 3]"
    assert "Distance: " + str(cs.norm_2(cs.DM([3.0, 4.0]))) == "Distance: 5"
    assert str(cs.MX.sym("q", 2)).startswith("MX(2x1")


def test_numeric_scalars_are_numpy_elements():
    """DM-valued Function results inside a list assigned to a NumPy row
    (`y_sim[i, :] = [fcos(t), fsine(t), 0]`, ur5_input_experiment.ipynb cell 16)."""
    x = cs.MX.sym("x")
    fcos = cs.Function("fcos", [x], [0.1 * cs.cos(x)])
    y = np.zeros((2, 3))
    y[1, :] = [fcos(0.0), fcos(np.pi), 0]
    assert np.allclose(y[1], [0.1, -0.1, 0.0])
    assert np.asarray(cs.DM(2.5)).shape == () and cs.DM(2.5).toarray().shape == (1, 1)
    assert np.asarray(cs.DM([1.0, 2.0])).shape == (2, 1)
    assert max(min(cs.DM(0.3), 0.2), -0.2) == 0.2          # clipping a DM command like the notebooks do
