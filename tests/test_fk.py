"""FK front-end: the notebooks' known answers and agreement with the oracle's
independent numeric FK / geometric Jacobian."""
import math

import numpy as np

from oracle_bridge import orc
from casclik_b200 import cs, fk

UR5_HOME = [0.0, -math.pi / 2, 0.0, -math.pi / 2, 0.0, 0.0]


def test_ur5_home_known_answers():
    d = fk.ur5()
    T = d["T_fk"](UR5_HOME).toarray()
    assert abs(np.linalg.norm(T[:3, 3]) - 1.0192) < 5e-5          # notebook prints 1.0192
    assert abs(np.linalg.norm(T[:3, 3]) - 1.0192017582309205) < 1e-14
    assert np.allclose(T[:3, 3], [0.0, 0.19145, 1.001059], atol=1e-9)
    assert np.allclose(T[:3, :3], [[1, 0, 0], [0, 0, 1], [0, -1, 0]], atol=1e-9)
    Q = d["dual_quaternion_fk"](UR5_HOME).toarray()[:, 0]
    golden = np.array([-0.707107, 0, 0, 0.707107, 0, -0.28624, 0.421616, 0])   # notebook print-out
    assert min(np.abs(Q - golden).max(), np.abs(Q + golden).max()) < 5e-6
    assert np.allclose(d["quaternion_fk"](UR5_HOME).toarray()[:, 0], Q[:4])


def test_joint_limits_and_names_from_urdf():
    d = fk.ur5()
    assert np.allclose(d["upper"], [6.28318531, 6.28318531, 3.14159265, 6.28318531, 6.28318531, 6.28318531])
    assert np.allclose(d["lower"], -np.array(d["upper"]))
    assert d["joint_names"] == ["shoulder_pan_joint", "shoulder_lift_joint", "elbow_joint",
                                "wrist_1_joint", "wrist_2_joint", "wrist_3_joint"]
    i = fk.iiwa14()
    assert np.allclose(i["upper"], [2.9668, 2.0942, 2.9668, 2.0942, 2.9668, 2.0942, 3.0541])
    assert len(i["joint_names"]) == 7
    assert np.allclose(i["T_fk"]([0.0] * 7).toarray()[:3, 3], [0, 0, 0.36 + 0.42 + 0.4 + 0.126])


def test_expression_fk_matches_independent_numeric_fk():
    rng = np.random.default_rng(0)
    for urdf, build, n in ((fk.UR5_URDF, fk.ur5, 6), (fk.IIWA14_URDF, fk.iiwa14, 7)):
        d = build()
        chain = orc.load_chain(urdf, "base_link", "tool0")
        q = rng.uniform(-2.0, 2.0, size=(50, n))
        R, p, _, _ = orc.fk_pose(chain, q)
        qs = cs.MX.sym("q", n)
        T = d["T_fk"](qs)
        Jp = cs.Function("Jp", [qs], [cs.jacobian(T[:3, 3], qs), cs.jacobian(cs.vec(T[:3, :3]), qs)])
        _, Jgeo = orc.position_jacobian(chain, q)
        _, dR = orc.rotation_jacobian(chain, q)
        for k in range(50):
            Tn = d["T_fk"](q[k]).toarray()
            assert np.abs(Tn[:3, :3] - R[k]).max() < 1e-14 and np.abs(Tn[:3, 3] - p[k]).max() < 1e-14
            Jn, JRn = (o.toarray() for o in Jp(q[k]))
            assert np.abs(Jn - Jgeo[k]).max() < 1e-13
            for j in range(n):
                assert np.abs(JRn[:, j].reshape(3, 3, order="F") - dR[k, j]).max() < 1e-13


def test_denavit_hartenberg_ur5():
    d = fk.from_denavit_hartenberg(["s"] * 6, [0., -0.425, -0.392, 0., 0., 0.],
                                   [0.089, 0., 0., 0.109, 0.095, 0.082],
                                   [math.pi / 2, 0., 0., math.pi / 2, -math.pi / 2, 0.])
    T = d["T_fk"]([0.0] * 6).toarray()
    # all joints at zero: arm stretched along -x, wrist offsets along y / z
    assert np.allclose(T[:3, 3], [-0.817, -0.191, -0.006], atol=1e-12)
    assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-14
    q = np.array([0.3, -1.0, 0.7, 0.2, -0.4, 1.1])
    Tq = d["T_fk"](q).toarray()
    assert np.allclose(Tq[:3, :3] @ Tq[:3, :3].T, np.eye(3), atol=1e-14)


def test_position_only_skill_prunes_to_a_small_program():
    """Tip-to-root accumulation leaves ~115 flops for UR5 position + Jacobian."""
    from casclik_b200.sym import dag
    d = fk.ur5()
    q = cs.MX.sym("q", 6)
    p = d["T_fk"](q)[:3, 3]
    J = cs.jacobian(p, q)
    h = dag.op_histogram(p.nodes() + J.nodes())
    flops = sum(h.get(k, 0) for k in ("add", "sub", "mul", "div", "sqrt"))
    assert flops <= 130 and h["sin"] == 5 and h["cos"] == 5
    assert all(n is dag.ZERO for n in J._a[:, 5])      # tool0 position does not depend on q6


def test_pose_error_forms_of_the_notebooks():
    from casclik_b200.fk import pose_errors as pe
    d = fk.ur5()
    q = cs.MX.sym("q", 6)
    T, Q = d["T_fk"](q), d["dual_quaternion_fk"](q)
    q_des = np.array([0.3, -1.2, 0.9, -0.4, 0.5, 0.1])
    T_des = d["T_fk"](q_des).toarray()
    Q_des = d["dual_quaternion_fk"](q_des).toarray()[:, 0]
    forms = {"T_dist1": (pe.T_dist1(T, T_des), 1), "T_dist2": (pe.T_dist2(T, T_des), 4),
             "T_dist3": (pe.T_dist3(T, T_des), 9), "Q_dist1": (pe.Q_dist1(Q, Q_des), 8),
             "Q_dist2": (pe.Q_dist2(Q, Q_des), 8)}
    f = cs.Function("f", [q], [e for e, _ in forms.values()])
    at_target = f(q_des)
    away = f(q_des + 0.2)
    for (name, (e, rows)), z, a in zip(forms.items(), at_target, away):
        assert e.shape == (rows, 1), name
        assert np.abs(z.toarray()).max() < 1e-12, name          # every form vanishes at the target
        assert np.abs(a.toarray()).max() > 1e-3, name
    # T_dist2 against NumPy
    Tn = d["T_fk"](q_des + 0.2).toarray()
    ref = np.concatenate([Tn[:3, 3] - T_des[:3, 3],
                          [np.linalg.norm(np.linalg.inv(T_des[:3, :3]) @ Tn[:3, :3] - np.eye(3))]])
    assert np.abs(away[1].toarray()[:, 0] - ref).max() < 1e-13
    # dual-quaternion algebra: A (x) B == dualH-(B) A, and unit norm of the FK quaternion
    A, B = d["dual_quaternion_fk"](q_des + 0.2).toarray()[:, 0], Q_des
    lhs = pe.dual_quaternion_product(A, B).toarray()[:, 0]
    rhs = (pe.dual_hamilton_operator_minus(B).toarray() @ A)
    assert np.abs(lhs - rhs).max() < 1e-14 and abs(np.linalg.norm(A[:4]) - 1.0) < 1e-14


def test_geometry_helper_modules_of_the_notebooks():
    """fk.casadi_geom / fk.numpy_geom stand in for urdf2casadi.casadi_geom / numpy_geom as the
    dual-quaternion notebooks use them (wrapped in cs.Functions of SX symbols, desired frames,
    identity dual quaternion); conventions are pinned by the FK itself: [x y z w | dual]."""
    from casclik_b200.fk import casadi_geom as cg, numpy_geom as ng
    d = fk.ur5()
    rng = np.random.default_rng(3)
    quat1, quat2 = cs.SX.sym("quat1", 8), cs.SX.sym("quat2", 8)
    dq_prod = cs.Function("dualquatprod", [quat1, quat2], [cg.dual_quaternion_product(quat1, quat2)])
    dq_conj = cs.Function("dualquatconj", [quat1], [cg.dual_quaternion_conj(quat1)])
    dq_inv = cs.Function("dualquatinv", [quat1], [cg.dual_quaternion_inv(quat1)])
    dq_T = cs.Function("dualquat2transfmat", [quat1], [cg.dual_quaternion_to_transformation_matrix(quat1)])
    dq_pos = cs.Function("dualquat2pos", [quat1], [cg.dual_quaternion_to_pos(quat1)])
    dq_norm = cs.Function("dualquatdualnorm", [quat1], [cg.dual_quaternion_norm2(quat1)[0],
                                                         cg.dual_quaternion_norm2(quat1)[1]])
    axis, ang = cs.SX.sym("axis", 3), cs.SX.sym("ang")
    dq_rot = cs.Function("dualquataxisrot", [axis, ang], [cg.dual_quaternion_axis_rotation(axis, ang)])
    dq_tr = cs.Function("dualquataxistransl", [axis, ang], [cg.dual_quaternion_axis_translation(axis, ang)])
    identity = np.array([0, 0, 0, 1, 0, 0, 0, 0.0])
    assert np.array_equal(ng.dual_quaternion_revolute([0., 0., 0.], [0., 0., 0.], [1., 0., 0.], 0.0), identity)
    for _ in range(10):
        q = rng.uniform(-3, 3, 6)
        Q = d["dual_quaternion_fk"](q)
        T = np.asarray(d["T_fk"](q).toarray())
        assert np.abs(np.asarray(dq_T(Q).toarray()) - T).max() < 1e-13
        assert np.abs(ng.dual_quaternion_to_transformation_matrix(Q.toarray()) - T).max() < 1e-13
        assert np.abs(np.asarray(dq_pos(Q).toarray())[:, 0] - T[:3, 3]).max() < 1e-13
        assert np.abs(np.asarray(dq_prod(Q, dq_conj(Q)).toarray())[:, 0] - identity).max() < 1e-13
        assert np.abs(np.asarray(dq_prod(Q, dq_inv(Q)).toarray())[:, 0] - identity).max() < 1e-13
        a, b = dq_norm(Q)
        assert abs(float(a) - 1.0) < 1e-13 and abs(float(b)) < 1e-13
        xyz, rpy = rng.uniform(-1, 1, 3), rng.uniform(-3, 3, 3)
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        th = rng.uniform(-3, 3)
        # joint transform two ways: dual quaternions vs homogeneous matrices
        Qj = ng.dual_quaternion_revolute(xyz, rpy, ax, th)
        Rj = np.asarray(dq_T(dq_rot(ax, th)).toarray())
        assert np.abs(ng.dual_quaternion_to_transformation_matrix(Qj) - ng.T_rpy(xyz, *rpy) @ Rj).max() < 1e-13
        assert np.abs(np.asarray(cs.DM(cg.dual_quaternion_revolute(xyz, rpy, ax, th)).toarray())[:, 0] - Qj).max() < 1e-14
        Qp = ng.dual_quaternion_prismatic(xyz, rpy, ax, 0.3)
        Tp = ng.T_rpy(xyz, *rpy) @ np.asarray(dq_T(dq_tr(ax, 0.3)).toarray())
        assert np.abs(ng.dual_quaternion_to_transformation_matrix(Qp) - Tp).max() < 1e-13
        assert np.abs(ng.rotation_rpy(*rpy) - ng.T_rpy(xyz, *rpy)[:3, :3]).max() == 0.0
        assert np.abs(np.asarray(cs.DM(cg.dual_quaternion_rpy(rpy)).toarray())[:, 0] - ng.dual_quaternion_rpy(rpy)).max() < 1e-15
        assert np.abs(np.asarray(cs.DM(cg.dual_quaternion_translation(xyz)).toarray())[:, 0]
                      - ng.dual_quaternion_translation(xyz)).max() == 0.0
