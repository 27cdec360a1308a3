"""Pose-error expressions used as constraint expressions in the reference notebooks
(ur5_dual_quaternion_vs_transformation_matrix.ipynb cells 3, 18, 20;
ur5_transformation_matrix_comparison_of_controllers.ipynb cells 36-37), as library functions.

T is a 4x4 homogeneous transform expression (e.g. `T_fk(q)`), Q a dual quaternion expression
([qx qy qz qw | dx dy dz dw], e.g. `dual_quaternion_fk(q)`); *_des are numeric or symbolic targets.
"""
import numpy as np

from .. import sym as cs
from .converter import quaternion_product, dual_quaternion_product


def _m(x):
    return x if isinstance(x, cs.GenericMatrixCommon) else cs.DM(np.asarray(x, dtype=np.float64))


def T_dist1(T, T_des):
    """|| T_des^-1 T - I ||_F  (1 row)."""
    return cs.norm_fro(cs.mtimes(cs.inv(_m(T_des)), T) - np.eye(4))


def T_dist2(T, T_des):
    """[p - p_des ; || R_des^-1 R - I ||_F]  (4 rows)."""
    T_des = _m(T_des)
    return cs.vertcat(T[:3, 3] - T_des[:3, 3],
                      cs.norm_fro(cs.mtimes(cs.inv(T_des[:3, :3]), T[:3, :3]) - np.eye(3)))


def T_dist3(T, T_des):
    """Three-point strategy: rows of R (as points) + p against the target's  (9 rows)."""
    T_des = _m(T_des)
    return cs.vertcat(*[T[i, :3].T + T[:3, 3] - T_des[i, :3].T - T_des[:3, 3] for i in range(3)])


def quaternion_conj(q):
    q = _m(q)
    return cs.vertcat(-q[0], -q[1], -q[2], q[3])


def dual_quaternion_conj(Q):
    Q = _m(Q)
    return cs.vertcat(quaternion_conj(Q[:4]), quaternion_conj(Q[4:]))


def hamilton_operator_minus(q):
    """H-(q): p (x) q = H-(q) p, [x y z w] layout."""
    q = _m(q)
    x, y, z, w = q[0], q[1], q[2], q[3]
    return cs.vertcat(cs.horzcat(w, z, -y, x),
                      cs.horzcat(-z, w, x, y),
                      cs.horzcat(y, -x, w, z),
                      cs.horzcat(-x, -y, -z, w))


def dual_hamilton_operator_minus(Q):
    """A (x) Q = dualH-(Q) A for dual quaternions."""
    Q = _m(Q)
    Hr, Hd = hamilton_operator_minus(Q[:4]), hamilton_operator_minus(Q[4:])
    Z = cs.DM.zeros(4, 4)
    return cs.vertcat(cs.horzcat(Hr, Z), cs.horzcat(Hd, Hr))


Q_IDENTITY = np.array([0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0])


def Q_dist1(Q, Q_des):
    """Q (x) conj(Q_des) - identity  (8 rows)."""
    return dual_quaternion_product(Q, dual_quaternion_conj(Q_des)) - Q_IDENTITY


def Q_dist2(Q, Q_des):
    """dualH-(Q_des) C (Q_des - Q), C = diag(-1,-1,-1,1,-1,-1,-1,1)  (8 rows)."""
    Q_des = _m(Q_des)
    C = cs.diag(cs.DM([-1., -1., -1., 1., -1., -1., -1., 1.]))
    return cs.mtimes(dual_hamilton_operator_minus(Q_des), cs.mtimes(C, Q_des - Q))


__all__ = ["T_dist1", "T_dist2", "T_dist3", "Q_dist1", "Q_dist2", "quaternion_conj",
           "dual_quaternion_conj", "hamilton_operator_minus", "dual_hamilton_operator_minus",
           "quaternion_product", "dual_quaternion_product"]
