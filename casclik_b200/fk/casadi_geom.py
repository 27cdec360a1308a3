"""Symbolic quaternion / dual-quaternion helpers: the part of `urdf2casadi.casadi_geom` the reference
notebooks call (ur5_dual_quaternion_*.ipynb cell 3 wraps each of these in a cs.Function).
urdf2casadi itself is not part of the reference tree; these follow its conventions as the notebooks
show them: quaternions are [x, y, z, w], dual quaternions [real(4); dual(4)] with
dual = 1/2 t (x) real for a rotation followed by a translation t in the parent frame.
Arguments may be symbolic (SX / MX) or numeric; results are cs matrices."""
import numpy as np

from .. import sym as cs
from .converter import quaternion_product, dual_quaternion_product, _axis_unit  # noqa: F401


def _m(x):
    return x if isinstance(x, cs.GenericMatrixCommon) else cs.DM(np.asarray(x, dtype=np.float64))


def quaternion_conj(q):
    q = _m(q)
    return cs.vertcat(-q[0], -q[1], -q[2], q[3])


def dual_quaternion_conj(Q):
    """Quaternion conjugate of both parts (the inverse of a unit dual quaternion)."""
    Q = _m(Q)
    return cs.vertcat(quaternion_conj(Q[:4]), quaternion_conj(Q[4:]))


def dual_quaternion_norm2(Q):
    """Q (x) conj(Q) as a dual number -> (real part |r|^2, dual part 2 r.d)."""
    Q = _m(Q)
    r, d = Q[:4], Q[4:]
    return cs.dot(r, r), 2.0 * cs.dot(r, d)


def dual_quaternion_inv(Q):
    """conj(Q) / norm2(Q), with the dual-number reciprocal 1/(a + eps b) = 1/a - eps b/a^2."""
    Q = _m(Q)
    a, b = dual_quaternion_norm2(Q)
    C = dual_quaternion_conj(Q)
    return cs.vertcat(C[:4] / a, C[4:] / a - C[:4] * (b / (a * a)))


def quaternion_to_rotation(q):
    """3x3 rotation matrix of a (unit) quaternion [x, y, z, w]."""
    q = _m(q)
    x, y, z, w = q[0], q[1], q[2], q[3]
    return cs.vertcat(
        cs.horzcat(1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)),
        cs.horzcat(2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)),
        cs.horzcat(2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)))


def dual_quaternion_to_pos(Q):
    """Translation of a unit dual quaternion: vector part of 2 d (x) conj(r)."""
    Q = _m(Q)
    t = 2.0 * quaternion_product(Q[4:], quaternion_conj(Q[:4]))
    return t[:3]


def dual_quaternion_to_transformation_matrix(Q):
    Q = _m(Q)
    top = cs.horzcat(quaternion_to_rotation(Q[:4]), dual_quaternion_to_pos(Q))
    return cs.vertcat(top, cs.DM(np.array([[0.0, 0.0, 0.0, 1.0]])))


def quaternion_rpy(rpy):
    """Quaternion of R = Rz(yaw) Ry(pitch) Rx(roll) (URDF fixed-axis convention)."""
    rpy = _m(rpy)
    hr, hp, hy = 0.5 * rpy[0], 0.5 * rpy[1], 0.5 * rpy[2]
    cr, sr, cp, sp, cy, sy = cs.cos(hr), cs.sin(hr), cs.cos(hp), cs.sin(hp), cs.cos(hy), cs.sin(hy)
    return cs.vertcat(sr * cp * cy - cr * sp * sy,
                      cr * sp * cy + sr * cp * sy,
                      cr * cp * sy - sr * sp * cy,
                      cr * cp * cy + sr * sp * sy)


def dual_quaternion_rpy(rpy):
    """Pure rotation by roll / pitch / yaw."""
    return cs.vertcat(quaternion_rpy(rpy), cs.DM.zeros(4, 1))


def dual_quaternion_translation(xyz):
    """Pure translation."""
    xyz = _m(xyz)
    return cs.vertcat(cs.DM([0.0, 0.0, 0.0, 1.0]), 0.5 * xyz[0], 0.5 * xyz[1], 0.5 * xyz[2], 0.0)


def _axis3(axis):
    """A numeric axis is normalised; a symbolic one (the notebooks wrap these helpers in Functions of
    a symbolic axis) is taken as given and expected to be of unit length."""
    if isinstance(axis, cs.GenericMatrixCommon) and not axis.is_constant():
        return axis[0], axis[1], axis[2]
    if isinstance(axis, cs.GenericMatrixCommon):
        axis = np.asarray(cs.DM(axis).toarray(), dtype=np.float64).reshape(-1)
    a = _axis_unit(axis)
    return float(a[0]), float(a[1]), float(a[2])


def quaternion_axis_rotation(axis, angle):
    ax, ay, az = _axis3(axis)
    h = 0.5 * (angle if isinstance(angle, cs.GenericMatrixCommon) else cs.DM(angle))
    s, c = cs.sin(h), cs.cos(h)
    return cs.vertcat(ax * s, ay * s, az * s, c)


def dual_quaternion_axis_rotation(axis, angle):
    """Rotation by `angle` about the (constant) axis."""
    return cs.vertcat(quaternion_axis_rotation(axis, angle), cs.DM.zeros(4, 1))


def dual_quaternion_axis_translation(axis, dist):
    """Translation by `dist` along the (constant) axis."""
    ax, ay, az = _axis3(axis)
    d = dist if isinstance(dist, cs.GenericMatrixCommon) else cs.DM(dist)
    return cs.vertcat(cs.DM([0.0, 0.0, 0.0, 1.0]), 0.5 * ax * d, 0.5 * ay * d, 0.5 * az * d, 0.0)


def dual_quaternion_revolute(xyz, rpy, axis, angle):
    """Joint origin (translation xyz, then rotation rpy) followed by a rotation about `axis`."""
    origin = dual_quaternion_product(dual_quaternion_translation(xyz), dual_quaternion_rpy(rpy))
    return dual_quaternion_product(origin, dual_quaternion_axis_rotation(axis, angle))


def dual_quaternion_prismatic(xyz, rpy, axis, dist):
    """Joint origin followed by a translation along `axis`."""
    origin = dual_quaternion_product(dual_quaternion_translation(xyz), dual_quaternion_rpy(rpy))
    return dual_quaternion_product(origin, dual_quaternion_axis_translation(axis, dist))
