"""Numeric (NumPy) counterparts of casadi_geom: the part of `urdf2casadi.numpy_geom` the reference
notebooks call (desired frames, identity dual quaternion, dual quaternion -> 4x4).  Same conventions:
quaternions [x, y, z, w], dual quaternions [real(4); dual(4)], R = Rz(yaw) Ry(pitch) Rx(roll)."""
import numpy as np

from .converter import rotation_rpy, _axis_unit  # noqa: F401


def T_rpy(xyz, roll, pitch, yaw):
    """4x4 homogeneous transform: rotation rpy, translation xyz."""
    T = np.eye(4)
    T[:3, :3] = rotation_rpy(roll, pitch, yaw)
    T[:3, 3] = np.asarray(xyz, dtype=np.float64).reshape(3)
    return T


def quaternion_product(p, q):
    p, q = np.asarray(p, dtype=np.float64).reshape(4), np.asarray(q, dtype=np.float64).reshape(4)
    px, py, pz, pw = p
    qx, qy, qz, qw = q
    return np.array([pw * qx + px * qw + py * qz - pz * qy,
                     pw * qy - px * qz + py * qw + pz * qx,
                     pw * qz + px * qy - py * qx + pz * qw,
                     pw * qw - px * qx - py * qy - pz * qz])


def quaternion_conj(q):
    q = np.asarray(q, dtype=np.float64).reshape(4)
    return np.array([-q[0], -q[1], -q[2], q[3]])


def dual_quaternion_product(A, B):
    A, B = np.asarray(A, dtype=np.float64).reshape(8), np.asarray(B, dtype=np.float64).reshape(8)
    return np.concatenate([quaternion_product(A[:4], B[:4]),
                           quaternion_product(A[:4], B[4:]) + quaternion_product(A[4:], B[:4])])


def dual_quaternion_conj(Q):
    Q = np.asarray(Q, dtype=np.float64).reshape(8)
    return np.concatenate([quaternion_conj(Q[:4]), quaternion_conj(Q[4:])])


def quaternion_rpy(roll, pitch, yaw):
    cr, sr, cp, sp = np.cos(0.5 * roll), np.sin(0.5 * roll), np.cos(0.5 * pitch), np.sin(0.5 * pitch)
    cy, sy = np.cos(0.5 * yaw), np.sin(0.5 * yaw)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


def dual_quaternion_translation(xyz):
    x, y, z = np.asarray(xyz, dtype=np.float64).reshape(3)
    return np.array([0.0, 0.0, 0.0, 1.0, 0.5 * x, 0.5 * y, 0.5 * z, 0.0])


def dual_quaternion_rpy(rpy):
    return np.concatenate([quaternion_rpy(*np.asarray(rpy, dtype=np.float64).reshape(3)), np.zeros(4)])


def dual_quaternion_axis_rotation(axis, angle):
    a = _axis_unit(axis)
    return np.concatenate([a * np.sin(0.5 * angle), [np.cos(0.5 * angle)], np.zeros(4)])


def dual_quaternion_axis_translation(axis, dist):
    a = _axis_unit(axis)
    return np.concatenate([[0.0, 0.0, 0.0, 1.0], 0.5 * dist * a, [0.0]])


def dual_quaternion_revolute(xyz, rpy, axis, angle):
    """Joint origin (translation xyz, then rotation rpy) followed by a rotation about `axis`."""
    origin = dual_quaternion_product(dual_quaternion_translation(xyz), dual_quaternion_rpy(rpy))
    return dual_quaternion_product(origin, dual_quaternion_axis_rotation(axis, angle))


def dual_quaternion_prismatic(xyz, rpy, axis, dist):
    """Joint origin followed by a translation along `axis`."""
    origin = dual_quaternion_product(dual_quaternion_translation(xyz), dual_quaternion_rpy(rpy))
    return dual_quaternion_product(origin, dual_quaternion_axis_translation(axis, dist))


def dual_quaternion_to_pos(Q):
    Q = np.asarray(Q, dtype=np.float64).reshape(8)
    return 2.0 * quaternion_product(Q[4:], quaternion_conj(Q[:4]))[:3]


def dual_quaternion_to_transformation_matrix(Q):
    Q = np.asarray(Q, dtype=np.float64).reshape(8)
    x, y, z, w = Q[:4]
    T = np.eye(4)
    T[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    T[:3, 3] = dual_quaternion_to_pos(Q)
    return T
