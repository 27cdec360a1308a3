"""URDF / Denavit-Hartenberg -> forward-kinematics expressions.

Stands in for the un-vendored `urdf2casadi.converter` the reference notebooks call
(`converter.from_file(root, tip, filename)`, e.g. ur5_transformation_matrix_comparison_of_
controllers.ipynb cell 4; `converter.from_denavit_hartenberg(...)`,
ur5_moe2016_example2.ipynb cell 2).  Returned dict keys follow what the notebooks read:
`T_fk`, `quaternion_fk`, `dual_quaternion_fk`, `q`, `upper`, `lower`, `joint_names`, `joint_list`.

Conventions (validated against the notebooks' known answer |p(UR5_home)| = 1.0192):
  * joint transform = T_origin(xyz, rpy) * Rot(axis, q_i); rpy is fixed-axis XYZ, R = Rz(y) Ry(p) Rx(r);
  * fixed joints contribute T_origin only; prismatic joints T_origin * Trans(axis * q_i);
  * dual quaternion layout [qx qy qz qw | dx dy dz dw], dual = 1/2 * t (x) r.

The chain is accumulated from the tip towards the root with origin and joint rotation as separate
factors.  That ordering makes the translation column a chain of matrix-vector products, so when a
skill only uses `T_fk(q)[:3, 3]` dead-code elimination leaves ~20 flops per joint instead of a
4x4 product (this is where the op count in DESIGN.md comes from).
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

from .. import sym as cs
from ..sym import dag
from ..sym.matrix import _filled


# ----------------------------------------------------------------------------------------------
# URDF parsing
# ----------------------------------------------------------------------------------------------

class Joint(object):
    def __init__(self, name, jtype, parent, child, xyz, rpy, axis, lower, upper):
        self.name, self.type, self.parent, self.child = name, jtype, parent, child
        self.xyz, self.rpy, self.axis = xyz, rpy, axis
        self.lower, self.upper = lower, upper

    @property
    def actuated(self):
        return self.type in ("revolute", "continuous", "prismatic")

    def __repr__(self):
        return "Joint(%s %s %s->%s)" % (self.name, self.type, self.parent, self.child)


def _floats(text, default):
    if text is None:
        return tuple(default)
    return tuple(float(v) for v in text.split())


def parse_urdf(filename):
    """-> list of Joint in file order."""
    root = ET.parse(filename).getroot()
    joints = []
    for j in root.findall("joint"):
        origin = j.find("origin")
        axis = j.find("axis")
        limit = j.find("limit")
        xyz = _floats(origin.get("xyz") if origin is not None else None, (0, 0, 0))
        rpy = _floats(origin.get("rpy") if origin is not None else None, (0, 0, 0))
        ax = _floats(axis.get("xyz") if axis is not None else None, (1, 0, 0))
        jtype = j.get("type")
        lower = upper = None
        if limit is not None and limit.get("lower") is not None:
            lower, upper = float(limit.get("lower")), float(limit.get("upper"))
        if jtype == "continuous":
            lower, upper = -math.inf, math.inf
        joints.append(Joint(j.get("name"), jtype, j.find("parent").get("link"),
                            j.find("child").get("link"), xyz, rpy, ax, lower, upper))
    return joints


def chain(joints, root, tip):
    """Joints on the path root -> tip (walks child->parent from the tip)."""
    by_child = {j.child: j for j in joints}
    path = []
    link = tip
    while link != root:
        if link not in by_child:
            raise ValueError("no kinematic path from %r to %r" % (root, tip))
        j = by_child[link]
        path.append(j)
        link = j.parent
    return list(reversed(path))


# ----------------------------------------------------------------------------------------------
# numeric pieces
# ----------------------------------------------------------------------------------------------

def rotation_rpy(roll, pitch, yaw):
    """Fixed-axis XYZ: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr = math.cos(roll), math.sin(roll)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cy, sy = math.cos(yaw), math.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def origin_matrix(xyz, rpy):
    T = np.eye(4)
    if any(v != 0.0 for v in rpy):
        T[:3, :3] = rotation_rpy(*rpy)
    T[:3, 3] = xyz
    return T


def quaternion_from_matrix(R):
    """[x, y, z, w] of a rotation matrix (numeric)."""
    tr = np.trace(R)
    if tr > 0:
        s = 2.0 * math.sqrt(1.0 + tr)
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s,
                         (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = 2.0 * math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k])
    q = np.zeros(4)
    q[i] = 0.25 * s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    q[3] = (R[k, j] - R[j, k]) / s
    return q


# ----------------------------------------------------------------------------------------------
# symbolic pieces
# ----------------------------------------------------------------------------------------------

def _axis_unit(axis):
    a = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(a)
    if n == 0.0:
        raise ValueError("zero joint axis")
    return a / n


def rotation_axis_angle(axis, angle):
    """4x4 symbolic Rot(axis, angle); exact sparsity for coordinate axes."""
    a = _axis_unit(axis)
    th = angle if isinstance(angle, cs.GenericMatrixCommon) else cs.DM(angle)
    for k in range(3):
        if abs(abs(a[k]) - 1.0) < 1e-15:  # coordinate axis (possibly negative)
            if a[k] < 0:
                th = -th
            c, s = cs.cos(th), cs.sin(th)
            i, j = (k + 1) % 3, (k + 2) % 3
            T = cs.SX.eye(4)
            T[i, i] = c
            T[i, j] = -s
            T[j, i] = s
            T[j, j] = c
            return T
    c, s = cs.cos(th), cs.sin(th)
    v = 1.0 - c
    x, y, z = (float(e) for e in a)
    T = cs.SX.eye(4)
    T[0, 0] = c + x * x * v
    T[0, 1] = x * y * v - z * s
    T[0, 2] = x * z * v + y * s
    T[1, 0] = y * x * v + z * s
    T[1, 1] = c + y * y * v
    T[1, 2] = y * z * v - x * s
    T[2, 0] = z * x * v - y * s
    T[2, 1] = z * y * v + x * s
    T[2, 2] = c + z * z * v
    return T


def translation_axis(axis, dist):
    a = _axis_unit(axis)
    T = cs.SX.eye(4)
    for k in range(3):
        T[k, 3] = float(a[k]) * dist
    return T


def _affine_times(A, B):
    """Product of two 4x4 homogeneous transforms, exploiting the [0 0 0 1] bottom rows."""
    a, b = A._a, B._a
    out = _filled(4, 4, dag.ZERO)
    out[3, 3] = dag.ONE
    for i in range(3):
        for j in range(4):
            acc = dag.ZERO
            for k in range(3):
                acc = dag.add(acc, dag.mul(a[i, k], b[k, j]))
            if j == 3:
                acc = dag.add(acc, a[i, 3])
            out[i, j] = acc
    return cs.SX._wrap(out)


def _tip_frame_joint(S, arg, kind, axis):
    """Axis direction and a point on the axis of one joint in TIP coordinates, from the suffix transform
    S = (joint frame -> tip): a_b = R_s' a, o_b = -R_s' p_s.  Ingredients of dag.ChainBlock."""
    a = _axis_unit(axis)
    s = S._a
    ab, ob = [], []
    for k in range(3):
        acc_a, acc_o = dag.ZERO, dag.ZERO
        for i in range(3):
            acc_a = dag.add(acc_a, dag.mul(s[i, k], dag.const(float(a[i]))))
            acc_o = dag.add(acc_o, dag.mul(s[i, k], s[i, 3]))
        ab.append(acc_a)
        ob.append(dag.neg(acc_o))
    return (arg.nodes()[0], kind, ab, ob)


def _register_block(T, joints):
    """Tell the AD layer that T = [R p] is a kinematic chain (closed-form tip-frame partials)."""
    t = T._a
    dag.register_chain_block(dag.ChainBlock([[t[i, k] for k in range(4)] for i in range(3)], joints))


def quaternion_product(p, q):
    """Hamilton product, [x y z w] layout, symbolic or numeric column vectors."""
    p = p if isinstance(p, cs.GenericMatrixCommon) else cs.DM(p)
    q = q if isinstance(q, cs.GenericMatrixCommon) else cs.DM(q)
    px, py, pz, pw = p[0], p[1], p[2], p[3]
    qx, qy, qz, qw = q[0], q[1], q[2], q[3]
    return cs.vertcat(pw * qx + px * qw + py * qz - pz * qy,
                      pw * qy - px * qz + py * qw + pz * qx,
                      pw * qz + px * qy - py * qx + pz * qw,
                      pw * qw - px * qx - py * qy - pz * qz)


def dual_quaternion_product(A, B):
    A = A if isinstance(A, cs.GenericMatrixCommon) else cs.DM(A)
    B = B if isinstance(B, cs.GenericMatrixCommon) else cs.DM(B)
    ar, ad, br, bd = A[:4], A[4:], B[:4], B[4:]
    return cs.vertcat(quaternion_product(ar, br),
                      quaternion_product(ar, bd) + quaternion_product(ad, br))


def _dq_from_rt(quat, xyz):
    """Dual quaternion of rotation `quat` followed (in the parent frame) by translation xyz."""
    t = cs.vertcat(xyz[0], xyz[1], xyz[2], 0.0)
    return cs.vertcat(quat, 0.5 * quaternion_product(t, quat))


def _axis_angle_quaternion(axis, angle):
    a = _axis_unit(axis)
    h = 0.5 * angle
    s, c = cs.sin(h), cs.cos(h)
    return cs.vertcat(float(a[0]) * s, float(a[1]) * s, float(a[2]) * s, c)


# ----------------------------------------------------------------------------------------------
# public builders
# ----------------------------------------------------------------------------------------------

def _limits(joint):
    lo = -math.inf if joint.lower is None else joint.lower
    hi = math.inf if joint.upper is None else joint.upper
    return lo, hi


def from_joint_list(path):
    """Build the FK dictionary from an ordered list of Joint objects (root -> tip)."""
    act = [j for j in path if j.actuated]
    n = len(act)
    q = cs.SX.sym("q", n)
    index = {j.name: i for i, j in enumerate(act)}

    # homogeneous transform, accumulated tip -> root with split factors
    U = cs.SX.eye(4)
    joints = []
    for j in reversed(path):
        if j.type in ("revolute", "continuous"):
            U = _affine_times(rotation_axis_angle(j.axis, q[index[j.name]]), U)
            joints.append(_tip_frame_joint(U, q[index[j.name]], "revolute", j.axis))
        elif j.type == "prismatic":
            U = _affine_times(translation_axis(j.axis, q[index[j.name]]), U)
            joints.append(_tip_frame_joint(U, q[index[j.name]], "prismatic", j.axis))
        elif j.type != "fixed":
            raise NotImplementedError("joint type %r" % j.type)
        U = _affine_times(cs.SX(origin_matrix(j.xyz, j.rpy)), U)
    _register_block(U, joints)

    # (dual) quaternion, accumulated root -> tip
    Q = cs.DM([0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0])
    for j in path:
        r0 = quaternion_from_matrix(origin_matrix(j.xyz, j.rpy)[:3, :3])
        Q = dual_quaternion_product(Q, _dq_from_rt(cs.DM(r0), j.xyz))
        if j.type in ("revolute", "continuous"):
            rq = _axis_angle_quaternion(j.axis, q[index[j.name]])
            Q = dual_quaternion_product(Q, cs.vertcat(rq, cs.DM.zeros(4)))
        elif j.type == "prismatic":
            a = _axis_unit(j.axis)
            d = q[index[j.name]]
            Q = dual_quaternion_product(
                Q, _dq_from_rt(cs.DM([0.0, 0.0, 0.0, 1.0]), [float(a[k]) * d for k in range(3)]))

    lims = [_limits(j) for j in act]
    return {
        "joint_names": [j.name for j in act],
        "joint_list": [j.name for j in path],
        "q": q,
        "lower": [l for l, _ in lims],
        "upper": [u for _, u in lims],
        "T_fk": cs.Function("T_fk", [q], [U], ["q"], ["T_fk"]),
        "quaternion_fk": cs.Function("quaternion_fk", [q], [Q[:4]], ["q"], ["quaternion_fk"]),
        "dual_quaternion_fk": cs.Function("dual_quaternion_fk", [q], [Q], ["q"],
                                          ["dual_quaternion_fk"]),
    }


def from_file(root, tip, filename):
    """urdf2casadi-style entry point."""
    return from_joint_list(chain(parse_urdf(filename), root, tip))


def from_denavit_hartenberg(joint_angles, link_lengths, link_offsets, link_twists,
                            joint_names=None, upper_limits=None, lower_limits=None):
    """Classic DH: A_i = Rz(theta_i) Tz(d_i) Tx(a_i) Rx(alpha_i).  An entry of `joint_angles` /
    `link_offsets` equal to the string "s" marks the actuated quantity of that joint (the
    notebooks pass joint_angles=["s"]*6, ur5_moe2016_example2.ipynb cell 2)."""
    n = len(link_lengths)
    q = cs.SX.sym("q", n)
    U = cs.SX.eye(4)
    joints = []
    for i in reversed(range(n)):
        theta = q[i] if isinstance(joint_angles[i], str) else joint_angles[i]
        d = q[i] if isinstance(link_offsets[i], str) else link_offsets[i]
        ca, sa = math.cos(link_twists[i]), math.sin(link_twists[i])
        # Tz(d) Tx(a) Rx(alpha) as one transform; constant unless the joint is prismatic
        B = cs.SX.eye(4)
        B[1, 1], B[1, 2], B[2, 1], B[2, 2] = ca, -sa, sa, ca
        B[0, 3] = link_lengths[i]
        B[2, 3] = d
        U = _affine_times(B, U)
        if isinstance(link_offsets[i], str):
            joints.append(_tip_frame_joint(U, q[i], "prismatic", (0, 0, 1)))
        U = _affine_times(rotation_axis_angle((0, 0, 1), theta), U)
        if isinstance(joint_angles[i], str):
            joints.append(_tip_frame_joint(U, q[i], "revolute", (0, 0, 1)))
    _register_block(U, joints)
    names = list(joint_names) if joint_names is not None else ["joint_%d" % i for i in range(n)]
    return {
        "joint_names": names,
        "joint_list": names,
        "q": q,
        "lower": list(lower_limits) if lower_limits is not None else [-math.inf] * n,
        "upper": list(upper_limits) if upper_limits is not None else [math.inf] * n,
        "T_fk": cs.Function("T_fk", [q], [U], ["q"], ["T_fk"]),
    }
