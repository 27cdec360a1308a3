#!/usr/bin/env python
"""Independent pins for the two things tests/golden/controller_vectors.json cannot pin by itself
("What the reference-generated fixture does and does not pin"):

(i)  THE EXPRESSION LAYER.  The fixture generator runs the reference's controller code over a stand-in
     `casadi` that is the product's own symbolic layer (casclik_b200.sym), so an error in its graph
     construction, AD or evaluation would sit in fixture, oracle bridge and kernels alike.  Here every
     skill of the catalogue (tests/golden_skills.py) is rebuilt with SYMPY as the symbolic engine — a
     small sympy-backed `cs` module below, and forward kinematics re-derived from the URDF / DH tables
     with sympy matrices (urdf: T = prod T_origin * Rot(axis, q), oracle/clik_oracle.py:load_chain parses
     the file; DH: Rz Tz Tx Rx) — and e, de/d[q; x], de/dt (sympy.diff) plus expression-valued gains,
     bounds and targets are evaluated at the fixture's first instances.  Nothing of casclik_b200 is
     imported by this script.  tests/test_independent_pins.py compares the product's lowering with them.
     Covers norm_fro, reshape, matrix / expression gains, DH and URDF chains, time-varying paths.

(ii) THE FIXTURE QPs.  The fixture's minimisers come from the oracle's own active-set solver.  Every
     fixture QP (reference-built H, A, lba, uba) is solved a second time with scipy.optimize (trust-constr
     interior point, started from zero) and the result stored next to it.

Output: tests/golden/independent_pins.json.   Run:  python tests/golden/make_independent_pins.py
"""
import json
import math
import os
import sys

import numpy as np
import sympy as sp
from scipy.optimize import minimize

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
N_INST = 6


# ---- a sympy-backed stand-in for the `cs` surface the catalogue uses -----------------------------------

def _mat(x):
    if isinstance(x, M):
        return x.m
    if isinstance(x, sp.MatrixBase):
        return sp.Matrix(x)
    if isinstance(x, sp.Basic):
        return sp.Matrix([[x]])
    if isinstance(x, (list, tuple)) and any(isinstance(v, M) for v in x):
        return sp.Matrix.vstack(*[_mat(v) for v in x])
    a = np.asarray(x, dtype=object)
    if a.ndim == 0:
        return sp.Matrix([[sp.Float(float(x), 17)]])
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return sp.Matrix(a.shape[0], a.shape[1],
                     lambda i, j: a[i, j] if isinstance(a[i, j], sp.Basic) else sp.Float(float(a[i, j]), 17))


def _elementwise(a, b, op):
    A, B = _mat(a), _mat(b)
    if A.shape == (1, 1) and B.shape != (1, 1):
        A = sp.Matrix(B.shape[0], B.shape[1], lambda i, j: A[0, 0])
    if B.shape == (1, 1) and A.shape != (1, 1):
        B = sp.Matrix(A.shape[0], A.shape[1], lambda i, j: B[0, 0])
    if A.shape != B.shape:
        raise ValueError("shape mismatch %s vs %s" % (A.shape, B.shape))
    return M(sp.Matrix(A.shape[0], A.shape[1], lambda i, j: op(A[i, j], B[i, j])))


class M(object):
    """Dense symbolic matrix with CasADi's operator semantics (elementwise, 1x1 broadcasts)."""
    __array_priority__ = 1000

    def __init__(self, m):
        self.m = sp.Matrix(m) if isinstance(m, sp.MatrixBase) else _mat(m)

    @staticmethod
    def sym(name, n=1, m=1):
        if n == 1 and m == 1:
            return M(sp.Matrix([[sp.Symbol(name, real=True)]]))
        return M(sp.Matrix(n, m, lambda i, j: sp.Symbol("%s_%d" % (name, i + n * j), real=True)))

    shape = property(lambda self: self.m.shape)

    def size(self):
        return self.m.shape

    @property
    def T(self):
        return M(self.m.T)

    def __getitem__(self, k):
        if not isinstance(k, tuple):
            if self.m.shape[1] == 1 or self.m.shape[0] == 1:
                flat = list(self.m)
                sel = flat[k]
                return M(sp.Matrix(sel if isinstance(sel, list) else [sel]))
            raise IndexError("single index on a matrix")
        r, c = k
        rows = list(range(self.m.shape[0]))[r] if isinstance(r, slice) else [r]
        cols = list(range(self.m.shape[1]))[c] if isinstance(c, slice) else [c]
        return M(self.m.extract(rows, cols))

    def __add__(self, o): return _elementwise(self, o, lambda a, b: a + b)
    def __radd__(self, o): return _elementwise(o, self, lambda a, b: a + b)
    def __sub__(self, o): return _elementwise(self, o, lambda a, b: a - b)
    def __rsub__(self, o): return _elementwise(o, self, lambda a, b: a - b)
    def __mul__(self, o): return _elementwise(self, o, lambda a, b: a * b)
    def __rmul__(self, o): return _elementwise(o, self, lambda a, b: a * b)
    def __truediv__(self, o): return _elementwise(self, o, lambda a, b: a / b)
    def __neg__(self): return M(-self.m)


class NumericResult(object):
    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)

    def toarray(self):
        return self.a


class Function(object):
    def __init__(self, name, ins, outs, *rest):
        self.ins, self.outs = [_mat(i) for i in ins], [_mat(o) for o in outs]

    def __call__(self, *args):
        subs = {}
        for formal, actual in zip(self.ins, args):
            if isinstance(actual, M):
                vals = list(actual.m)
            else:
                vals = [sp.Float(float(v), 17) for v in np.asarray(actual, dtype=float).reshape(-1)]
            for s, v in zip(list(formal), vals):
                subs[s] = v
        out = [o.subs(subs) for o in self.outs]
        if all(not e.free_symbols for o in out for e in o):
            return NumericResult(np.array(out[0].evalf(17).tolist(), dtype=float))
        return M(out[0])


class CS(object):
    MX = M
    SX = M
    Function = Function

    @staticmethod
    def vertcat(*args):
        if len(args) == 1 and isinstance(args[0], (list, tuple)):
            args = tuple(args[0])
        return M(sp.Matrix.vstack(*[_mat(a) for a in args]))

    @staticmethod
    def sin(x): return M(_mat(x).applyfunc(sp.sin))
    @staticmethod
    def cos(x): return M(_mat(x).applyfunc(sp.cos))

    @staticmethod
    def mtimes(a, b):
        A, B = _mat(a), _mat(b)
        if A.shape == (1, 1) or B.shape == (1, 1):
            return _elementwise(a, b, lambda x, y: x * y)
        return M(A * B)

    @staticmethod
    def norm_fro(a):
        return M(sp.Matrix([[sp.sqrt(sum(e * e for e in _mat(a)))]]))

    @staticmethod
    def reshape(a, r, c):
        flat = list(_mat(a).T) if _mat(a).shape[1] > 1 else list(_mat(a))     # column-major
        return M(sp.Matrix(r, c, lambda i, j: flat[i + r * j]))


# ---- constraint records ---------------------------------------------------------------------------------

class _Cnstr(object):
    def __init__(self, label, expression, gain=1.0, priority=None, constraint_type="hard", slack_weight=1.0,
                 set_min=None, set_max=None, target=None, **kw):
        self.label, self.expression, self.gain = label, expression, gain
        self.set_min, self.set_max, self.target = set_min, set_max, target


class EqualityConstraint(_Cnstr): pass
class SetConstraint(_Cnstr): pass
class VelocityEqualityConstraint(_Cnstr): pass
class VelocitySetConstraint(_Cnstr): pass


class SkillSpecification(object):
    def __init__(self, label, time_var, robot_var, robot_vel_var=None, virtual_var=None, virtual_vel_var=None,
                 input_var=None, constraints=(), **kw):
        self.label, self.time_var, self.robot_var = label, time_var, robot_var
        self.virtual_var, self.input_var, self.constraints = virtual_var, input_var, list(constraints)


class Mod(object):
    PseudoInverseController = ReactiveQPController = None


for _c in (EqualityConstraint, SetConstraint, VelocityEqualityConstraint, VelocitySetConstraint, SkillSpecification):
    setattr(Mod, _c.__name__, _c)


# ---- forward kinematics re-derived with sympy ---------------------------------------------------------------

def _rot_axis(axis, th):
    a = np.asarray(axis, dtype=float)
    a = a / np.linalg.norm(a)
    K = sp.Matrix([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]]).applyfunc(lambda v: sp.Float(float(v), 17))
    aa = sp.Matrix(3, 3, lambda i, j: sp.Float(float(a[i] * a[j]), 17))
    return sp.cos(th) * sp.eye(3) + sp.sin(th) * K + (1 - sp.cos(th)) * aa


def _homog(R, p):
    T = sp.eye(4)
    T[:3, :3] = R
    T[:3, 3] = sp.Matrix(p)
    return T


def urdf_fk(urdf, root, tip):
    import clik_oracle as orc      # only its URDF parser (numeric origin matrices, axes, limits)
    chain = orc.load_chain(urdf, root, tip)
    n = sum(1 for j in chain if j[0] in ("revolute", "continuous"))
    q = M.sym("qf", n)
    T = sp.eye(4)
    k = 0
    lower, upper = [], []
    for jt, xyz, Ro, axis, lo, hi in chain:
        Rn = sp.Matrix(3, 3, lambda i, j: sp.Float(float(Ro[i, j]), 17))
        T = T * _homog(Rn, [sp.Float(float(v), 17) for v in xyz])
        if jt in ("revolute", "continuous"):
            T = T * _homog(_rot_axis(axis, q.m[k]), [0, 0, 0])
            lower.append(lo)
            upper.append(hi)
            k += 1
    return {"T_fk": Function("T_fk", [q], [M(T)]), "lower": lower, "upper": upper}


def dh_fk(joint_angles, link_lengths, link_offsets, link_twists):
    n = len(link_lengths)
    q = M.sym("qf", n)
    T = sp.eye(4)
    for i in range(n):
        th = q.m[i]
        ca, sa = sp.Float(math.cos(link_twists[i]), 17), sp.Float(math.sin(link_twists[i]), 17)
        Rz = _homog(sp.Matrix([[sp.cos(th), -sp.sin(th), 0], [sp.sin(th), sp.cos(th), 0], [0, 0, 1]]), [0, 0, 0])
        Tz = _homog(sp.eye(3), [0, 0, sp.Float(link_offsets[i], 17)])
        Tx = _homog(sp.eye(3), [sp.Float(link_lengths[i], 17), 0, 0])
        Rx = _homog(sp.Matrix([[1, 0, 0], [0, ca, -sa], [0, sa, ca]]), [0, 0, 0])
        T = T * Rz * Tz * Tx * Rx
    return {"T_fk": Function("T_fk", [q], [M(T)])}


class FK(object):
    ROBOTS = os.path.join(ROOT, "casclik_b200", "fk", "robots")

    def ur5(self):
        return urdf_fk(os.path.join(self.ROBOTS, "ur5_chain.urdf"), "base_link", "tool0")

    def iiwa14(self):
        return urdf_fk(os.path.join(self.ROBOTS, "iiwa14_chain.urdf"), "base_link", "tool0")

    def from_denavit_hartenberg(self, joint_angles, link_lengths, link_offsets, link_twists, **kw):
        return dh_fk(joint_angles, link_lengths, link_offsets, link_twists)


# ---- (i) expression pins ------------------------------------------------------------------------------------

def expression_pins(vectors):
    import golden_skills as gs
    gs._fk = lambda: FK()                       # the catalogue's robots come from the sympy FK above
    ns = gs.Namespace(CS, Mod)
    pins = {}
    done = {}
    for name in sorted(vectors):
        builder = gs.CASES[name][0]
        rec = vectors[name]
        inp = {k: np.array(v, dtype=np.float64) for k, v in rec["inputs"].items()}
        key = (builder.__name__, json.dumps(rec["inputs"], sort_keys=True)[:2000])
        if key in done:                                        # same skill + same inputs under other options
            pins[name] = {"same_as": done[key]}
            continue
        done[key] = name
        spec, _ = builder(ns, np.random.default_rng(0), 8)
        state = list(spec.robot_var.m) + (list(spec.virtual_var.m) if spec.virtual_var is not None else [])
        tsym = spec.time_var.m[0, 0]
        syms = [tsym] + list(spec.robot_var.m)
        if spec.virtual_var is not None:
            syms += list(spec.virtual_var.m)
        if spec.input_var is not None:
            syms += list(spec.input_var.m)
        K = min(N_INST, inp["q"].shape[1])
        cons = []
        for c in spec.constraints:
            e = _mat(c.expression)
            J = e.jacobian(sp.Matrix(state))
            Jt = e.diff(tsym)
            extra = {}
            for field in ("gain", "set_min", "set_max", "target"):
                v = getattr(c, field)
                if isinstance(v, M) and any(x.free_symbols for x in v.m):
                    extra[field] = v.m
            outs = [e, J, Jt] + list(extra.values())
            f = sp.lambdify(syms, outs, modules="mpmath", cse=True)
            rows = []
            import mpmath
            mpmath.mp.dps = 40
            for i in range(K):
                args = [inp["t"][i]] + list(inp["q"][:, i])
                if spec.virtual_var is not None:
                    args += list(inp["x"][:, i])
                if spec.input_var is not None:
                    args += list(inp["y"][:, i])
                vals = f(*[mpmath.mpf(float(a)) for a in args])
                def tolist(mm):
                    return [[float(mm[r, cc]) for cc in range(mm.cols)] for r in range(mm.rows)]
                row = {"e": tolist(vals[0]), "J": tolist(vals[1]), "Jt": tolist(vals[2])}
                for kx, fld in enumerate(extra):
                    row[fld] = tolist(vals[3 + kx])
                rows.append(row)
            cons.append({"label": c.label, "instances": rows})
        pins[name] = {"constraints": cons, "n_instances": K,
                      "state": [str(s) for s in state]}
        print("expression pins:", name, "ok", flush=True)
    return pins


# ---- (ii) second QP solver ----------------------------------------------------------------------------------

def scipy_qp(h, A, lb, ub):
    """Interior-point (trust-constr) solve in the scaled variables z = sqrt(h) x, where the objective is
    1/2 |z|^2 (h spans 1e-3 .. 1: unscaled, SLSQP stalls on most of the fixture's problems)."""
    from scipy.optimize import LinearConstraint
    n = len(h)
    s = 1.0 / np.sqrt(h)
    As = A * s[None, :]
    res = minimize(lambda z: 0.5 * float(z @ z), np.zeros(n), jac=lambda z: z, hess=lambda z: np.eye(n),
                   constraints=[LinearConstraint(As, lb, ub)], method="trust-constr",
                   options={"gtol": 1e-12, "xtol": 1e-14, "barrier_tol": 1e-12, "maxiter": 3000})
    x = res.x * s
    # polish: the interior-point solution identifies the active rows (near a bound); the minimiser on
    # that face is one dense KKT solve (NumPy).  The face is accepted when the polished point is feasible
    # and its multipliers have the right signs (H x + A' lam = 0: lam >= 0 at an upper, <= 0 at a lower
    # bound) — i.e. when it satisfies the KKT conditions of the QP, whose solution is unique.
    r = A @ x
    for rel in (1e-7, 1e-6, 1e-5, 1e-4, 1e-3):
        tol = rel * (1.0 + np.abs(r))
        act_u, act_l = np.abs(r - ub) <= tol, np.abs(r - lb) <= tol
        rows = np.nonzero(act_u | act_l)[0]
        if not len(rows):
            continue
        Aa = A[rows]
        ba = np.where(act_u[rows], ub[rows], lb[rows])
        k = len(rows)
        KKT = np.block([[np.diag(h), Aa.T], [Aa, np.zeros((k, k))]])
        sol, *_ = np.linalg.lstsq(KKT, np.concatenate([np.zeros(n), ba]), rcond=None)
        xp, lam = sol[:n], sol[n:]
        rp = A @ xp
        feas = np.all(rp <= ub + 1e-9 * (1 + np.abs(rp))) and np.all(rp >= lb - 1e-9 * (1 + np.abs(rp)))
        eq = act_u[rows] & act_l[rows]
        signs = np.all((lam >= -1e-9) | ~act_u[rows] | eq) and np.all((lam <= 1e-9) | ~act_l[rows] | eq)
        if feas and signs:
            x = xp
            break
    return x, bool(res.status in (1, 2))


def qp_pins(vectors):
    out = {}
    for name, rec in sorted(vectors.items()):
        if rec["controller"] != "qp":
            continue
        rows = []
        for o in rec["outputs"]:
            h, A = np.array(o["h"]), np.array(o["A"])
            lb, ub = np.array(o["lb"]), np.array(o["ub"])
            # the +-1e10 "no bound" defaults of a SetConstraint are dropped (infinite bounds for scipy)
            lb2 = np.where(np.abs(lb) >= 1e9, -np.inf, lb)
            ub2 = np.where(np.abs(ub) >= 1e9, np.inf, ub)
            x, ok = scipy_qp(h, A, lb2, ub2)
            rows.append({"x": x.tolist(), "success": ok})
        out[name] = rows
        print("scipy QP pins:", name, sum(r["success"] for r in rows), "/", len(rows), flush=True)
    return out


def main():
    with open(os.path.join(HERE, "controller_vectors.json")) as f:
        vectors = json.load(f)
    pins = {"expressions": expression_pins(vectors), "scipy_qp": qp_pins(vectors),
            "how": "tests/golden/make_independent_pins.py: sympy %s (expressions, mpmath 40 digits), scipy trust-constr (QPs)"
                   % sp.__version__}
    with open(os.path.join(HERE, "independent_pins.json"), "w") as f:
        json.dump(pins, f)
    print("wrote independent_pins.json")


if __name__ == "__main__":
    main()
