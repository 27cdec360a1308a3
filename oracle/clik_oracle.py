"""CPU oracle for the CLIK controller step (NumPy, float64).  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  Nothing under `casclik_b200/` imports it; the product path has no CPU fallback.

PINNING.  The reference ships no tests and no golden `solve()` vectors, and its arithmetic runs
inside CasADi 3.4.1 (requirements.txt:1; `solve` -> Linsol QR, `conic` -> bundled qpOASES) and
urdf2casadi (no pin), neither of which is in /root/reference or installable here.
The oracle is pinned against outputs of the reference's OWN controller code run in the build
container: tests/golden/make_controller_vectors.py imports the unmodified /root/reference/casclik
package with a stand-in `casadi` module (casclik_b200.sym evaluated by NumPy; `conic` -> this file's
QP solver on the reference-built H, A, lba, uba) and drives PseudoInverseController.solve /
ReactiveQPController.solve one instance at a time on 19 skill/option cases ->
tests/golden/controller_vectors.json (modes, velocities, QP matrices and minimisers;
tests/test_golden_controllers.py: modes bit-exact, velocities to ~1e-12).  What that does NOT pin is
CasADi's and qpOASES' own floating-point arithmetic (they would differ from the stand-in at rounding
level for the well-conditioned skills; see DESIGN.md §5 for the ill-conditioned ones): in that sense,
and only in that sense, parity with a CasADi-backed run of the reference remains UNPINNED.
Also pinned, in tests/test_oracle_kat.py:
  * FK: |p(UR5_home)| = 1.0192 and the home dual quaternion printed in the notebooks;
  * the mode tables produced by running the reference's own create_activation_map
    (pseudo_inverse.py:107-130) with a stub casadi module (tests/golden/activation_maps.json);
  * hand-derived known answers of the in-tree formulas (cart KATs P1-P3, Q1-Q2).

Every function below restates one block of the reference, cited as file:line under
/root/reference/casclik/.  The restatement is literal on purpose (explicit pseudo-inverse
matrices, explicit null-space projector, the first-equality double application), so that it is
the *reference's* floating-point formula that the CUDA kernels are compared with, not a tidied one.
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

EQ, SET, VELEQ, VELSET = "eq", "set", "veleq", "velset"


# ----------------------------------------------------------------------------------------------
# numeric constraint blocks: what the compiled CasADi functions evaluate at one (t, q, x, y)
# ----------------------------------------------------------------------------------------------

@dataclass
class Block:
    """Numeric value of one constraint for a batch of N instances (priority-sorted position is the
    position in the list handed to the step functions)."""
    kind: str                      # EQ | SET | VELEQ | VELSET
    e: np.ndarray                  # (N, m)      expression
    J: np.ndarray                  # (N, m, n)   d e / d [robot_var; virtual_var]
    Jt: np.ndarray                 # (N, m)      d e / d time_var
    gain: object = 1.0             # float or (m, m) / (N, m, m) matrix
    set_min: Optional[np.ndarray] = None   # (m,) or (N, m)
    set_max: Optional[np.ndarray] = None
    target: object = 0.0           # VELEQ: float, (m,) or (N, m)
    soft: bool = False
    slack_weight: float = 1.0

    @property
    def rows(self):
        return self.e.shape[1]


def _gain_times(gain, vec):
    """cs.mtimes(gain, vec): scalar * vector or matrix @ vector (constraints.py:36-51)."""
    g = np.asarray(gain, dtype=vec.dtype)
    if g.ndim == 0 or g.size == 1:
        return g.reshape(-1)[0] * vec
    if g.ndim == 2:
        return np.einsum("ij,nj->ni", g, vec)
    return np.einsum("nij,nj->ni", g, vec)


def _bcast(val, N, m):
    a = np.asarray(val)
    if a.dtype not in (np.float64, np.longdouble):
        a = a.astype(np.float64)
    if a.ndim == 0:
        a = np.full((m,), float(a))
    if a.ndim == 2 and a.shape[1] == 1 and a.shape[0] == m:
        a = a.reshape(-1)
    return np.broadcast_to(a, (N, m))


# ----------------------------------------------------------------------------------------------
# pseudo_inverse.py:107-130  create_activation_map
# ----------------------------------------------------------------------------------------------

def activation_map(n_sets: int) -> List[List[int]]:
    """2^S bit lists, entry k = 1 iff the k-th SetConstraint (priority order) is active; stable
    sort by number of active sets.  Set 0 is the least significant bit of the unsorted index."""
    if n_sets == 0:
        return []
    rows = []
    for idx in range(2 ** n_sets):
        rows.append([(idx >> k) & 1 for k in range(n_sets)])
    return sorted(rows, key=lambda r: sum(r))


# ----------------------------------------------------------------------------------------------
# pseudo_inverse.py:92-105  pinv
# ----------------------------------------------------------------------------------------------

def _solve(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Batched A X = B.  float64: LAPACK dgesv (partial pivoting).  Any other dtype (the tests use
    np.longdouble as a higher-precision referee): the same algorithm written out in NumPy."""
    if A.dtype == np.float64:
        return np.linalg.solve(A, B)
    A = A.copy()
    B = B.copy()
    N, k, _ = A.shape
    ar = np.arange(N)
    for c in range(k):
        piv = c + np.argmax(np.abs(A[:, c:, c]), axis=1)
        rows_c, rows_p = A[ar, c].copy(), A[ar, piv].copy()
        A[ar, c], A[ar, piv] = rows_p, rows_c
        rows_c, rows_p = B[ar, c].copy(), B[ar, piv].copy()
        B[ar, c], B[ar, piv] = rows_p, rows_c
        f = A[:, c + 1:, c] / A[:, c, c][:, None]
        A[:, c + 1:, :] -= f[:, :, None] * A[:, c, :][:, None, :]
        B[:, c + 1:, :] -= f[:, :, None] * B[:, c, :][:, None, :]
    X = np.zeros_like(B)
    for r in range(k - 1, -1, -1):
        acc = B[:, r, :] - np.einsum("nk,nkj->nj", A[:, r, r + 1:], X[:, r + 1:, :])
        X[:, r, :] = acc / A[:, r, r][:, None]
    return X


def damped_pinv(J: np.ndarray, method: str = "damped", damping: float = 1e-7) -> np.ndarray:
    """J: (N, m, n) -> (N, n, m).
    damped:   cols >= rows: (solve(J J' + lam I, J))'   else  solve(J' J + lam I, J')
    standard: cs.pinv -> the same two branches without damping (CasADi's generic pinv takes the
              tall branch for square matrices)."""
    m, n = J.shape[-2:]
    Jt = np.swapaxes(J, -1, -2)
    if method == "damped":
        wide = n >= m
        lam = damping
    elif method == "standard":
        wide = not (m >= n)
        lam = 0.0
    else:
        raise ValueError(method)
    if wide:
        inner = J @ Jt + J.dtype.type(lam) * np.eye(m, dtype=J.dtype)
        return np.swapaxes(_solve(inner, J), -1, -2)
    inner = Jt @ J + J.dtype.type(lam) * np.eye(n, dtype=J.dtype)
    return _solve(inner, Jt)


# ----------------------------------------------------------------------------------------------
# pseudo_inverse.py:132-190  get_in_tangent_cone_function (scalar sets)
# ----------------------------------------------------------------------------------------------

def in_tangent_cone(e, de, set_min, set_max):
    """in_tc = (min - e < 1e-12) ? ((e - max < 1e-12) ? 1 : (de < 0)) : (de > 0); all (N,)."""
    leq_high = np.where(e - set_max < 1e-12, True, de < 0.0)
    return np.where(set_min - e < 1e-12, leq_high, de > 0.0)


# ----------------------------------------------------------------------------------------------
# pseudo_inverse.py:259-451 (one mode) and :512-556 (mode search)
# ----------------------------------------------------------------------------------------------

def in_tangent_cone_multidim(e, de, set_min, set_max):
    """pseudo_inverse.py:222-252 for a vector-valued set; e, de, bounds are (N, m).
    inside <=> every e - min >= 1e-12 and every e - max <= 1e-12; otherwise the motion must point
    inwards, with a 45-degree rule when every component is outside (a "corner")."""
    le = e - set_min
    ue = e - set_max
    le_good = le >= 1e-12
    ue_good = ue <= 1e-12
    inside = le_good.all(axis=1) & ue_good.all(axis=1)
    out_dir = (np.sign(le) + np.sign(ue)) / 2.0
    corner = (np.sign(le) == np.sign(ue)).all(axis=1)
    proj = np.einsum("nj,nj->n", out_dir, de)
    dists = (np.sqrt(np.einsum("nj,nj->n", de, de)) + 1e-10) * np.sqrt(np.einsum("nj,nj->n", out_dir, out_dir))
    with np.errstate(all="ignore"):
        corner_handler = np.where(proj < 0.0, np.abs(-proj) / dists < math.cos(math.pi / 4), False)
    going_in = np.where(corner, corner_handler, proj < 0.0)
    return np.where(inside, True, going_in)


def _mode_velocity(blocks: Sequence[Block], active_bits: Sequence[int], n: int, N: int, opts):
    """Velocity of one mode for all N instances + the list of inactive scalar sets to test."""
    ff = opts.get("feedforward", True)
    method = opts.get("pinv_method", "damped")
    lam = opts.get("damping_factor", 1e-7)
    conv_last = opts.get("converge_final_set_to_max", False)
    multidim = opts.get("multidim_sets", False)

    dt = blocks[0].J.dtype
    v = np.zeros((N, n), dtype=dt)
    Jlist, rJlist, to_test = [], [], []
    set_idx = 0
    eye = np.eye(n, dtype=dt)

    def nullspace_term(Ji, des):
        J0 = np.concatenate(Jlist, axis=1)
        rJ0 = np.concatenate(rJlist, axis=1)
        N0 = eye - damped_pinv(J0, method, lam) @ rJ0                # :389-392
        NJ = N0 @ damped_pinv(Ji, method, lam)                         # :393
        return np.einsum("nij,nj->ni", NJ, des)                        # :394

    for b in blocks:
        is_first = len(Jlist) == 0                                     # :276
        if b.kind == SET and b.rows > 1 and not multidim:
            raise NotImplementedError("multi-row SetConstraint needs multidim_sets (:299-312)")
        if b.kind == EQ:
            des = -_gain_times(b.gain, b.e)                            # :318 / :383
            if ff:
                des = des - b.Jt                                       # :320-321
            if is_first:                                               # :317-326
                v = v + np.einsum("nij,nj->ni", damped_pinv(b.J, method, lam), des)
                Jlist.append(b.J)
                rJlist.append(b.J)
            # the chain that starts at :327 is a *new* if: a first EqualityConstraint also
            # runs the generic branch :382-396
            v = v + nullspace_term(b.J, des)
            Jlist.append(b.J)
            rJlist.append(b.J)
        elif b.kind == VELEQ:
            des = _bcast(b.target, N, b.rows)                          # :328 / :431
            if ff:
                des = des - b.Jt
            if is_first:                                               # :327-335
                v = v + np.einsum("nij,nj->ni", damped_pinv(b.J, method, lam), des)
            else:                                                      # :430-443
                v = v + nullspace_term(b.J, des)
            Jlist.append(b.J)
            rJlist.append(b.J)
        elif b.kind == SET:
            if active_bits[set_idx] and conv_last and b is blocks[-1]:  # :337-356
                des = _gain_times(b.gain, _bcast(b.set_max, N, b.rows) - b.e)
                if ff:
                    des = des - b.Jt
                if not Jlist:
                    raise ValueError("converge_final_set_to_max with an empty active list: the "
                                     "reference fails at setup (cs.vertcat(*[]), Appendix A16)")
                v = v + nullspace_term(b.J, des)
            if active_bits[set_idx]:                                   # :399-405
                Jlist.append(b.J)
                if multidim:                                           # :289-298, :401-402
                    hi = _bcast(b.set_max, N, b.rows)
                    lo = _bcast(b.set_min, N, b.rows)
                    act = ((b.e - hi > 0.0) | (b.e - lo < 0.0)).astype(b.J.dtype)
                    rJlist.append(act[:, :, None] * b.J)               # S @ Ji, S = diag(active)
                else:
                    rJlist.append(b.J)
            else:                                                      # :406-415
                to_test.append(b)
            set_idx += 1
        # VELSET: no branch matches -> ignored (Appendix A6)
    return v, to_test


def as_dtype(blocks: Sequence[Block], dtype):
    """Copy of the numeric blocks in another floating type (np.longdouble = referee precision)."""
    def c(a):
        return None if a is None else np.asarray(a, dtype=dtype)
    return [Block(b.kind, c(b.e), c(b.J), c(b.Jt), c(b.gain), c(b.set_min), c(b.set_max),
                  c(b.target), b.soft, b.slack_weight) for b in blocks]


def pinv_step(blocks: Sequence[Block], n_state: int, options: Optional[dict] = None):
    """PseudoInverseController.solve for N instances.  Returns (v (N, n_state), mode (N,) int32);
    mode = -1 and v = 0 where no mode is admissible (pseudo_inverse.py:551-555).
    Works in the dtype of the blocks (float64 = the reference's arithmetic)."""
    opts = dict(options or {})
    N = blocks[0].e.shape[0]
    n_sets = sum(1 for b in blocks if b.kind == SET)
    amap = activation_map(n_sets) or [[]]
    v_out = np.zeros((N, n_state), dtype=blocks[0].J.dtype)
    mode_out = np.full((N,), -1, dtype=np.int32)
    todo = np.ones((N,), dtype=bool)
    for mode_idx, bits in enumerate(amap):
        if not todo.any():
            break
        idx = np.nonzero(todo)[0]
        sub = [_take(b, idx) for b in blocks]
        v, to_test = _mode_velocity(sub, bits, n_state, len(idx), opts)
        ok = np.ones((len(idx),), dtype=bool)
        for b in to_test:
            if b.rows == 1:
                de = b.Jt[:, 0] + np.einsum("nj,nj->n", b.J[:, 0, :], v)  # :151-158
                lo = _bcast(b.set_min, len(idx), 1)[:, 0]
                hi = _bcast(b.set_max, len(idx), 1)[:, 0]
                ok &= in_tangent_cone(b.e[:, 0], de, lo, hi).astype(bool)
            else:                                                      # :211-218, multidim_sets
                de = b.Jt + np.einsum("nij,nj->ni", b.J, v)
                lo = _bcast(b.set_min, len(idx), b.rows)
                hi = _bcast(b.set_max, len(idx), b.rows)
                ok &= in_tangent_cone_multidim(b.e, de, lo, hi).astype(bool)
        acc = idx[ok]
        v_out[acc] = v[ok]
        mode_out[acc] = mode_idx
        todo[acc] = False
    return v_out, mode_out


def _take(b: Block, idx):
    def t(a, full_ndim):
        if a is None:
            return None
        a = np.asarray(a)
        return a[idx] if a.ndim == full_ndim else a
    g = b.gain
    if isinstance(g, np.ndarray) and g.ndim == 3:
        g = g[idx]
    return Block(b.kind, b.e[idx], b.J[idx], b.Jt[idx], g, t(b.set_min, 2), t(b.set_max, 2),
                 t(b.target, 2) if isinstance(b.target, np.ndarray) else b.target,
                 b.soft, b.slack_weight)


# ----------------------------------------------------------------------------------------------
# reactive_qp.py:175-189, :191-246  problem matrices
# ----------------------------------------------------------------------------------------------

MU = 0.001  # reactive_qp.py:44 weight_shifter


def qp_matrices(blocks: Sequence[Block], n_rob: int, n_virt: int, w_rob=None, w_virt=None,
                w_slack=None):
    """-> (hdiag (nx,), A (N, m, nx), lb (N, m), ub (N, m)) with x = [robot vel; virtual vel; slack]."""
    N = blocks[0].e.shape[0]
    n_slack = sum(b.rows for b in blocks if b.soft)
    w_rob = np.ones(n_rob) if w_rob is None else np.asarray(w_rob, float)
    w_virt = np.ones(n_virt) if w_virt is None else np.asarray(w_virt, float)
    if w_slack is None:
        w_slack = np.concatenate([b.slack_weight * np.ones(b.rows) for b in blocks if b.soft]
                                 or [np.zeros(0)])
    h = [MU * w_rob]
    if n_virt > 0:
        h.append(MU * w_virt)
    if n_slack > 0:
        h.append(MU + np.asarray(w_slack, float))                      # :187
    hdiag = np.concatenate(h)
    rows_A, rows_lb, rows_ub = [], [], []
    slack_ind = 0
    for b in blocks:
        A = b.J                                                        # :210-213
        lb = -b.Jt                                                     # :215-216
        ub = -b.Jt
        if b.kind == EQ:
            ke = _gain_times(b.gain, b.e)
            lb = lb - ke
            ub = ub - ke
        elif b.kind == SET:
            lb = lb + _gain_times(b.gain, _bcast(b.set_min, N, b.rows) - b.e)
            ub = ub + _gain_times(b.gain, _bcast(b.set_max, N, b.rows) - b.e)
        elif b.kind == VELEQ:
            tg = _bcast(b.target, N, b.rows)
            lb = lb + tg
            ub = ub + tg
        elif b.kind == VELSET:
            lb = lb + _bcast(b.set_min, N, b.rows)
            ub = ub + _bcast(b.set_max, N, b.rows)
        if n_slack > 0:                                                # :233-238
            S = np.zeros((N, b.rows, n_slack))
            if b.soft:
                S[:, np.arange(b.rows), slack_ind + np.arange(b.rows)] = -1.0
                slack_ind += b.rows
            A = np.concatenate([A, S], axis=2)
        rows_A.append(A)
        rows_lb.append(lb)
        rows_ub.append(ub)
    return (hdiag, np.concatenate(rows_A, axis=1), np.concatenate(rows_lb, axis=1),
            np.concatenate(rows_ub, axis=1))


# ----------------------------------------------------------------------------------------------
# the QP itself.  The reference hands (H, A, lb, ub) to qpOASES through cs.conic
# (reactive_qp.py:256-260, :493).  H is diagonal positive, so the minimiser is unique; the oracle
# finds it with a dual active-set iteration written for clarity (dense least squares at every
# step), and `kkt_residuals` certifies optimality independently of any solver.
# ----------------------------------------------------------------------------------------------

def solve_qp_single(hdiag, A, lb, ub, max_iter=200, tol=1e-12):
    """min 1/2 x' diag(h) x  s.t. lb <= A x <= ub.
    Returns (x, lam, status) with lam_i > 0 at an active upper bound, < 0 at an active lower
    bound (the sign convention of CasADi's lam_a); status 0 ok, 1 iteration cap, 2 infeasible."""
    m, n = A.shape
    s = 1.0 / np.sqrt(hdiag)
    At = A * s[None, :]                      # rows in z = sqrt(H) x coordinates
    z = np.zeros(n)
    W: List[int] = []                        # working set (row indices)
    sg: List[float] = []                     # +1 upper, -1 lower
    u = np.zeros(0)
    # a row counts as violated beyond 1e-12 * max(1, |its own bound|): the default +-1e10 bounds of a
    # SetConstraint (constraints.py:199-206) must not loosen the test for the other rows
    tol_u = tol * np.maximum(1.0, np.abs(np.where(np.isfinite(ub), ub, 1.0)))
    tol_l = tol * np.maximum(1.0, np.abs(np.where(np.isfinite(lb), lb, 1.0)))
    for _ in range(max_iter):
        r = At @ z
        viol_u = np.where(r - ub > tol_u, r - ub, -np.inf)
        viol_l = np.where(lb - r > tol_l, lb - r, -np.inf)
        viol_u[W] = -np.inf
        viol_l[W] = -np.inf
        iu, il = int(np.argmax(viol_u)), int(np.argmax(viol_l))
        if viol_u[iu] >= viol_l[il]:
            p, sp, vp = iu, 1.0, viol_u[iu]
        else:
            p, sp, vp = il, -1.0, viol_l[il]
        if not np.isfinite(vp):
            lam = np.zeros(m)
            for j, (w, g) in enumerate(zip(W, sg)):
                lam[w] = g * u[j]
            return z * s, lam, 0
        npv = sp * At[p]
        up = 0.0
        while True:
            if W:
                Nm = (np.array(sg)[:, None] * At[W]).T          # n x k
                rr, *_ = np.linalg.lstsq(Nm, npv, rcond=None)
                d = npv - Nm @ rr
            else:
                rr = np.zeros(0)
                d = npv.copy()
            dn = d @ npv
            t1, k = math.inf, -1
            for j in range(len(W)):
                if rr[j] > 1e-14 and u[j] / rr[j] < t1:
                    t1, k = u[j] / rr[j], j
            slack_p = sp * (At[p] @ z) - (ub[p] if sp > 0 else -lb[p])   # > 0 : violated
            if dn > 1e-14 * max(1.0, npv @ npv):
                t2 = slack_p / dn
            else:
                t2 = math.inf
            t = min(t1, t2)
            if not math.isfinite(t):
                return z * s, np.zeros(m), 2
            if math.isfinite(t2):
                z = z - t * d
            u = u - t * rr
            up += t
            if t == t2:
                W.append(p)
                sg.append(sp)
                u = np.append(u, up)
                break
            W.pop(k)
            sg.pop(k)
            u = np.delete(u, k)
    return z * s, np.zeros(m), 1


def solve_qp(hdiag, A, lb, ub):
    """Batched wrapper: A (N, m, nx), lb/ub (N, m).  -> x (N, nx), lam (N, m), status (N,)."""
    N, m, n = A.shape
    X = np.zeros((N, n))
    L = np.zeros((N, m))
    S = np.zeros((N,), dtype=np.int32)
    for i in range(N):
        X[i], L[i], S[i] = solve_qp_single(hdiag, A[i], lb[i], ub[i])
    return X, L, S


def kkt_residuals(hdiag, A, lb, ub, x, lam=None):
    """Solver-independent optimality certificate for one instance.  If lam is None the
    multipliers are recovered by least squares on the rows active at x.
    -> dict(primal, stationarity, sign, objective)"""
    r = A @ x
    primal = max(0.0, float(np.max(lb - r)), float(np.max(r - ub)))
    g = hdiag * x
    if lam is None:
        tol = 1e-7 * (1.0 + np.abs(r))
        act = np.nonzero((np.abs(r - lb) <= tol) | (np.abs(r - ub) <= tol))[0]
        lam = np.zeros(A.shape[0])
        if len(act):
            sol, *_ = np.linalg.lstsq(A[act].T, -g, rcond=None)
            lam[act] = sol
    stat = float(np.max(np.abs(g + A.T @ lam))) if A.size else float(np.max(np.abs(g)))
    tolb = 1e-7 * (1.0 + np.abs(r))
    at_u = np.abs(r - ub) <= tolb
    at_l = np.abs(r - lb) <= tolb
    bad = 0.0
    for i in range(A.shape[0]):
        if at_u[i] and at_l[i]:
            continue                       # equality row: any sign
        if at_u[i]:
            bad = max(bad, -lam[i])
        elif at_l[i]:
            bad = max(bad, lam[i])
        else:
            bad = max(bad, abs(lam[i]))
    return {"primal": primal, "stationarity": stat, "sign": float(bad),
            "objective": 0.5 * float(x @ (hdiag * x))}


def qp_step(blocks: Sequence[Block], n_rob: int, n_virt: int, **weights):
    """ReactiveQPController.solve for N instances -> (x (N, nx), lam (N, m), status (N,))."""
    hdiag, A, lb, ub = qp_matrices(blocks, n_rob, n_virt, **weights)
    return solve_qp(hdiag, A, lb, ub)


# ----------------------------------------------------------------------------------------------
# independent forward kinematics (numeric, geometric Jacobian) for the fixture robots.
# Conventions of SURVEY.md Appendix B (urdf2casadi): T = prod T_origin(xyz, rpy) Rot(axis, q_i),
# rpy fixed-axis XYZ.  Deliberately does not import casclik_b200.
# ----------------------------------------------------------------------------------------------

def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def load_chain(urdf_file: str, root: str, tip: str):
    """-> list of (type, xyz, R_origin, axis, lower, upper) from root to tip."""
    tree = ET.parse(urdf_file).getroot()
    by_child = {}
    for j in tree.findall("joint"):
        o = j.find("origin")
        xyz = [float(v) for v in (o.get("xyz", "0 0 0") if o is not None else "0 0 0").split()]
        rpy = [float(v) for v in (o.get("rpy", "0 0 0") if o is not None else "0 0 0").split()]
        a = j.find("axis")
        axis = [float(v) for v in (a.get("xyz") if a is not None else "1 0 0").split()]
        lim = j.find("limit")
        lo = float(lim.get("lower")) if lim is not None and lim.get("lower") else -math.inf
        hi = float(lim.get("upper")) if lim is not None and lim.get("upper") else math.inf
        by_child[j.find("child").get("link")] = (j.get("type"), np.array(xyz), _rpy(*rpy),
                                                 np.array(axis), lo, hi,
                                                 j.find("parent").get("link"))
    path = []
    link = tip
    while link != root:
        rec = by_child[link]
        path.append(rec[:6])
        link = rec[6]
    return list(reversed(path))


def _rot_axis(axis, th):
    """Rodrigues, batched over th (N,) -> (N, 3, 3)."""
    a = axis / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    c, s = np.cos(th)[:, None, None], np.sin(th)[:, None, None]
    return c * np.eye(3) + s * K + (1 - c) * np.outer(a, a)


def fk_pose(chain, q):
    """q (N, n) -> R (N,3,3), p (N,3), plus per-joint world axes and origins for the Jacobian."""
    N = q.shape[0]
    R = np.broadcast_to(np.eye(3), (N, 3, 3)).copy()
    p = np.zeros((N, 3))
    axes, origins = [], []
    k = 0
    for (jt, xyz, Ro, axis, _, _) in chain:
        p = p + np.einsum("nij,j->ni", R, xyz)
        R = R @ Ro
        if jt in ("revolute", "continuous"):
            axes.append(np.einsum("nij,j->ni", R, axis / np.linalg.norm(axis)))
            origins.append(p.copy())
            R = R @ _rot_axis(axis, q[:, k])
            k += 1
        elif jt != "fixed":
            raise NotImplementedError(jt)
    return R, p, axes, origins


def position_jacobian(chain, q):
    """Geometric Jacobian of the tip position: column i = z_i x (p - o_i).  -> p (N,3), J (N,3,n)."""
    _, p, axes, origins = fk_pose(chain, q)
    cols = [np.cross(z, p - o) for z, o in zip(axes, origins)]
    return p, np.stack(cols, axis=2)


def rotation_jacobian(chain, q):
    """d vec(R) / d q_i = [z_i]x R.  -> R (N,3,3), dR (N, n, 3, 3)."""
    R, _, axes, _ = fk_pose(chain, q)
    dR = []
    for z in axes:
        K = np.zeros((q.shape[0], 3, 3))
        K[:, 0, 1], K[:, 0, 2] = -z[:, 2], z[:, 1]
        K[:, 1, 0], K[:, 1, 2] = z[:, 2], -z[:, 0]
        K[:, 2, 0], K[:, 2, 1] = -z[:, 1], z[:, 0]
        dR.append(K @ R)
    return R, np.stack(dR, axis=1)
