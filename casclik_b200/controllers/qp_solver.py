"""`cs.conic`-style solver object backed by the batched CUDA QP kernel (C ABI clik_qp_dense).

Replaces the reference's `cs.conic("solver", "qpoases", {"h": sp, "a": sp}, opts)` object
(reference casclik/controllers/reactive_qp.py:256-260) and its call
`solver(h=H, a=A, lba=Blb, uba=Bub[, x0=...])` (:493, :512-513).  H must be diagonal positive
(it always is in CASCLIK: :175-189); there is no linear term.
"""
import ctypes

import numpy as np

from .. import runtime
from .. import sym as cs


def _arr(x):
    if isinstance(x, cs.GenericMatrixCommon):
        return x.toarray()
    return np.asarray(x, dtype=np.float64)


class ConicSolver(object):
    def __init__(self, name, solver_name, structure, opts):
        self.name, self.solver_name, self.structure, self.opts = name, solver_name, structure, opts

    def solve_dense_batch(self, hdiag, A, lb, ub, x0=None, max_iter=0):
        """hdiag (N, nx), A (N, m, nx), lb/ub (N, m) NumPy -> x (N, nx), status (N,), active (N, 2)."""
        runtime.require_device()
        import torch
        hdiag, A, lb, ub = (np.asarray(v, dtype=np.float64) for v in (hdiag, A, lb, ub))
        N, m, nx = A.shape
        dev = torch.device("cuda", runtime.current_device())

        def up(a):  # instance-major host array -> coordinate-major device tensor
            return torch.from_numpy(np.array(a.reshape(N, -1).T, dtype=np.float64, order="C", copy=True)).to(dev)
        h_d, A_d, lb_d, ub_d = up(hdiag), up(A), up(lb), up(ub)
        x0_d = up(np.asarray(x0, dtype=np.float64)) if x0 is not None else None
        sol = torch.empty((nx, N), dtype=torch.float64, device=dev)
        status = torch.empty((N,), dtype=torch.int32, device=dev)
        active = torch.empty((2, N), dtype=torch.int32, device=dev)
        p = lambda t_: ctypes.c_void_p(t_.data_ptr()) if t_ is not None else None  # noqa: E731
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        runtime.check(runtime.load_library().clik_qp_dense(
            dev.index, N, nx, m, p(h_d), p(A_d), p(lb_d), p(ub_d), p(x0_d), p(sol), p(status),
            p(active), int(max_iter), stream))
        return sol.T.cpu().numpy(), status.cpu().numpy(), active.T.cpu().numpy()

    def __call__(self, h=None, a=None, lba=None, uba=None, x0=None, g=None, lbx=None, ubx=None,
                 **unused):
        H, A = _arr(h), _arr(a)
        if g is not None and np.any(_arr(g) != 0.0):
            raise NotImplementedError("linear cost terms are not supported (CASCLIK never passes g)")
        if lbx is not None or ubx is not None:
            raise NotImplementedError("variable bounds are not supported (CASCLIK never passes them)")
        if np.any(H - np.diag(np.diag(H)) != 0.0):
            raise NotImplementedError("only diagonal H is supported")
        nx = H.shape[0]
        lb = _arr(lba).reshape(1, -1)
        ub = _arr(uba).reshape(1, -1)
        x0a = _arr(x0).reshape(1, nx) if x0 is not None else None
        x, status, active = self.solve_dense_batch(np.diag(H).reshape(1, nx), A.reshape(1, -1, nx),
                                                   lb, ub, x0a)
        if status[0] != runtime.QP_SOLVED:
            raise RuntimeError("conic solver %s: QP %s" % (
                self.name, {runtime.QP_INFEASIBLE: "is infeasible",
                            runtime.QP_INVALID: "has non-finite data or a non-finite solution (is H positive?)"
                            }.get(int(status[0]), "hit the iteration cap")))
        xs = x[0]
        return {"x": cs.DM(xs.reshape(-1, 1)), "cost": cs.DM(0.5 * float(xs @ (np.diag(H) * xs))),
                "active_upper": int(active[0, 0]), "active_lower": int(active[0, 1])}
