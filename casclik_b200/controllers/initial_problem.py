"""Initial-value problem of the reactive QP controller: choose virtual velocities and slack
for a given (fixed) robot velocity before the control loop starts.

Mirrors reference casclik/controllers/reactive_qp.py:300-459:
  variables  [virtual_vel ; slack],  H = diag([mu*w_virt ; (1+mu)*w_slack])           (:324-332)
  rows       only constraints that touch a virtual variable or own slack                 (:383-386)
  bounds     -de/dt - (de/dq) robot_vel + the usual per-type terms                       (:354-370)
It runs once per skill, so the matrices are evaluated on the host from the expression graph and
the QP goes through the dense CUDA QP entry point (clik_qp_dense).
"""
import numpy as np

from .. import sym as cs
from ..codegen.lower import kind_of, KIND_EQ, KIND_SET, KIND_VELEQ, _col
from .qp_solver import ConicSolver


class InitialProblem(object):
    def __init__(self, ctrl):
        spec = ctrl.skill_spec
        self.spec = spec
        nvirt = spec.n_virtual_var if spec.virtual_var is not None else 0
        nslack = spec.n_slack_var
        mu = ctrl.weight_shifter
        self.nvirt, self.nslack = nvirt, nslack
        weights = []
        if nvirt > 0:
            weights.append(mu * ctrl.virtual_var_weights)
        if nslack > 0:
            weights.append((1 + mu) * ctrl.slack_var_weights)
        self.active = False
        if not weights:
            return
        H = cs.diag(cs.vertcat(*weights))
        rows_A, rows_lb, rows_ub = [], [], []
        slack_ind = 0
        found_any = False
        for c in spec.constraints:
            e = c.expression
            rows = e.size()[0]
            kind = kind_of(c)
            found_virt = found_slack = False
            if nvirt > 0:
                Jv = cs.jacobian(e, spec.virtual_var)
                if Jv.nnz() > 0:
                    found_virt = True
                    expr = Jv
                else:
                    expr = cs.DM.zeros(rows, nvirt)
            Jt = cs.jacobian(e, spec.time_var)
            rob_der = c.jtimes(spec.robot_var, spec.robot_vel_var)
            lb = -Jt - rob_der
            ub = -Jt - rob_der
            if kind == KIND_EQ:
                ke = c.gain_times(e)
                lb, ub = lb + (-ke), ub + (-ke)
            elif kind == KIND_SET:
                smin = cs.MX(cs.vertcat(*_col(c.set_min, rows, "set_min", c.label)))
                smax = cs.MX(cs.vertcat(*_col(c.set_max, rows, "set_max", c.label)))
                lb, ub = lb + c.gain_times(smin - e), ub + c.gain_times(smax - e)
            elif kind == KIND_VELEQ:
                tg = cs.MX(cs.vertcat(*_col(c.target, rows, "target", c.label)))
                lb, ub = lb + tg, ub + tg
            else:
                lb = lb + cs.MX(cs.vertcat(*_col(c.set_min, rows, "set_min", c.label)))
                ub = ub + cs.MX(cs.vertcat(*_col(c.set_max, rows, "set_max", c.label)))
            if nslack > 0:
                smat = cs.DM.zeros(rows, nslack)
                if c.constraint_type == "soft":
                    smat[:, slack_ind:slack_ind + rows] = -cs.DM.eye(rows)
                    slack_ind += rows
                    found_slack = True
                expr = cs.horzcat(expr, smat) if nvirt > 0 else smat
            if found_virt or found_slack:
                found_any = True
                rows_A.append(expr)
                rows_lb.append(lb)
                rows_ub.append(ub)
        if not found_any:
            return
        ins = [spec.time_var, spec.robot_var, spec.robot_vel_var]
        names = ["time_var", "robot_var", "robot_vel_var"]
        if spec._has_virtual:
            ins.append(spec.virtual_var)
            names.append("virtual_var")
        if spec._has_input:
            ins.append(spec.input_var)
            names.append("input_var")
        self.funcs = {
            "H": cs.Function("H_initial", ins, [H], names, ["H"]),
            "A": cs.Function("A_initial", ins, [cs.vertcat(*rows_A)], names, ["A"]),
            "Blb": cs.Function("Blb_initial", ins, [cs.vertcat(*rows_lb)], names, ["Blb"]),
            "Bub": cs.Function("Bub_initial", ins, [cs.vertcat(*rows_ub)], names, ["Bub"]),
        }
        self.solver = ConicSolver("solver", ctrl.options["solver_name"], {},
                                  ctrl.options["initial_solver_opts"])
        self.active = True

    def solve(self, t0, q0, x0=None, dq0=None, y0=None):
        spec = self.spec
        if dq0 is None:
            dq0 = [0.0] * spec.n_robot_var
        vals = [t0, q0, dq0]
        if spec._has_virtual:
            vals.append([0.0] * self.nvirt if x0 is None else x0)
        if spec._has_input:
            vals.append([0.0] * spec.n_input_var if y0 is None else y0)
        H, A, lb, ub = (self.funcs[k](*vals) for k in ("H", "A", "Blb", "Bub"))
        res = self.solver(h=H, a=A, lba=lb, uba=ub)
        x = np.asarray(res["x"].toarray()).reshape(-1)
        res_virt = cs.DM(x[:self.nvirt].reshape(-1, 1)) if self.nvirt > 0 else None
        res_slack = (cs.DM(x[self.nvirt:self.nvirt + self.nslack].reshape(-1, 1))
                     if self.nslack > 0 else None)
        return res_virt, res_slack
