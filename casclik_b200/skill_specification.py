"""SkillSpecification — the container a controller is built from.

API mirror of reference casclik/skill_specification.py:23-250: same constructor, properties,
`count_constraints()` keys and `print_constraints()` text.  What the controllers rely on
 and what is therefore part of the contract:

  * `constraints` is stably sorted by `priority` when assigned (reference :139-142);
  * `n_slack_var` = total rows of the "soft" constraints, in that sorted order, and `slack_var`
    is a fresh symbol vector of that length or None (reference :143-151);
  * `_has_virtual` / `_has_input` are structural: does any constraint expression, gain, target
    or set bound depend on the virtual / input symbols (reference :154-200).
"""
import sys

from . import sym as cs
from .constraints import (EqualityConstraint, SetConstraint, VelocityEqualityConstraint,
                          VelocitySetConstraint)


def _vector_rows(var):
    return 0 if var is None else var.size()[0]


class SkillSpecification(object):
    """A labelled, priority-sorted list of constraints over (time, robot, virtual, input) symbols.

    Args:
        label (str): name of the skill
        time_var (cs.MX.sym): time symbol
        robot_var (cs.MX.sym): controllable robot coordinates
        robot_vel_var (cs.MX.sym): their rates (created when None)
        virtual_var (cs.MX.sym): internal virtual coordinates (optional)
        virtual_vel_var (cs.MX.sym): their rates (created when None)
        input_var (cs.MX.sym): external inputs; never differentiated for control
        constraints (list): constraint objects
    """

    def __init__(self, label, time_var, robot_var, robot_vel_var=None, virtual_var=None,
                 virtual_vel_var=None, input_var=None, constraints=()):
        self._constraints = []
        self._virtual_var = None
        self._input_var = None
        self.n_virtual_var = 0
        self.n_input_var = 0
        self.label = label
        self.time_var = time_var
        self.robot_var = robot_var
        self.robot_vel_var = robot_vel_var
        self.virtual_var = virtual_var
        self.virtual_vel_var = virtual_vel_var
        self.input_var = input_var
        self.constraints = constraints

    # -- symbols -----------------------------------------------------------------------------------
    @property
    def robot_var(self):
        return self._robot_var

    @robot_var.setter
    def robot_var(self, var):
        self._robot_var = var
        self.n_robot_var = _vector_rows(var)

    @staticmethod
    def _checked_rate(var, of, what, of_what):
        if not isinstance(var, cs.MX):
            raise TypeError(what + " must be cs.MX.sym.")
        if var.size() != of.size():
            raise ValueError(of_what + " and " + what + " must have the same dimensions")
        return var

    @property
    def robot_vel_var(self):
        return self._robot_vel_var

    @robot_vel_var.setter
    def robot_vel_var(self, var):
        if var is None:
            var = cs.MX.sym("robot_vel_var", self.n_robot_var)
        else:
            var = self._checked_rate(var, self.robot_var, "robot_vel_var", "robot_var")
        self._robot_vel_var = var

    @property
    def virtual_var(self):
        return self._virtual_var

    @virtual_var.setter
    def virtual_var(self, var):
        self._virtual_var = var
        self.n_virtual_var = _vector_rows(var)
        self._check_var_existence()

    @property
    def virtual_vel_var(self):
        return self._virtual_vel_var

    @virtual_vel_var.setter
    def virtual_vel_var(self, var):
        if var is None:
            var = cs.MX.sym("virtual_vel_var", self.n_virtual_var)
        else:
            var = self._checked_rate(var, self.virtual_var, "virtual_vel_var", "virtual_var")
        self._virtual_vel_var = var

    @property
    def input_var(self):
        return self._input_var

    @input_var.setter
    def input_var(self, var):
        self._input_var = var
        self.n_input_var = _vector_rows(var)
        self._check_var_existence()

    # -- constraints -----------------------------------------------------------------------------
    @property
    def constraints(self):
        return self._constraints

    @constraints.setter
    def constraints(self, cnstr_list):
        # sorted() is stable: equal priorities keep insertion order
        self._constraints = sorted(cnstr_list, key=lambda c: c.priority)
        self.n_slack_var = sum(c.expression.size()[0] for c in self._constraints
                               if c.constraint_type == "soft")
        self.slack_var = cs.MX.sym("slack_var", self.n_slack_var) if self.n_slack_var else None
        self._check_var_existence()

    def _depends(self, var):
        """Does anything the controllers evaluate depend on `var`?"""
        if var is None:
            return False
        for c in self._constraints:
            if cs.jacobian(c.expression, var).nnz() > 0:
                return True
            for attr in ("target", "set_min", "set_max", "gain"):
                val = getattr(c, attr, None)
                if isinstance(val, cs.MX) and cs.jacobian(val, var).nnz() > 0:
                    return True
        return False

    def _check_var_existence(self):
        self._has_virtual = self._depends(self._virtual_var)
        self._has_input = self._depends(self._input_var)

    # -- reporting -------------------------------------------------------------------------------
    def count_constraints(self):
        """dict with keys all, equality, velocity_equality, set, velocity_set, hard, soft."""
        kinds = (("equality", EqualityConstraint), ("set", SetConstraint),
                 ("velocity_equality", VelocityEqualityConstraint),
                 ("velocity_set", VelocitySetConstraint))
        out = {"all": len(self._constraints), "hard": 0, "soft": 0}
        out.update({k: 0 for k, _ in kinds})
        for c in self._constraints:
            if c.constraint_type in ("hard", "soft"):
                out[c.constraint_type] += 1
            for key, cls in kinds:
                if isinstance(c, cls):
                    out[key] += 1
                    break
        return out

    def print_constraints(self):
        n = self.count_constraints()
        lines = ["SkillSpecification: " + self.label]
        lines += ["#%d: %s" % (i, c.label) for i, c in enumerate(self._constraints)]
        lines += ["Has virtual var: " + str(self._has_virtual),
                  "Has input var: " + str(self._has_input),
                  "N constraints: " + str(n["all"]),
                  "N equality:",
                  "\tPos:%d\tVel:%d" % (n["equality"], n["velocity_equality"]),
                  "N set:",
                  "\tPos:%d\tVel:%d" % (n["set"], n["velocity_set"])]
        sys.stdout.write("\n".join(lines) + "\n")
        sys.stdout.flush()
