"""Constraint types of a skill — the user-facing data holders of the controller step.

API mirror of the reference's constraint classes (same class names, constructor signatures,
attribute names, defaults and error types), written against this package's own expression layer:

    EqualityConstraint          reference casclik/constraints.py:88-145
    SetConstraint               reference casclik/constraints.py:148-296  (defaults of +-1e10: :199-206)
    VelocityEqualityConstraint  reference casclik/constraints.py:299-333
    VelocitySetConstraint       reference casclik/constraints.py:336-368

How the controllers read them (the contract the CUDA kernels implement):
    Eq      J v = -gain*e - de/dt            Set     gain*(min-e) <= J v + de/dt <= gain*(max-e)
    VelEq   J v = target - de/dt             VelSet  min <= J v + de/dt <= max
with J = de/d[robot_var; virtual_var].  Attributes may be mutated after construction (the
notebooks change `.priority` and `.expression`), so nothing is cached here.
"""
import numpy as np

from . import sym as cs

_HARD_SOFT = ("hard", "soft")
_BIG = 1e10


def _is_expr(x):
    return isinstance(x, cs.GenericMatrixCommon)


def _n_rows(expression):
    return expression.size()[0]


def _gain_matches(gain, expr_shape, label):
    """Accepted gains: float, square MX/DM (1x1 or rows x rows), square ndarray (rows x rows),
    list of numbers with one entry per row.  Anything else is a TypeError, as in the reference
    (casclik/constraints.py:32-65)."""
    rows = expr_shape[0]
    if isinstance(gain, float):
        return True
    if _is_expr(gain):
        g = gain.size()
        return g[0] == g[1] and g[1] in (rows, 1)
    if isinstance(gain, np.ndarray):
        return gain.ndim == 2 and gain.shape[0] == gain.shape[1] == rows
    if isinstance(gain, list):
        if not all(isinstance(v, (float, int)) for v in gain):
            raise TypeError("Unknown gain type in " + label + ". Supported are: float, MX, DM, "
                            "numpy.ndarray, and list of floats/ints")
        return len(gain) == rows
    raise TypeError("Unknown gain type in " + label + ". Supported are: float, MX, DM, "
                    "numpy.ndarray, and list of floats/ints.")


def _bound_matches(bound, rows, label, which):
    """A set bound is a number (scalar constraints), a constant column MX/DM, or an ndarray with
    one entry per row."""
    if isinstance(bound, (float, int)):
        return rows == 1
    if _is_expr(bound):
        if isinstance(bound, cs.MX) and bound.is_symbolic():
            return False
        s = bound.size()
        return s[0] == rows and s[1] == 1
    if isinstance(bound, np.ndarray):
        if bound.ndim == 1:
            return bound.shape[0] == rows
        return bound.ndim == 2 and bound.shape == (rows, 1)
    raise TypeError("Unknown " + which + " type in " + label + ". Supported are float, MX, DM, "
                    "and numpy.ndarray")


class BaseConstraint(object):
    """label + expression (column vector) + gain; derivative helpers."""

    constraint_class = "BaseConstraint"

    def __init__(self, label, expression, gain):
        self.label = label
        self.expression = expression
        self.gain = gain

    def __repr__(self):
        return "%s<%s at 0x%s>" % (self.label, self.constraint_class, id(self))

    def size(self):
        return self.expression.size()

    def _check_sizes(self):
        shape = self.size()
        if shape[1] != 1:
            return False
        return _gain_matches(self.gain, shape, self.label)

    def jacobian(self, var):
        """d expression / d var."""
        return cs.jacobian(self.expression, var)

    def jtimes(self, varA, varB):
        """(d expression / d varA) * varB."""
        return cs.jtimes(self.expression, varA, varB)

    def nullspace(self, var):
        """I - pinv(J) J with J = d expression / d var."""
        J = self.jacobian(var)
        return cs.MX.eye(var.size()[0]) - cs.mtimes(cs.pinv(J), J)

    # -- used by the controllers ---------------------------------------------------------------
    def gain_times(self, vec):
        """gain * vec with the gain conventions above (list gains act row-wise)."""
        g = self.gain
        if isinstance(g, list):
            g = cs.diag(cs.DM(g))
        return cs.mtimes(g, vec)

    def _merge(self, other, with_bounds):
        if self.priority != other.priority:
            raise TypeError("Added constraints must have same priority.")
        if self.constraint_type != other.constraint_type:
            raise TypeError("Added constrains must have same constraint type")
        na, nb = _n_rows(self.expression), _n_rows(other.expression)

        def block(g, n):
            if isinstance(g, list):
                return cs.diag(cs.DM(g))
            if isinstance(g, (float, int)) or (_is_expr(g) and g.size() == (1, 1)):
                return g * cs.MX.eye(n)
            return g

        gain = cs.MX.zeros(na + nb, na + nb)
        gain[:na, :na] = block(self.gain, na)
        gain[na:, na:] = block(other.gain, nb)
        kw = dict(expression=cs.vertcat(self.expression, other.expression), gain=gain,
                  constraint_type=self.constraint_type, priority=self.priority)
        if with_bounds:
            kw["set_min"] = cs.vertcat(self.set_min, other.set_min)
            kw["set_max"] = cs.vertcat(self.set_max, other.set_max)
        return type(self)(self.label + "+" + other.label, **kw)


class EqualityConstraint(BaseConstraint):
    """Drive `expression` to zero:  J v = -gain*expression - d expression/dt."""

    constraint_class = "EqualityConstraint"

    def __init__(self, label, expression, gain=1.0, constraint_type="hard", priority=1,
                 slack_weight=1.0):
        BaseConstraint.__init__(self, label, expression, gain)
        self.constraint_type = constraint_type
        self.priority = priority
        self.slack_weight = slack_weight
        if not self._check_sizes():
            raise ValueError("Gain and expression dimensions do not match.")

    def __add__(self, other):
        return self._merge(other, with_bounds=False)


class SetConstraint(BaseConstraint):
    """Keep `expression` inside [set_min, set_max] (defaults -1e10 / +1e10 per row)."""

    constraint_class = "SetConstraint"

    def __init__(self, label, expression, gain=1.0, set_min=None, set_max=None,
                 constraint_type="hard", priority=1, slack_weight=1.0):
        BaseConstraint.__init__(self, label, expression, gain)
        self.constraint_type = constraint_type
        self.priority = priority
        rows = _n_rows(expression)
        self.set_min = -_BIG * np.ones(rows) if set_min is None else set_min
        self.set_max = _BIG * np.ones(rows) if set_max is None else set_max
        self.slack_weight = slack_weight
        if not self._check_sizes():
            raise ValueError("Gain, set limits, or expression dimensions do not match in "
                             + self.label)

    def _check_sizes(self):
        rows = self.size()[0]
        ok_gain = BaseConstraint._check_sizes(self)
        ok_min = _bound_matches(self.set_min, rows, self.label, "set_min")
        ok_max = _bound_matches(self.set_max, rows, self.label, "set_max")
        return ok_gain and ok_min and ok_max

    def __add__(self, other):
        return self._merge(other, with_bounds=True)


class VelocityEqualityConstraint(BaseConstraint):
    """Prescribe the rate of `expression`:  J v = target - d expression/dt.
    (No size check, like the reference: casclik/constraints.py:311.)"""

    constraint_class = "VelocityEqualityConstraint"

    def __init__(self, label, expression, gain=1.0, constraint_type="hard", priority=1,
                 target=0.0, slack_weight=1.0):
        BaseConstraint.__init__(self, label, expression, gain)
        self.constraint_type = constraint_type
        self.priority = priority
        self.target = target
        self.slack_weight = slack_weight


class VelocitySetConstraint(BaseConstraint):
    """Bound the rate of `expression`:  set_min <= J v + d expression/dt <= set_max.
    Only the optimisation controllers use it; the pseudo-inverse controller ignores it
    (reference pseudo_inverse.py:278-280)."""

    constraint_class = "VelocitySetConstraint"

    def __init__(self, label, expression, gain=1.0, set_min=-_BIG, set_max=_BIG,
                 constraint_type="hard", priority=1, slack_weight=1.0):
        BaseConstraint.__init__(self, label, expression, gain)
        self.constraint_type = constraint_type
        self.priority = priority
        self.set_min = set_min
        self.set_max = set_max
        self.slack_weight = slack_weight


__all__ = ["BaseConstraint", "EqualityConstraint", "SetConstraint",
           "VelocityEqualityConstraint", "VelocitySetConstraint"]
