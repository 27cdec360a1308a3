"""Scalar expression DAG: hash-consed nodes, algebraic simplification, forward-mode AD,
vectorised NumPy evaluation.

This is the bottom layer of the expression compiler.  It plays the role CasADi's SX graph plays
under the reference (the reference builds MX graphs in casclik/constraints.py:67-80 and evaluates
them through CasADi's generated C — SURVEY.md §2.2); nothing here is derived from CasADi's code.

Design points
  * every node is interned (hash-consing) so structurally equal sub-expressions are one node:
    common sub-expression elimination is a property of construction, not a later pass;
  * node ids grow with creation, so ascending id is a topological order;
  * constructors fold constants and apply the cheap identities (x+0, x*1, x*0, --x, ...), which
    is what keeps forward-mode derivatives of kinematic chains small;
  * `forward()` propagates one tangent direction through the reachable sub-DAG (Jacobian columns
    are one pass per symbol; `jtimes` is one pass in total).
"""
from __future__ import annotations

import math
import struct
from typing import Dict, Iterable, List, Sequence

import numpy as np

# ----------------------------------------------------------------------------------------------
# node
# ----------------------------------------------------------------------------------------------

#: operator name -> arity
OPS = {
    "const": 0, "sym": 0,
    "add": 2, "sub": 2, "mul": 2, "div": 2, "neg": 1,
    "sin": 1, "cos": 1, "tan": 1, "asin": 1, "acos": 1, "atan": 1, "atan2": 2,
    "exp": 1, "log": 1, "sqrt": 1, "pow": 2, "fabs": 1, "sign": 1,
    "floor": 1, "ceil": 1, "fmin": 2, "fmax": 2,
    "lt": 2, "le": 2, "eq": 2, "ne": 2, "and": 2, "or": 2, "not": 1,
    "if_else": 3,
}

_COMMUTATIVE = {"add", "mul", "fmin", "fmax", "eq", "ne", "and", "or"}


class Node:
    """One scalar operation.  Immutable; compare with `is`."""

    __slots__ = ("op", "args", "val", "name", "id")

    def __init__(self, op, args, val, name, nid):
        self.op = op
        self.args = args
        self.val = val
        self.name = name
        self.id = nid

    def __repr__(self):
        if self.op == "const":
            return repr(self.val)
        if self.op == "sym":
            return self.name
        return "%s(%s)" % (self.op, ", ".join(repr(a) for a in self.args))

    @property
    def is_const(self):
        return self.op == "const"

    @property
    def is_sym(self):
        return self.op == "sym"


_table: Dict[tuple, Node] = {}
_next_id = [0]


def _new(op, args=(), val=None, name=None) -> Node:
    n = Node(op, args, val, name, _next_id[0])
    _next_id[0] += 1
    return n


def _key_of_float(v: float):
    if v != v:
        return "nan"
    if v == 0.0:
        v = 0.0  # merge -0.0 and +0.0
    return struct.pack("<d", v)


def const(v) -> Node:
    v = float(v)
    if v == 0.0:
        v = 0.0
    key = ("const", _key_of_float(v))
    n = _table.get(key)
    if n is None:
        n = _new("const", (), v)
        _table[key] = n
    return n


ZERO = const(0.0)
ONE = const(1.0)
MINUS_ONE = const(-1.0)
TWO = const(2.0)
HALF = const(0.5)


def symbol(name: str) -> Node:
    """A fresh symbol.  Symbols are never merged, even with equal names (CasADi semantics)."""
    return _new("sym", (), None, name)


def _intern(op: str, *args: Node) -> Node:
    if op in _COMMUTATIVE and args[0].id > args[1].id:
        args = (args[1], args[0])
    key = (op,) + tuple(a.id for a in args)
    n = _table.get(key)
    if n is None:
        n = _new(op, tuple(args))
        _table[key] = n
    return n


def as_node(x) -> Node:
    if isinstance(x, Node):
        return x
    if isinstance(x, (bool, np.bool_)):
        return ONE if x else ZERO
    return const(x)


# ----------------------------------------------------------------------------------------------
# scalar numeric semantics (shared by constant folding and the NumPy evaluator)
# ----------------------------------------------------------------------------------------------

def _sign(x):
    return np.sign(x)


_NP1 = {
    "neg": np.negative, "sin": np.sin, "cos": np.cos, "tan": np.tan, "asin": np.arcsin,
    "acos": np.arccos, "atan": np.arctan, "exp": np.exp, "log": np.log, "sqrt": np.sqrt,
    "fabs": np.abs, "sign": _sign, "floor": np.floor, "ceil": np.ceil,
    "not": lambda a: (a == 0.0).astype(np.float64) if isinstance(a, np.ndarray) else float(a == 0.0),
}


def _b(x):
    return x.astype(np.float64) if isinstance(x, np.ndarray) else float(x)


_NP2 = {
    "add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide,
    "atan2": np.arctan2, "pow": np.power, "fmin": np.fmin, "fmax": np.fmax,
    "lt": lambda a, b: _b(a < b), "le": lambda a, b: _b(a <= b),
    "eq": lambda a, b: _b(a == b), "ne": lambda a, b: _b(a != b),
    "and": lambda a, b: _b(np.logical_and(a != 0, b != 0)),
    "or": lambda a, b: _b(np.logical_or(a != 0, b != 0)),
}


def _fold(op, *vals):
    with np.errstate(all="ignore"):
        if op in _NP1:
            return float(_NP1[op](np.float64(vals[0])))
        if op in _NP2:
            return float(_NP2[op](np.float64(vals[0]), np.float64(vals[1])))
        if op == "if_else":
            return float(vals[1] if vals[0] != 0.0 else vals[2])
    raise KeyError(op)


# ----------------------------------------------------------------------------------------------
# simplifying constructors
# ----------------------------------------------------------------------------------------------

def add(a, b) -> Node:
    a, b = as_node(a), as_node(b)
    if a.is_const and b.is_const:
        return const(a.val + b.val)
    if a is ZERO:
        return b
    if b is ZERO:
        return a
    if b.op == "neg":
        return sub(a, b.args[0])
    if a.op == "neg":
        return sub(b, a.args[0])
    return _intern("add", a, b)


def sub(a, b) -> Node:
    a, b = as_node(a), as_node(b)
    if a.is_const and b.is_const:
        return const(a.val - b.val)
    if b is ZERO:
        return a
    if a is ZERO:
        return neg(b)
    if a is b:
        return ZERO
    if b.op == "neg":
        return add(a, b.args[0])
    return _intern("sub", a, b)


def neg(a) -> Node:
    a = as_node(a)
    if a.is_const:
        return const(-a.val)
    if a.op == "neg":
        return a.args[0]
    if a.op == "sub":
        return sub(a.args[1], a.args[0])
    return _intern("neg", a)


def mul(a, b) -> Node:
    a, b = as_node(a), as_node(b)
    if a.is_const and b.is_const:
        return const(a.val * b.val)
    if a is ZERO or b is ZERO:
        return ZERO
    if a is ONE:
        return b
    if b is ONE:
        return a
    if a is MINUS_ONE:
        return neg(b)
    if b is MINUS_ONE:
        return neg(a)
    # pull negations outwards so that x*y and (-x)*y share the product
    if a.op == "neg" and b.op == "neg":
        return mul(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(mul(a.args[0], b))
    if b.op == "neg":
        return neg(mul(a, b.args[0]))
    if a.is_const and a.val < 0:
        return neg(mul(const(-a.val), b))
    if b.is_const and b.val < 0:
        return neg(mul(a, const(-b.val)))
    return _intern("mul", a, b)


def div(a, b) -> Node:
    a, b = as_node(a), as_node(b)
    if a.is_const and b.is_const:
        with np.errstate(all="ignore"):
            return const(float(np.float64(a.val) / np.float64(b.val)))
    if a is ZERO:
        return ZERO
    if b is ONE:
        return a
    if b is MINUS_ONE:
        return neg(a)
    if a.op == "neg" and b.op == "neg":
        return div(a.args[0], b.args[0])
    if a.op == "neg":
        return neg(div(a.args[0], b))
    if b.op == "neg":
        return neg(div(a, b.args[0]))
    if b.is_const:
        # exact only when the reciprocal is a power of two
        m, _ = math.frexp(b.val)
        if abs(m) == 0.5:
            return mul(a, const(1.0 / b.val))
    return _intern("div", a, b)


def _unary(op, a) -> Node:
    a = as_node(a)
    if a.is_const:
        return const(_fold(op, a.val))
    if op == "sin" and a.op == "neg":
        return neg(_intern("sin", a.args[0]))
    if op == "cos" and a.op == "neg":
        return _intern("cos", a.args[0])
    if op == "fabs" and a.op == "neg":
        return _intern("fabs", a.args[0])
    if op == "fabs" and a.op in ("fabs", "sqrt"):
        return a
    if op in ("tan", "asin", "atan") and a.op == "neg":
        return neg(_intern(op, a.args[0]))
    return _intern(op, a)


def sin(a): return _unary("sin", a)
def cos(a): return _unary("cos", a)
def tan(a): return _unary("tan", a)
def asin(a): return _unary("asin", a)
def acos(a): return _unary("acos", a)
def atan(a): return _unary("atan", a)
def exp(a): return _unary("exp", a)
def log(a): return _unary("log", a)
def sqrt(a): return _unary("sqrt", a)
def fabs(a): return _unary("fabs", a)
def sign(a): return _unary("sign", a)
def floor(a): return _unary("floor", a)
def ceil(a): return _unary("ceil", a)


def _binary(op, a, b) -> Node:
    a, b = as_node(a), as_node(b)
    if a.is_const and b.is_const:
        return const(_fold(op, a.val, b.val))
    return _intern(op, a, b)


def atan2(a, b): return _binary("atan2", a, b)
def fmin(a, b): return _binary("fmin", a, b)
def fmax(a, b): return _binary("fmax", a, b)
def lt(a, b): return _binary("lt", a, b)
def le(a, b): return _binary("le", a, b)
def eq(a, b): return _binary("eq", a, b)
def ne(a, b): return _binary("ne", a, b)


def logic_and(a, b):
    a, b = as_node(a), as_node(b)
    if a.is_const:
        return ZERO if a.val == 0.0 else logic_not(logic_not(b))
    if b.is_const:
        return ZERO if b.val == 0.0 else logic_not(logic_not(a))
    return _intern("and", a, b)


def logic_or(a, b):
    a, b = as_node(a), as_node(b)
    if a.is_const:
        return ONE if a.val != 0.0 else logic_not(logic_not(b))
    if b.is_const:
        return ONE if b.val != 0.0 else logic_not(logic_not(a))
    return _intern("or", a, b)


def logic_not(a):
    a = as_node(a)
    if a.is_const:
        return ONE if a.val == 0.0 else ZERO
    if a.op == "not" and a.args[0].op in ("not", "lt", "le", "eq", "ne", "and", "or"):
        return a.args[0]
    return _intern("not", a)


def pow_(a, b) -> Node:
    a, b = as_node(a), as_node(b)
    if a.is_const and b.is_const:
        return const(_fold("pow", a.val, b.val))
    if b.is_const:
        if b.val == 0.0:
            return ONE
        if b.val == 1.0:
            return a
        if b.val == 2.0:
            return mul(a, a)
        if b.val == 0.5:
            return sqrt(a)
        if b.val == -1.0:
            return div(ONE, a)
    return _intern("pow", a, b)


def if_else(c, a, b) -> Node:
    c, a, b = as_node(c), as_node(a), as_node(b)
    if c.is_const:
        return a if c.val != 0.0 else b
    if a is b:
        return a
    return _intern("if_else", c, a, b)


# ----------------------------------------------------------------------------------------------
# traversal
# ----------------------------------------------------------------------------------------------

def topo(outputs: Iterable[Node]) -> List[Node]:
    """All nodes reachable from `outputs`, in topological (ascending id) order."""
    seen = {}
    stack = [o for o in outputs]
    while stack:
        n = stack.pop()
        if n.id in seen:
            continue
        seen[n.id] = n
        stack.extend(n.args)
    return [seen[k] for k in sorted(seen)]


def symbols_of(outputs: Iterable[Node]) -> List[Node]:
    return [n for n in topo(outputs) if n.op == "sym"]


def depends_on(outputs: Iterable[Node], syms: Iterable[Node]) -> bool:
    want = {s.id for s in syms}
    return any(n.id in want for n in topo(outputs))


# ----------------------------------------------------------------------------------------------
# forward-mode AD
# ----------------------------------------------------------------------------------------------

def forward(outputs: Sequence[Node], seeds: Dict[int, Node]) -> List[Node]:
    """Directional derivative of every output along `seeds` ({symbol id: tangent node})."""
    order = topo(outputs)
    d: Dict[int, Node] = {}
    for n in order:
        op = n.op
        if op == "const":
            t = ZERO
        elif op == "sym":
            t = seeds.get(n.id, ZERO)
        else:
            a = n.args
            da = [d[x.id] for x in a]
            if all(x is ZERO for x in da):
                t = ZERO
            elif op == "add":
                t = add(da[0], da[1])
            elif op == "sub":
                t = sub(da[0], da[1])
            elif op == "neg":
                t = neg(da[0])
            elif op == "mul":
                t = add(mul(da[0], a[1]), mul(a[0], da[1]))
            elif op == "div":
                # (a/b)' = (a' - (a/b) b') / b
                t = div(sub(da[0], mul(n, da[1])), a[1])
            elif op == "sin":
                t = mul(cos(a[0]), da[0])
            elif op == "cos":
                t = neg(mul(sin(a[0]), da[0]))
            elif op == "tan":
                t = mul(add(ONE, mul(n, n)), da[0])
            elif op == "asin":
                t = div(da[0], sqrt(sub(ONE, mul(a[0], a[0]))))
            elif op == "acos":
                t = neg(div(da[0], sqrt(sub(ONE, mul(a[0], a[0])))))
            elif op == "atan":
                t = div(da[0], add(ONE, mul(a[0], a[0])))
            elif op == "atan2":
                den = add(mul(a[0], a[0]), mul(a[1], a[1]))
                t = div(sub(mul(a[1], da[0]), mul(a[0], da[1])), den)
            elif op == "exp":
                t = mul(n, da[0])
            elif op == "log":
                t = div(da[0], a[0])
            elif op == "sqrt":
                t = div(da[0], mul(TWO, n))
            elif op == "pow":
                # general a**b
                t = ZERO
                if da[0] is not ZERO:
                    t = add(t, mul(mul(a[1], pow_(a[0], sub(a[1], ONE))), da[0]))
                if da[1] is not ZERO:
                    t = add(t, mul(mul(n, log(a[0])), da[1]))
            elif op == "fabs":
                t = mul(sign(a[0]), da[0])
            elif op in ("sign", "floor", "ceil", "lt", "le", "eq", "ne", "and", "or", "not"):
                t = ZERO
            elif op == "fmin":
                t = if_else(le(a[0], a[1]), da[0], da[1])
            elif op == "fmax":
                t = if_else(le(a[1], a[0]), da[0], da[1])
            elif op == "if_else":
                t = if_else(a[0], da[1], da[2])
            else:  # pragma: no cover
                raise NotImplementedError(op)
        d[n.id] = t
    return [d[o.id] for o in outputs]


def jacobian(outputs: Sequence[Node], wrt: Sequence[Node]) -> List[List[Node]]:
    """rows = outputs, cols = wrt symbols."""
    cols = []
    for s in wrt:
        if not s.is_sym:
            raise ValueError("jacobian: differentiation variable must be purely symbolic")
        cols.append(forward(outputs, {s.id: ONE}))
    return [[cols[j][i] for j in range(len(wrt))] for i in range(len(outputs))]


# ----------------------------------------------------------------------------------------------
# substitution and evaluation
# ----------------------------------------------------------------------------------------------

_CTOR = {
    "add": add, "sub": sub, "mul": mul, "div": div, "neg": neg, "sin": sin, "cos": cos, "tan": tan,
    "asin": asin, "acos": acos, "atan": atan, "atan2": atan2, "exp": exp, "log": log, "sqrt": sqrt,
    "pow": pow_, "fabs": fabs, "sign": sign, "floor": floor, "ceil": ceil, "fmin": fmin,
    "fmax": fmax, "lt": lt, "le": le, "eq": eq, "ne": ne, "and": logic_and, "or": logic_or,
    "not": logic_not, "if_else": if_else,
}


def substitute(outputs: Sequence[Node], mapping: Dict[int, Node]) -> List[Node]:
    """Replace symbols (by id) with nodes and rebuild (re-simplifying on the way)."""
    order = topo(outputs)
    new: Dict[int, Node] = {}
    for n in order:
        if n.op == "const":
            new[n.id] = n
        elif n.op == "sym":
            new[n.id] = mapping.get(n.id, n)
        else:
            args = [new[a.id] for a in n.args]
            if all(x is y for x, y in zip(args, n.args)):
                new[n.id] = n
            else:
                new[n.id] = _CTOR[n.op](*args)
    return [new[o.id] for o in outputs]


def evaluate(outputs: Sequence[Node], values: Dict[int, object]):
    """NumPy evaluation.  `values` maps symbol id -> float or 1-D array (batch); arrays broadcast.
    Returns a list with one float/array per output."""
    order = topo(outputs)
    v: Dict[int, object] = {}
    with np.errstate(all="ignore"):
        for n in order:
            op = n.op
            if op == "const":
                v[n.id] = n.val
            elif op == "sym":
                try:
                    v[n.id] = values[n.id]
                except KeyError:
                    raise ValueError("evaluate: no value for free symbol %s" % n.name)
            elif op in _NP1:
                v[n.id] = _NP1[op](v[n.args[0].id])
            elif op in _NP2:
                v[n.id] = _NP2[op](v[n.args[0].id], v[n.args[1].id])
            elif op == "if_else":
                c, a, b = (v[x.id] for x in n.args)
                if isinstance(c, np.ndarray) or isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
                    v[n.id] = np.where(np.asarray(c) != 0.0, a, b)
                else:
                    v[n.id] = a if c != 0.0 else b
            else:  # pragma: no cover
                raise NotImplementedError(op)
    return [v[o.id] for o in outputs]


# ----------------------------------------------------------------------------------------------
# statistics (operation count used by the roofline accounting)
# ----------------------------------------------------------------------------------------------

def op_histogram(outputs: Sequence[Node]) -> Dict[str, int]:
    h: Dict[str, int] = {}
    for n in topo(outputs):
        if n.op in ("const", "sym"):
            continue
        h[n.op] = h.get(n.op, 0) + 1
    return h
