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
    rows = [[cols[j][i] for j in range(len(wrt))] for i in range(len(outputs))]
    if _AD_MODE[0] == "auto" and _chain_blocks and len(wrt) > 1:
        for i, o in enumerate(outputs):
            alt = _reverse_row(o, wrt)
            # keep the smaller graph; the primal `o` is in both so that shared work is not counted against either
            if alt is not None and graph_cost(alt + [o]) < graph_cost(rows[i] + [o]):
                rows[i] = alt
    return rows


# ----------------------------------------------------------------------------------------------
# kinematic-chain blocks and reverse-mode rows
# ----------------------------------------------------------------------------------------------
# A forward-kinematics transform T(q) = [R p] of an n-joint chain is the one place where forward-mode
# AD is badly matched to the problem: every joint direction drags its own 3x3 tangent through the rest
# of the chain, although d T / d q_j has a closed form in the TIP frame,
#     dR = R [a_j]x dq_j ,   dp = R (o_j x a_j) dq_j        (revolute; a_j = joint axis, o_j = a point on it,
#     dR = 0 ,               dp = R a_j dq_j                  (prismatic)        both in tip coordinates)
# whose ingredients a_j, o_j come out of the SUFFIX transforms the chain product computes anyway.  The FK
# front-end registers that structure as a ChainBlock; `jacobian` then offers, per scalar output row, a
# second derivation: reverse-mode AD down to the block's entries (adjoints F_R, F_p), pulled back
# through the block as   d e / d q_j = a_j . (kappa(R' F_R) + (R' F_p) x o_j),   and keeps whichever of
# the two graphs is smaller.  Orientation-error rows (||R_des' R - I||_F and the like) shrink several
# fold; position rows usually stay forward.  Both derivations are exact, they differ by rounding only.

class ChainBlock(object):
    """T: 3x4 nested list of nodes [R | p]; joints: list of (arg node, "revolute" | "prismatic",
    a_b (3 nodes), o_b (3 nodes)) with axis / axis point in tip coordinates."""

    def __init__(self, T, joints):
        self.T = [list(r) for r in T]
        self.joints = [(a, k, list(ab), list(ob)) for a, k, ab, ob in joints]

    def entries(self):
        return [n for r in self.T for n in r]

    def all_nodes(self):
        out = self.entries()
        for a, _, ab, ob in self.joints:
            out += [a] + ab + ob
        return out


_chain_blocks: Dict[int, "ChainBlock"] = {}
import os as _os
_AD_MODE = [_os.environ.get("CLIK_AD_MODE", "auto")]        # "auto": smaller of forward / reverse-through-blocks per row; "forward": forward only


class ad_mode(object):
    """Context manager: `with dag.ad_mode("forward"):` forces plain forward-mode Jacobians (the tests'
    oracle bridge uses it, so the oracle never shares the block pull-back with the product)."""

    def __init__(self, mode):
        if mode not in ("auto", "forward"):
            raise ValueError(mode)
        self.mode = mode

    def __enter__(self):
        self.prev = _AD_MODE[0]
        _AD_MODE[0] = self.mode

    def __exit__(self, *exc):
        _AD_MODE[0] = self.prev


def register_chain_block(block: "ChainBlock"):
    for n in block.entries():
        if n.op not in ("const", "sym"):
            _chain_blocks[n.id] = block


class _NoReverseRule(Exception):
    pass


def _acc(adj, node, val):
    if node.op == "const" or val is ZERO:
        return
    cur = adj.get(node.id)
    adj[node.id] = val if cur is None else add(cur, val)


def _reverse_to_leaves(output: Node, cut: Dict[int, "ChainBlock"]):
    """Adjoints d output / d leaf for the leaves of output's graph, where symbols and the entries of
    registered chain blocks are leaves.  -> ({leaf id: adjoint node}, {leaf id: leaf node})"""
    seen: Dict[int, Node] = {}
    stack = [output]
    while stack:
        n = stack.pop()
        if n.id in seen:
            continue
        seen[n.id] = n
        if n.id in cut or n.op in ("const", "sym"):
            continue
        stack.extend(n.args)
    adj: Dict[int, Node] = {output.id: ONE}
    leaves: Dict[int, Node] = {}
    for nid in sorted(seen, reverse=True):
        n = seen[nid]
        g = adj.get(nid)
        if g is None or n.op == "const":
            continue
        if n.id in cut or n.op == "sym":
            leaves[nid] = n
            continue
        a = n.args
        op = n.op
        if op == "add":
            _acc(adj, a[0], g)
            _acc(adj, a[1], g)
        elif op == "sub":
            _acc(adj, a[0], g)
            _acc(adj, a[1], neg(g))
        elif op == "neg":
            _acc(adj, a[0], neg(g))
        elif op == "mul":
            _acc(adj, a[0], mul(g, a[1]))
            _acc(adj, a[1], mul(g, a[0]))
        elif op == "div":
            q = div(g, a[1])
            _acc(adj, a[0], q)
            _acc(adj, a[1], neg(mul(q, n)))
        elif op == "sqrt":
            _acc(adj, a[0], div(g, mul(TWO, n)))
        elif op == "sin":
            _acc(adj, a[0], mul(g, cos(a[0])))
        elif op == "cos":
            _acc(adj, a[0], neg(mul(g, sin(a[0]))))
        elif op == "exp":
            _acc(adj, a[0], mul(g, n))
        elif op == "log":
            _acc(adj, a[0], div(g, a[0]))
        else:
            raise _NoReverseRule(op)      # piecewise / rarely used ops: the row stays forward-mode
    return adj, leaves


def _block_pullback(block: "ChainBlock", adj) -> List[Node]:
    """d e / d (joint argument j) from the adjoints of the block's entries (missing = zero)."""
    R = [block.T[i][:3] for i in range(3)]
    FR = [[adj.get(block.T[i][k].id, ZERO) if block.T[i][k].op != "const" else ZERO for k in range(3)] for i in range(3)]
    Fp = [adj.get(block.T[i][3].id, ZERO) if block.T[i][3].op != "const" else ZERO for i in range(3)]

    def dot3(u, v):
        return add(add(mul(u[0], v[0]), mul(u[1], v[1])), mul(u[2], v[2]))

    # K = R' F_R ; kappa = (K32 - K23, K13 - K31, K21 - K12) ; g = R' F_p
    K = [[dot3([R[0][i], R[1][i], R[2][i]], [FR[0][k], FR[1][k], FR[2][k]]) for k in range(3)] for i in range(3)]
    kappa = [sub(K[2][1], K[1][2]), sub(K[0][2], K[2][0]), sub(K[1][0], K[0][1])]
    g = [dot3([R[0][i], R[1][i], R[2][i]], Fp) for i in range(3)]
    out = []
    for _, kind, ab, ob in block.joints:
        if kind == "prismatic":
            out.append(dot3(g, ab))
        else:
            gxo = [sub(mul(g[1], ob[2]), mul(g[2], ob[1])), sub(mul(g[2], ob[0]), mul(g[0], ob[2])),
                   sub(mul(g[0], ob[1]), mul(g[1], ob[0]))]
            out.append(dot3(ab, [add(kappa[0], gxo[0]), add(kappa[1], gxo[1]), add(kappa[2], gxo[2])]))
    return out


_OP_COST = {"add": 1, "sub": 1, "mul": 1, "div": 4, "sqrt": 4, "sin": 10, "cos": 10}


def graph_cost(nodes: Sequence[Node]) -> int:
    return sum(_OP_COST.get(n.op, 0 if n.op in ("const", "sym", "neg") else 2) for n in topo(nodes))


def _reverse_row(output: Node, wrt: Sequence[Node]):
    """Row of the Jacobian by reverse mode through the registered chain blocks, or None."""
    blocks = {}
    for n in topo([output]):
        b = _chain_blocks.get(n.id)
        if b is not None:
            blocks[id(b)] = b
    if not blocks:
        return None
    cut = {n.id: b for b in blocks.values() for n in b.entries() if n.op not in ("const", "sym")}
    try:
        adj, leaves = _reverse_to_leaves(output, cut)
    except _NoReverseRule:
        return None
    # total derivative w.r.t. a requested symbol s: direct adjoint + sum over joint arguments of
    # (pull-back to that argument) x d argument / d s   (arguments are the symbols themselves unless the
    # FK function was called with expressions)
    args, pulls = [], []
    for b in blocks.values():
        pb = _block_pullback(b, adj)
        for (arg, _, _, _), pj in zip(b.joints, pb):
            args.append(arg)
            pulls.append(pj)
    row = []
    for s_ in wrt:
        t = adj.get(s_.id, ZERO) if s_.id in leaves else ZERO
        darg = forward(args, {s_.id: ONE}) if args else []
        for pj, da in zip(pulls, darg):
            if da is not ZERO and pj is not ZERO:
                t = add(t, mul(pj, da))
        row.append(t)
    return row


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
    # chain blocks whose entries are being rebuilt travel with them (their axis data is substituted too)
    touched = {}
    for n in order:
        b = _chain_blocks.get(n.id)
        if b is not None:
            touched[id(b)] = b
    extra = [m for b in touched.values() for m in b.all_nodes()]
    if extra:
        order = topo(list(outputs) + extra)
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
    for b in touched.values():
        nb = ChainBlock([[new[m.id] for m in r] for r in b.T],
                        [(new[a.id], k, [new[m.id] for m in ab], [new[m.id] for m in ob]) for a, k, ab, ob in b.joints])
        if any(x is not y for x, y in zip(nb.all_nodes(), b.all_nodes())):
            register_chain_block(nb)
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
