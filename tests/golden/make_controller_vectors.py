"""Generate tests/golden/controller_vectors.json by running the REFERENCE's own controllers.

The unmodified reference package (/root/reference/casclik: constraints.py, skill_specification.py,
controllers/pseudo_inverse.py, controllers/reactive_qp.py) is imported and driven through its public
API — SkillSpecification(...), PseudoInverseController(...).setup_solver(), .solve(t, q, x, y) —
one instance at a time, exactly like its notebooks do.  CasADi itself cannot be installed in the
build container, so `import casadi` resolves to a stand-in module: casclik_b200.sym (expression
graph + forward AD, evaluated by its NumPy interpreter in float64).  So every line of the reference's
controller logic runs as written (mode expressions, damped pseudo-inverses, null-space chain,
in-tangent-cone functions, sequential mode search, H/A/lb/ub assembly); what is substituted is the
symbolic backend underneath it and, for the QP controller, the numerical QP solver behind
`casadi.conic` (qpOASES is not available: the stand-in solves the reference-built H, A, lba, uba
with oracle/clik_oracle.py's active-set method; H > 0, so the minimiser is unique and the fixture
stores the matrices next to the solution).

Run in the build container (needs /root/reference):  python tests/golden/make_controller_vectors.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import casclik_b200.sym as sym  # noqa: E402
import clik_oracle as orc  # noqa: E402

shim = types.ModuleType("casadi")
for _n in dir(sym):
    if not _n.startswith("_"):
        setattr(shim, _n, getattr(sym, _n))


class _CpuConic(object):
    """Stand-in for casadi.conic(name, "qpoases", {"h": sp, "a": sp}, opts): callable with
    h=, a=, lba=, uba= (and x0=, ignored: the minimiser is unique) -> {"x": DM}."""
    last = None

    def __init__(self, name, solver_name, sparsities, opts=None):
        self.name = name

    def __call__(self, h=None, a=None, lba=None, uba=None, x0=None, **kw):
        H = np.array(sym.DM(h).full(), dtype=float)
        A = np.array(sym.DM(a).full(), dtype=float)
        lb = np.array(sym.DM(lba).full(), dtype=float).reshape(-1)
        ub = np.array(sym.DM(uba).full(), dtype=float).reshape(-1)
        assert np.count_nonzero(H - np.diag(np.diag(H))) == 0
        hd = np.diag(H).copy()
        x, lam, status = orc.solve_qp_single(hd, A, lb, ub)
        assert status == 0
        kk = orc.kkt_residuals(hd, A, lb, ub, x)
        assert max(kk["primal"], kk["stationarity"], kk["sign"]) < 1e-8, kk
        _CpuConic.last = {"h": hd, "A": A, "lb": lb, "ub": ub, "x": x}
        return {"x": sym.DM(x.reshape(-1, 1))}


shim.conic = _CpuConic
sys.modules["casadi"] = shim
sys.path.insert(0, "/root/reference")
import casclik as ref  # noqa: E402

assert ref.__file__.startswith("/root/reference/"), ref.__file__

import golden_skills as gs  # noqa: E402

ns = gs.Namespace(shim, ref)


def _arg(a, i):
    return None if a is None else np.ascontiguousarray(a[:, i])


def _vec(dm):
    return None if dm is None else [float(v) for v in np.array(sym.DM(dm).full(), dtype=float).reshape(-1)]


def _select(outputs, kind, keep_first=8, per_mode=3, cap=48):
    """First few instances + a few examples of every mode (pinv) / working-set size (QP) seen."""
    def key(o):
        if kind == "pinv":
            return o["mode"]
        x = np.array(o["robot_vel"] + (o["virtual_vel"] or []) + (o["slack"] or []))
        r = np.array(o["A"]) @ x
        tol = 1e-9 * np.maximum(1.0, np.abs(r))
        return (tuple(np.nonzero(np.abs(r - np.array(o["lb"])) <= tol)[0]),
                tuple(np.nonzero(np.abs(r - np.array(o["ub"])) <= tol)[0]))
    keep = list(range(min(keep_first, len(outputs))))
    seen = {}
    for i, o in enumerate(outputs):
        k = key(o)
        if seen.get(k, 0) < per_mode and i not in keep:
            seen[k] = seen.get(k, 0) + 1
            keep.append(i)
    return sorted(keep[:cap])


# closed-loop simulations (the loop of the notebooks around solve(): q += clip(v)*dt, t = t0 + k*dt)
ROLLOUTS = {"pinv/ur5_moe2016": dict(steps=300, dt=0.008, max_speed=0.6, max_virtual_speed=None, n=4),
            "pinv/cart_path": dict(steps=200, dt=0.02, max_speed=0.275, max_virtual_speed=0.5, n=4),
            "pinv/iiwa_multitask": dict(steps=120, dt=0.01, max_speed=1.0, max_virtual_speed=None, n=4),
            "qp/ur5_qp": dict(steps=80, dt=0.008, max_speed=None, max_virtual_speed=None, n=3)}


def _rollout(ctrl, kind, inp, cfg):
    res = {"config": cfg, "q_final": [], "x_final": [], "last_mode": [], "modes_seen": []}
    for i in range(cfg["n"]):
        q = np.array(inp["q"][:, i], dtype=float)
        x = np.array(inp["x"][:, i], dtype=float) if "x" in inp else None
        y = _arg(inp.get("y"), i)
        t0 = float(inp["t"][i])
        seen = set()
        for k in range(cfg["steps"]):
            rv, vv, _ = ctrl.solve(t0 + k * cfg["dt"], q, x, y)
            v = np.array(sym.DM(rv).full(), dtype=float).reshape(-1)
            if cfg["max_speed"] is not None:
                v = np.clip(v, -cfg["max_speed"], cfg["max_speed"])
            q = q + v * cfg["dt"]
            if x is not None:
                w = np.array(sym.DM(vv).full(), dtype=float).reshape(-1)
                if cfg["max_virtual_speed"] is not None:
                    w = np.clip(w, -cfg["max_virtual_speed"], cfg["max_virtual_speed"])
                x = x + w * cfg["dt"]
            if kind == "pinv":
                seen.add(int(ctrl.current_mode))
        res["q_final"].append(q.tolist())
        res["x_final"].append(None if x is None else x.tolist())
        res["last_mode"].append(int(ctrl.current_mode) if kind == "pinv" else None)
        res["modes_seen"].append(sorted(seen))
    return res


out = {}
for name in sorted(gs.CASES):
    spec, inp, kind, kwargs = gs.build(ns, name)
    N = inp["q"].shape[1]
    rec = {"controller": kind, "kwargs": json.loads(json.dumps(kwargs)), "inputs": None, "outputs": []}
    if kind == "pinv":
        ctrl = ref.PseudoInverseController(skill_spec=spec, **kwargs)
        ctrl.setup_solver()
        for i in range(N):
            rv, vv, _ = ctrl.solve(float(inp["t"][i]), _arg(inp["q"], i), _arg(inp.get("x"), i), _arg(inp.get("y"), i))
            rec["outputs"].append({"robot_vel": _vec(rv), "virtual_vel": _vec(vv), "mode": int(ctrl.current_mode)})
        modes = sorted(set(o["mode"] for o in rec["outputs"]))
    else:
        ctrl = ref.ReactiveQPController(skill_spec=spec, **kwargs)
        ctrl.setup_problem_functions()
        ctrl.setup_solver()
        for i in range(N):
            rv, vv, sl = ctrl.solve(float(inp["t"][i]), _arg(inp["q"], i), _arg(inp.get("x"), i), _arg(inp.get("y"), i))
            m = _CpuConic.last
            rec["outputs"].append({"robot_vel": _vec(rv), "virtual_vel": _vec(vv), "slack": _vec(sl),
                                   "h": m["h"].tolist(), "A": m["A"].tolist(), "lb": m["lb"].tolist(),
                                   "ub": m["ub"].tolist()})
        modes = []
        # initial-value problem (reactive_qp.py:297-424, :426-459) on the first few instances
        ctrl.setup_initial_problem_solver()
        rec["initial"] = []
        rng = np.random.default_rng(gs.seed_of(name) + 500)
        for i in range(6):
            dq0 = rng.uniform(-0.3, 0.3, inp["q"].shape[0])
            virt, slack = ctrl.solve_initial_problem(float(inp["t"][i]), _arg(inp["q"], i), _arg(inp.get("x"), i),
                                                     dq0, _arg(inp.get("y"), i))
            m = _CpuConic.last
            rec["initial"].append({"dq0": dq0.tolist(), "virtual": _vec(virt), "slack": _vec(slack),
                                   "h": m["h"].tolist(), "A": m["A"].tolist(), "lb": m["lb"].tolist(),
                                   "ub": m["ub"].tolist()})
    keep = _select(rec["outputs"], kind)
    rec["outputs"] = [rec["outputs"][i] for i in keep]
    rec["inputs"] = {k: (np.asarray(v)[keep] if np.ndim(v) == 1 else np.asarray(v)[:, keep]).tolist()
                     for k, v in inp.items()}
    if kind == "pinv":
        modes = sorted(set(o["mode"] for o in rec["outputs"]))
    if name in ROLLOUTS:
        kept = {k: np.array(v, dtype=float) for k, v in rec["inputs"].items()}
        rec["rollout"] = _rollout(ctrl, kind, kept, ROLLOUTS[name])
        print("   rollout: modes seen", rec["rollout"]["modes_seen"])
    out[name] = rec
    print("%-42s %3d of %d instances kept  modes %s" % (name, len(keep), N, modes))

path = os.path.join(HERE, "controller_vectors.json")
with open(path, "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes")
