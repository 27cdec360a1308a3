"""Generate tests/golden/api_behaviour.json: what the REFERENCE's classes do for the scripted
constructor / validation / bookkeeping cases of tests/golden_api_cases.py (returned summary, or the
type of the exception raised).  The unmodified reference package is imported over the stand-in
casadi module, exactly as in make_controller_vectors.py.
Run in the build container (needs /root/reference):  python tests/golden/make_api_behaviour.py"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import casclik_b200.sym as sym  # noqa: E402

shim = types.ModuleType("casadi")
for _n in dir(sym):
    if not _n.startswith("_"):
        setattr(shim, _n, getattr(sym, _n))
sys.modules["casadi"] = shim
sys.path.insert(0, "/root/reference")
import casclik as ref  # noqa: E402

assert ref.__file__.startswith("/root/reference/"), ref.__file__

import golden_skills as gs  # noqa: E402
import golden_api_cases as api  # noqa: E402

ns = gs.Namespace(shim, ref)
out = {name: api.run_case(ns, name) for name in sorted(api.CASES)}
for name, res in out.items():
    print("%-36s %s" % (name, json.dumps(res)[:150]))
path = os.path.join(HERE, "api_behaviour.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", path)
