"""Generate tests/golden/activation_maps.json by running the REFERENCE's own
PseudoInverseController.create_activation_map (casclik/controllers/pseudo_inverse.py:107-130).

CasADi is not installable in the build container, but this one method only touches `cs.np`, so a
stub module that exposes NumPy as `casadi.np` is enough to execute the unmodified reference code.
Run in the build container (needs /root/reference):  python tests/golden/make_activation_maps.py
"""
import json
import os
import sys
import types

import numpy as np

stub = types.ModuleType("casadi")
stub.np = np
sys.modules["casadi"] = stub
sys.path.insert(0, "/root/reference")
from casclik.controllers.pseudo_inverse import PseudoInverseController  # noqa: E402


class _Shell(object):
    pass


out = {}
for n_sets in range(0, 8):
    shell = _Shell()
    shell.n_set_constraints = n_sets
    PseudoInverseController.create_activation_map(shell)
    out[str(n_sets)] = [[int(b) for b in row] for row in shell.activation_map]

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "activation_maps.json")
with open(path, "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("wrote", path, {k: len(v) for k, v in out.items()})
