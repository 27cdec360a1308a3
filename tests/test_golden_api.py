"""casclik_b200's classes against tests/golden/api_behaviour.json — what the REFERENCE's classes
return or raise for the scripted cases of golden_api_cases.py (fixture made by
tests/golden/make_api_behaviour.py with the unmodified reference package)."""
import json
import os

import pytest

import casclik_b200 as cc
from casclik_b200 import cs
import golden_api_cases as api
import golden_skills as gs

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_behaviour.json")) as f:
    GOLDEN = json.load(f)
NS = gs.Namespace(cs, cc)

# Deliberate differences, each a defect of the reference that a drop-in should not reproduce:
DEVIATIONS = {
    # an (n, 1) numpy column as set_max: the reference's size check compares the wrong variable
    # (`szi == 1`, constraints.py:262) and rejects what it accepts for set_min two branches earlier
    "set/column_array_bounds": {"ok": [[-1.0, -2.0], [1.0, 2.0], 1, "hard"]},
    # DM / MX weights: the reference's setters compare the bound method `weights.size2` with 1
    # (reactive_qp.py:69-74) and so reject every CasADi-typed weight vector
    "qp/dm_weights": {"ok": {"H_diag": [0.001, 0.002, 0.003, 0.001, 2.501, 2.501, 1.001], "H_offdiag_nnz": 0,
                             "mu": 0.001, "solver": "qpoases"}},
}


def test_fixture_and_cases_agree():
    assert sorted(GOLDEN) == sorted(api.CASES)
    assert set(DEVIATIONS) <= set(GOLDEN)


@pytest.mark.parametrize("name", sorted(api.CASES))
def test_class_api_behaves_like_the_reference(name):
    got = api.run_case(NS, name)
    want = DEVIATIONS.get(name, GOLDEN[name])
    assert got == want, (name, got, GOLDEN[name])
