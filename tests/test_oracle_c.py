"""The C restatement (oracle/clik_oracle.c, used as the timed CPU baseline) agrees with the NumPy
oracle, and the oracle's independent geometric FK agrees with the expression graph."""
import os
import sys

import numpy as np

from oracle_bridge import orc, oracle_pinv, close
from casclik_b200 import scenarios, fk

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import c_port  # noqa: E402


def test_c_port_matches_numpy_oracle_ur5_track():
    sc = scenarios.get("ur5_track")
    inp = sc.sample(2000, seed=11)
    ref_v, _ = oracle_pinv(sc.spec, inp)
    chain = orc.load_chain(fk.UR5_URDF, "base_link", "tool0")
    got, used = c_port.pinv_track(chain, inp["q"], inp["y"], gain=1.0, lam=1e-7, threads=2)
    assert used == 2
    assert close(got, ref_v, 1e-9, 1e-12).all(), np.abs(got - ref_v).max()


def test_independent_fk_matches_expression_graph():
    sc = scenarios.get("ur5_track")
    inp = sc.sample(500, seed=4)
    chain = orc.load_chain(fk.UR5_URDF, "base_link", "tool0")
    p, J = orc.position_jacobian(chain, inp["q"].T.copy())
    from oracle_bridge import blocks_from_skill
    blocks, n = blocks_from_skill(sc.spec, inp["t"], inp["q"], None, inp["y"])
    assert n == 6
    assert np.abs(blocks[0].e - (p - inp["y"].T)).max() < 1e-14
    assert np.abs(blocks[0].J - J).max() < 1e-14      # AD Jacobian == geometric Jacobian
    assert np.all(blocks[0].Jt == 0.0)
