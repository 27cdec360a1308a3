"""Forward-kinematics front-end (URDF / DH -> expressions); see converter.py."""
import os

from . import converter  # noqa: F401
from .converter import from_file, from_denavit_hartenberg, parse_urdf, chain  # noqa: F401
from . import pose_errors  # noqa: F401
from . import casadi_geom, numpy_geom  # noqa: F401  (urdf2casadi.casadi_geom / numpy_geom stand-ins)

ROBOT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "robots")
UR5_URDF = os.path.join(ROBOT_DIR, "ur5_chain.urdf")
IIWA14_URDF = os.path.join(ROBOT_DIR, "iiwa14_chain.urdf")


def ur5():
    """UR5 base_link -> tool0 (the chain every UR5 notebook of the reference uses)."""
    return from_file("base_link", "tool0", UR5_URDF)


def iiwa14():
    """KUKA LBR iiwa 14 R820 base_link -> tool0 (7-DOF; BASELINE.json config 3)."""
    return from_file("base_link", "tool0", IIWA14_URDF)
