"""Drop-in check: the simulation cell of examples/notebooks/ur5_moe2016_example2.ipynb (cells 2-12)
written against casclik_b200 with only the imports changed, for the pseudo-inverse and the QP
controller, and compared step by step with the oracle-driven loop."""
import numpy as np
import pytest

from oracle_bridge import orc, oracle_pinv, oracle_qp_problem, close
import casclik_b200 as cc
from casclik_b200 import cs
from casclik_b200.fk import converter, UR5_URDF

pytestmark = pytest.mark.gpu


def _notebook_skill():
    # --- cell 2 -----------------------------------------------------------------------------------
    fk_dict = converter.from_file(root="base_link", tip="tool0", filename=UR5_URDF)
    link_lengths = [0., -0.425, -0.392, 0., 0., 0.]
    link_twists = [cs.np.pi / 2, 0., 0., cs.np.pi / 2, -cs.np.pi / 2, 0.]
    link_offsets = [0.089, 0., 0., 0.109, 0.095, 0.082]
    joint_angles = ["s" for i in range(6)]
    fk_dict = converter.from_denavit_hartenberg(
        joint_angles=joint_angles, link_lengths=link_lengths, link_offsets=link_offsets,
        link_twists=link_twists, joint_names=fk_dict["joint_names"],
        upper_limits=fk_dict["upper"], lower_limits=fk_dict["lower"])
    # --- cell 4 -----------------------------------------------------------------------------------
    t = cs.MX.sym("t")
    q = cs.MX.sym("q", len(fk_dict["joint_names"]))
    dq = cs.MX.sym("dq", len(fk_dict["joint_names"]))
    T_fk = fk_dict["T_fk"]
    p_fk = cs.Function("p_fk", [t, q], [T_fk(q)[:3, 3]])
    # --- cells 7-8 --------------------------------------------------------------------------------
    x_min, x_max = 0.1, 0.6
    y_min, y_max = -0.5, 0.4
    z_min, z_max = -0.3, 0.25
    omega = 0.1
    path_des = cs.vertcat(0.5 * cs.sin(omega * t) * cs.sin(omega * t) + 0.2,
                          0.5 * cs.cos(omega * t) + 0.25 * cs.sin(omega * t),
                          0.5 * cs.sin(omega * t) * cs.cos(omega * t) + 0.1)
    n = len(fk_dict["joint_names"])
    colav_x_cnstr = cc.SetConstraint(label="colav_x", expression=p_fk(t, q)[0], set_min=x_min,
                                     set_max=x_max, priority=n + 2, constraint_type="hard", gain=5e2)
    colav_y_cnstr = cc.SetConstraint(label="colav_y", expression=p_fk(t, q)[1], set_min=y_min,
                                     set_max=y_max, priority=n + 1, constraint_type="hard", gain=5e2)
    colav_z_cnstr = cc.SetConstraint(label="colav_z", expression=p_fk(t, q)[2], set_min=z_min,
                                     set_max=z_max, priority=n + 3, constraint_type="hard", gain=5e2)
    path_cnstr = cc.EqualityConstraint(label="move_point2", expression=p_fk(t, q) - path_des,
                                       priority=n + 4, constraint_type="soft", gain=0.15)
    path_cnstr.eval = cs.Function("path_eval", [t, q], [cs.norm_2(path_cnstr.expression)])
    skill = cc.SkillSpecification(label="box_move", time_var=t, robot_var=q, robot_vel_var=dq,
                                  constraints=[colav_x_cnstr, colav_y_cnstr, colav_z_cnstr, path_cnstr])
    return skill, T_fk, path_cnstr


UR5_home = cs.np.array([-(50.0 / 180.0) * cs.np.pi, -(160.0 / 180.0) * cs.np.pi,
                        -(110.0 / 180.0) * cs.np.pi, -(90.0 / 180.0) * cs.np.pi,
                        -(90.0 / 180.0) * cs.np.pi, 0.0])
dt = 0.008
max_speed = cs.np.pi / 5


@pytest.mark.parametrize("key", ["pinv", "qp"])
def test_notebook_simulation_cell(key, capsys):
    skill, T_fk, path_cnstr = _notebook_skill()
    skill.print_constraints()
    assert capsys.readouterr().out.startswith(
        "SkillSpecification: box_move\n#0: colav_y\n#1: colav_x\n#2: colav_z\n#3: move_point2\n")
    controller_classes = {"qp": cc.ReactiveQPController, "pinv": cc.PseudoInverseController}
    # --- cell 11 ----------------------------------------------------------------------------------
    ctrl = controller_classes[key](skill_spec=skill)
    ctrl.setup_problem_functions()
    ctrl.setup_solver()
    # --- cell 12 (shortened to 60 steps) ------------------------------------------------------------
    timesteps = 60
    ctrl.setup_initial_problem_solver()
    slack_res = ctrl.solve_initial_problem(0, UR5_home)[-1]
    t_sim = cs.np.array([dt * i for i in range(timesteps + 1)])
    q_sim = cs.np.zeros((len(t_sim), 6))
    q_sim[0, :] = UR5_home
    dq_sim = cs.np.zeros((len(t_sim), 6))
    p_sim = cs.np.zeros((len(t_sim), 3))
    p_sim[0, :] = T_fk(UR5_home)[:3, 3].toarray()[:, 0]
    e_sim = cs.np.zeros(len(t_sim))
    e_sim[0] = path_cnstr.eval(t_sim[0], q_sim[0, :])
    mode_sim = cs.np.zeros(len(t_sim))
    q_ref = UR5_home.copy()
    for i in range(len(t_sim) - 1):
        res = ctrl.solve(t_sim[i], q_sim[i, :], warmstart_slack_var=slack_res)
        dq_sim[i, :] = res[0].toarray()[:, 0]
        if res[-1] is not None:
            slack_res = res[-1].toarray()[:, 0]
        # the same step from the oracle, at the SAME state (so errors do not compound)
        inp = {"t": np.array([t_sim[i]]), "q": q_sim[i, :].reshape(6, 1)}
        if key == "pinv":
            v_ref, mode_ref = oracle_pinv(skill, inp)
            assert ctrl.current_mode == int(mode_ref[0])
            assert close(dq_sim[i, :], v_ref[:, 0], 1e-9, 1e-12).all()
        else:
            h, A, lb, ub = oracle_qp_problem(skill, inp)
            xo, _, sto = orc.solve_qp_single(h, A[0], lb[0], ub[0])
            assert sto == 0 and np.abs(dq_sim[i, :] - xo[:6]).max() < 1e-7 * (1 + np.abs(xo).max())
            assert np.abs(slack_res - xo[6:]).max() < 1e-7 * (1 + np.abs(xo).max())
        for idx, dqi in enumerate(dq_sim[i, :]):
            dq_sim[i, idx] = max(min(dqi, max_speed), -max_speed)
        q_sim[i + 1, :] = q_sim[i, :] + dq_sim[i, :] * dt
        p_sim[i + 1, :] = T_fk(q_sim[i + 1, :])[:3, 3].toarray()[:, 0]
        e_sim[i + 1] = path_cnstr.eval(t_sim[i], q_sim[i + 1, :])
        if key == "pinv":
            mode_sim[i + 1] = ctrl.current_mode
    assert np.isfinite(q_sim).all() and e_sim[-1] < e_sim[0]      # the arm moves towards the path
    if key == "qp":
        assert slack_res.shape == (3,)
