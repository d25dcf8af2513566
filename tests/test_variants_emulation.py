"""More formulations / robots through the kernel mathematics (host emulation) against the oracle:
B2 with a payload external-force frame on the base body, B2G without the arm (locked joints merged into the base),
walk / stand gaits, the first SQP iterate of the reference (x = initial guess)."""
import numpy as np
import pytest

from emu_util import Emu, random_problem
from oracle.model import OracleRobot
from oracle.ocp import OracleOCP

TOL = 1e-9


def _check(prod_robot, ora_robot, kind, N, rng, gait=("trot", 0.8), exact_guess=False):
    prod_robot.set_gait_sequence(*gait)
    o = OracleOCP(ora_robot, kind, N, gait_type=gait[0], gait_period=gait[1])
    e = Emu(prod_robot, kind, N)
    assert (e.n, e.m, e.np_) == (o.n, o.m, o.np_)
    x, p = random_problem(o, rng)
    if exact_guess:
        x = o.initial_guess()
    g_ref, _, _ = o.g_data(x, p)
    J_ref = o.jac_g(x, p)
    g, Jv = e.eval(x, p)
    assert np.abs(g[0] - g_ref).max() <= TOL * max(1.0, np.abs(g_ref).max())
    assert np.abs(e.dense(Jv[0]) - J_ref).max() <= TOL * np.abs(J_ref).max()


@pytest.mark.parametrize("payload", ["front", "rear"])
def test_b2_payload_external_force_on_base(payload):
    from pino_locoman_b200.utils.robot import B2
    rng = np.random.default_rng(2)
    prod = B2(payload=payload)
    ora = OracleRobot("b2", payload=payload)
    assert prod.nf == ora.nf == 15
    for kind in ("whole_body_rnea", "centroidal_acc", "whole_body_aba"):
        _check(prod, ora, kind, 3, rng)


def test_b2g_without_arm_locks_all_arm_joints():
    from pino_locoman_b200.utils.robot import B2G
    prod = B2G(ignore_arm=True)
    ora = OracleRobot("b2g", ignore_arm=True)
    assert prod.nq == ora.nq == 19 and prod.nf == 12 and prod.arm_ee_frame is None
    assert abs(prod.mass - 77.26826983) < 1e-6           # arm links merged into the base body
    _check(prod, ora, "whole_body_rnea", 3, np.random.default_rng(3))


@pytest.mark.parametrize("gait", [("walk", 0.8), ("stand", 0.8), ("trot", 0.5)])
def test_other_gaits(robots, gait):
    prod, ora = robots
    try:
        _check(prod["b2g"], ora["b2g"], "whole_body_rnea", 4, np.random.default_rng(4), gait=gait)
        _check(prod["go2"], ora["go2"], "centroidal_vel", 4, np.random.default_rng(5), gait=gait)
    finally:
        for r in prod.values():
            r.set_gait_sequence("trot", 0.8)


def test_reference_initial_guess_point(robots):
    """x = opti.initial(): DX = 0 and U = u_des -- the point every first SQP iteration is evaluated at."""
    prod, ora = robots
    rng = np.random.default_rng(6)
    for rn, kind in (("b2", "whole_body_rnea"), ("b2g", "whole_body_aba"), ("go2", "centroidal_vel"), ("b2", "whole_body_acc")):
        _check(prod[rn], ora[rn], kind, 3, rng, exact_guess=True)


@pytest.mark.parametrize("rn,kind,kw", [("b2", "centroidal_acc", {"include_base": False}), ("b2g", "whole_body_acc", {"include_base": False}),
                                        ("go2", "centroidal_vel", {"include_base": False}), ("b2g", "centroidal_vel", {"include_base": False}),
                                        ("b2g", "centroidal_acc", {"include_base": False}), ("go2", "whole_body_acc", {"include_base": False}),
                                        ("b2g", "whole_body_rnea", {"include_acc": False}), ("b2", "whole_body_rnea", {"include_acc": False})])
def test_formulations_without_base_inputs(robots, rn, kind, kw):
    """include_base=False (ocp_centroidal_vel.py:104-120, ocp_centroidal_acc.py:129-140, ocp_whole_body_acc.py:130-141): the base
    velocity / acceleration is solved for from the six gap rows inside the node evaluation; include_acc=False
    (ocp_whole_body_rnea.py:183-191): finite-difference accelerations.  Rows and the chain-rule Jacobian (dense base-integrator,
    foot- and arm-velocity rows; RNEA rows reaching into dv_{i+1}) against the oracle's complex-step differentiation."""
    prod, ora = robots
    rng = np.random.default_rng(12)
    for exact in (False, True):
        o = OracleOCP(ora[rn], kind, 4, **kw)      # (N = 4 with tau_nodes = 3: both RNEA node types)
        e = Emu(prod[rn], kind, 4, **kw)
        assert (e.n, e.m, e.np_) == (o.n, o.m, o.np_)
        x, p = random_problem(o, rng)
        if exact:
            x = o.initial_guess()
        g_ref, _, _ = o.g_data(x, p)
        J_ref = o.jac_g(x, p)
        g, Jv = e.eval(x, p)
        assert not np.isnan(Jv).any()                      # every pattern entry is written
        assert np.abs(g[0] - g_ref).max() <= TOL * max(1.0, np.abs(g_ref).max())
        assert np.abs(e.dense(Jv[0]) - J_ref).max() <= TOL * np.abs(J_ref).max()
        g2, _ = e.eval(x, p, want_jac=False)               # residual-only mode (line-search trials)
        assert np.array_equal(g2, g)
