"""GPU: the plugin surface end to end -- make_ocp(...).solve() in a receding-horizon loop (run_mpc.py:115-143) against
the oracle's restatement, host buffers in / host buffers out."""
import numpy as np
import pytest

from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("warm", [True, False])
@pytest.mark.parametrize("rn,kind,N", [("b2", "whole_body_rnea", 6), ("go2", "centroidal_vel", 5)])
def test_mpc_loop_matches_oracle(robots, rn, kind, N, warm):
    """warm = False: run_mpc.py with its warm_start flag off -- every solve() starts from opti.initial() (DX = 0,
    U = u_des), only the OSQP iterates carry over."""
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    prod, ora = robots
    B, loops, dt_min = 2, 3, 0.01
    ocp = make_ocp(dynamics=kind, default_args=OCP_ARGS[kind], robot=prod[rn], nodes=N, solver="osqp", batch=B)
    oracles = [OracleOCP(ora[rn], kind, N) for _ in range(B)]
    sqps = [OracleSQP(o) for o in oracles]
    t0 = np.array([0.0, 0.21])
    base_vel = np.array([0.2, 0, 0, 0, 0, 0])

    def configure(o, x_init, t):
        o.set_time_params(dt_min, 0.08)
        o.set_swing_params(0.07, [0.1, -0.2])
        o.set_tracking_targets(base_vel, np.zeros(3), np.zeros(3))
        o.update_initial_state(x_init)
        o.update_gait_sequence(t)
        if kind == "whole_body_rnea":
            o.update_previous_torques(np.zeros(o.nj))

    x_init = np.stack([ocp.x_nom, ocp.x_nom])
    configure(ocp, x_init, t0)
    ocp.init_solver()
    x_ref = [None] * B
    xi_ref = [ocp.x_nom.copy() for _ in range(B)]
    for b in range(B):
        configure(oracles[b], xi_ref[b], t0[b])
        sqps[b].init_solver()
    for k in range(loops):
        t = t0 + k * dt_min
        ocp.update_initial_state(x_init)
        ocp.update_gait_sequence(t)
        if warm:
            ocp.warm_start()
        sol = ocp.solve(retract_all=False)
        x_init = ocp.state_integrate(x_init, ocp.DX_prev[1])
        assert ocp.stats.shape == (B, 8) and ocp.solve_time > 0
        for b in range(B):
            o = oracles[b]
            o.update_initial_state(xi_ref[b])
            o.update_gait_sequence(t[b])
            xw = o.warm_start() if warm else o.initial_guess()
            x_ref[b], info = sqps[b].solve(xw, o.p_vector())
            o.retract_stacked_sol(x_ref[b])
            xi_ref[b] = o.dyn.state_integrate()(xi_ref[b], o.DX_prev[1])
            scale = max(1.0, np.abs(x_ref[b]).max())
            assert np.abs(sol[b] - x_ref[b]).max() <= 1e-6 * scale, (k, b)
            assert int(ocp.stats[b, 0]) == info["qp_iters"]
            assert np.abs(x_init[b] - xi_ref[b]).max() <= 1e-6
    assert len(ocp.q_sol) == loops and ocp.q_sol[0].shape == (B, ocp.nq)


def test_solution_history_is_not_overwritten(robots):
    """forces_sol / a_sol / tau_sol hold copies (the reference stores np.array(...)): entries of earlier solves must
    survive the reuse of the pinned result buffers two solves later."""
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    prod, _ = robots
    kind = "whole_body_rnea"
    ocp = make_ocp(dynamics=kind, default_args=OCP_ARGS[kind], robot=prod["b2"], nodes=5, solver="osqp", batch=2)
    ocp.set_time_params(0.01, 0.08)
    ocp.set_swing_params(0.07, [0.1, -0.2])
    ocp.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]), np.zeros(3), np.zeros(3))
    x_init = np.stack([ocp.x_nom, ocp.x_nom])
    ocp.update_initial_state(x_init)
    ocp.update_gait_sequence(0.0)
    ocp.update_previous_torques(np.zeros(ocp.nj))
    ocp.init_solver()
    snaps = []
    for k in range(4):
        ocp.update_gait_sequence(k * 0.01)
        ocp.warm_start()
        ocp.solve(retract_all=False)
        snaps.append((ocp.forces_sol[k].copy(), ocp.a_sol[k].copy(), ocp.tau_sol[k].copy(), ocp.q_sol[k].copy()))
    assert len(ocp.forces_sol) == 4
    for k in range(4):
        assert np.array_equal(ocp.forces_sol[k], snaps[k][0]) and np.array_equal(ocp.a_sol[k], snaps[k][1])
        assert np.array_equal(ocp.tau_sol[k], snaps[k][2]) and np.array_equal(ocp.q_sol[k], snaps[k][3])
    assert not np.array_equal(ocp.forces_sol[0], ocp.forces_sol[2])


def test_solve_without_warm_start_restarts_from_initial_point(robots):
    """solve() does not move opti.initial(): two solves without warm_start() start from the same point (only the OSQP
    iterates differ); set_initial() moves it."""
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    prod, _ = robots
    kind = "centroidal_acc"
    ocp = make_ocp(dynamics=kind, default_args=OCP_ARGS[kind], robot=prod["b2"], nodes=5, solver="osqp", batch=1)
    ocp.set_time_params(0.01, 0.08)
    ocp.set_swing_params(0.07, [0.1, -0.2])
    ocp.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]), np.zeros(3), np.zeros(3))
    ocp.update_initial_state(ocp.x_nom)
    ocp.update_gait_sequence(0.0)
    ocp.init_solver()
    s1 = ocp.solve().copy()
    assert ocp._x0 is None
    # a fresh OCP stepped from set_initial(initial_guess) gives the same first step
    ocp2 = make_ocp(dynamics=kind, default_args=OCP_ARGS[kind], robot=prod["b2"], nodes=5, solver="osqp", batch=1)
    ocp2.set_time_params(0.01, 0.08)
    ocp2.set_swing_params(0.07, [0.1, -0.2])
    ocp2.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]), np.zeros(3), np.zeros(3))
    ocp2.update_initial_state(ocp2.x_nom)
    ocp2.update_gait_sequence(0.0)
    ocp2.init_solver()
    ocp2.set_initial(ocp2.initial_guess())
    assert np.array_equal(ocp2.solve(), s1)


def test_solver_errors(robots):
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    prod, _ = robots
    ocp = make_ocp(dynamics="centroidal_acc", default_args=OCP_ARGS["centroidal_acc"], robot=prod["b2"], nodes=5, solver="fatrop", batch=1)
    with pytest.raises(NotImplementedError):
        ocp.init_solver()
