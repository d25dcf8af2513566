"""GPU: the plugin surface end to end -- make_ocp(...).solve() in a receding-horizon loop (run_mpc.py:115-143) against
the oracle's restatement, host buffers in / host buffers out."""
import numpy as np
import pytest

from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rn,kind,N", [("b2", "whole_body_rnea", 6), ("go2", "centroidal_vel", 5)])
def test_mpc_loop_matches_oracle(robots, rn, kind, N):
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
        ocp.warm_start()
        sol = ocp.solve(retract_all=False)
        x_init = ocp.state_integrate(x_init, ocp.DX_prev[1])
        assert ocp.stats.shape == (B, 8) and ocp.solve_time > 0
        for b in range(B):
            o = oracles[b]
            o.update_initial_state(xi_ref[b])
            o.update_gait_sequence(t[b])
            xw = o.warm_start()
            x_ref[b], info = sqps[b].solve(xw, o.p_vector())
            o.retract_stacked_sol(x_ref[b])
            xi_ref[b] = o.dyn.state_integrate()(xi_ref[b], o.DX_prev[1])
            scale = max(1.0, np.abs(x_ref[b]).max())
            assert np.abs(sol[b] - x_ref[b]).max() <= 1e-6 * scale, (k, b)
            assert int(ocp.stats[b, 0]) == info["qp_iters"]
            assert np.abs(x_init[b] - xi_ref[b]).max() <= 1e-6
    assert len(ocp.q_sol) == loops and ocp.q_sol[0].shape == (B, ocp.nq)


def test_solver_errors(robots):
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    prod, _ = robots
    ocp = make_ocp(dynamics="centroidal_acc", default_args=OCP_ARGS["centroidal_acc"], robot=prod["b2"], nodes=5, solver="fatrop", batch=1)
    with pytest.raises(NotImplementedError):
        ocp.init_solver()
