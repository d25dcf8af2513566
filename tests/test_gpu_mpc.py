"""GPU: the device-resident receding-horizon step (plm_mpc_step / BatchedMPC) against the same loop driven through the
plugin surface from the host (run_mpc.py:115-143), whose parity with the oracle tests/test_gpu_ocp_api.py establishes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _configure(o, kind, x_init, t, dt_min):
    o.set_time_params(dt_min, 0.08)
    o.set_swing_params(0.07, [0.1, -0.2])
    o.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]), np.zeros(3), np.zeros(3))
    o.update_initial_state(x_init)
    o.update_gait_sequence(t)
    if kind == "whole_body_rnea":
        o.update_previous_torques(np.zeros(o.nj))


@pytest.mark.parametrize("rn,kind,N,gait,warm,kw", [("b2g", "whole_body_rnea", 6, "trot", True, {}), ("go2", "centroidal_vel", 5, "walk", True, {}),
                                                    ("b2", "whole_body_acc", 5, "stand", True, {}), ("b2g", "whole_body_rnea", 6, "trot", False, {}),
                                                    ("go2", "centroidal_vel", 5, "walk", False, {}),
                                                    ("go2", "centroidal_vel", 5, "trot", True, {"include_base": False}),
                                                    ("b2", "whole_body_rnea", 5, "trot", True, {"include_acc": False})])
def test_device_resident_mpc_matches_host_loop(rn, kind, N, gait, warm, kw):
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.mpc import BatchedMPC
    from pino_locoman_b200.optimization import make_ocp
    from pino_locoman_b200.utils import robot as prob
    B, loops, dt_min = 3, 4, 0.01
    t0 = np.array([0.0, 0.21, -0.13])      # a negative start time: Python's floor-mod, not C fmod

    def make():
        r = {"b2g": prob.B2G, "go2": prob.Go2, "b2": prob.B2}[rn]()
        r.set_gait_sequence(gait, 0.8)
        return make_ocp(dynamics=kind, default_args=dict(OCP_ARGS[kind], **kw), robot=r, nodes=N, solver="osqp", batch=B)

    host, dev = make(), make()
    x_init = np.stack([host.x_nom] * B)
    x_init[1, 7:host.nq] += 0.05          # instances differ
    _configure(host, kind, x_init, t0, dt_min)
    _configure(dev, kind, x_init, t0, dt_min)
    host.init_solver()
    dev.init_solver()
    mpc = BatchedMPC(dev, warm_start=warm, t0=t0)
    co, so = host.handle.p_off["contact_schedule"], host.handle.p_off["swing_schedule"]
    for k in range(loops):
        t = t0 + k * dt_min
        host.update_initial_state(x_init)
        host.update_gait_sequence(t)
        if warm:
            host.warm_start()
        sol = host.solve(retract_all=False)
        x_init = host.state_integrate(x_init, host.DX_prev[1])
        stats = mpc.step().cpu().numpy()
        p_dev = mpc.p.cpu().numpy()
        # gait schedules: bit for bit
        assert np.array_equal(p_dev[:, co:co + 4 * N], host._p[:, co:co + 4 * N]), k
        assert np.array_equal(p_dev[:, so:so + 4 * N], host._p[:, so:so + 4 * N]), k
        x_dev = mpc.solution().cpu().numpy()
        scale = max(1.0, np.abs(sol).max())
        assert np.abs(x_dev - sol).max() <= 1e-9 * scale, (k, np.abs(x_dev - sol).max())
        assert np.array_equal(stats[:, :2], host.stats[:, :2])          # ADMM iterations and status
        assert np.abs(mpc.x_init().cpu().numpy() - x_init).max() <= 1e-9
    assert mpc.k == loops


def test_device_resident_mpc_updates_previous_torques():
    """update_tau_prev: tau_prev <- tau of node 1 after every step (compiled-solver branch, run_mpc.py:108-111)."""
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.mpc import BatchedMPC
    from pino_locoman_b200.optimization import make_ocp
    from pino_locoman_b200.utils import robot as prob
    kind, B, N, loops, dt_min = "whole_body_rnea", 2, 5, 3, 0.01
    t0 = np.array([0.0, 0.3])
    args = dict(OCP_ARGS[kind], tau_nodes=3)

    def make():
        r = prob.B2G()
        r.set_gait_sequence("trot", 0.8)
        return make_ocp(dynamics=kind, default_args=args, robot=r, nodes=N, solver="osqp", batch=B)

    host, dev = make(), make()
    x_init = np.stack([host.x_nom] * B)
    x_init[1, 7:host.nq] -= 0.04
    _configure(host, kind, x_init, t0, dt_min)
    _configure(dev, kind, x_init, t0, dt_min)
    for o in (host, dev):
        o._set("W_diag", np.full(o.nj, 1e-3))       # the tau_prev term must matter
        o.init_solver()
    mpc = BatchedMPC(dev, warm_start=True, t0=t0, update_tau_prev=True)
    to = host.handle.p_off["tau_prev"]
    for k in range(loops):
        host.update_initial_state(x_init)
        host.update_gait_sequence(t0 + k * dt_min)
        host.warm_start()
        sol = host.solve(retract_all=False)
        x_init = host.state_integrate(x_init, host.DX_prev[1])
        tau1 = host.get_tau_sol(1).copy()
        host.update_previous_torques(tau1)
        mpc.step()
        x_dev = mpc.solution().cpu().numpy()
        assert np.abs(x_dev - sol).max() <= 1e-9 * max(1.0, np.abs(sol).max()), k
        assert np.array_equal(mpc.p.cpu().numpy()[:, to:to + host.nj], x_dev[:, host.handle.x_off[1] + host.ndx_opt + host.tau_idx:][:, :host.nj])
        assert np.abs(mpc.p.cpu().numpy()[:, to:to + host.nj] - tau1).max() <= 1e-9 * max(1.0, np.abs(tau1).max())
