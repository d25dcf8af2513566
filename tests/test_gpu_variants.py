"""GPU parity of the formulations without base inputs (include_base=False) against the oracle:
ocp_centroidal_vel.py:104-120,174-185 (v_b = base_vel_dynamics(h, q, v_j)), ocp_centroidal_acc.py:108-140 and
ocp_whole_body_acc.py:109-141 (a_b = base_acc_dynamics(q, v, a_j, forces)); no dynamics-gap rows."""
import numpy as np
import pytest

from emu_util import random_problem
from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP
from test_gpu_parity_n20 import nominal_problem

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CASES = [("b2", "centroidal_acc"), ("b2g", "whole_body_acc"), ("go2", "centroidal_vel"), ("b2g", "centroidal_vel")]


@pytest.mark.parametrize("rn,kind", CASES)
def test_eval_matches_oracle(robots, rn, kind):
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(40)
    N = 6
    o = OracleOCP(ora[rn], kind, N, include_base=False)
    h = Handle(prod[rn], kind, N, max_batch=3, include_base=False)
    assert (h.n, h.m, h.np) == (o.n, o.m, o.np_)
    probs = [random_problem(o, rng) for _ in range(3)]
    probs[2] = (o.initial_guess(), probs[2][1])            # DX = 0: the point of every first SQP iteration
    x = torch.tensor(np.stack([q[0] for q in probs]), device="cuda")
    p = torch.tensor(np.stack([q[1] for q in probs]), device="cuda")
    grad, J, g, lbg, ubg = h.sqp_data(x, p)
    g2, _, _ = h.g_data(x, p, bounds=False)
    assert torch.equal(g, g2)                              # residual-only mode (what the line search evaluates)
    Jd = h.jac_dense(J).cpu().numpy()
    for b, (xb, pb) in enumerate(probs):
        g_ref, lb_ref, ub_ref = o.g_data(xb, pb)
        J_ref = o.jac_g(xb, pb)
        f_ref, grad_ref = o.f_data(xb, pb)
        assert np.abs(g[b].cpu().numpy() - g_ref).max() <= 1e-9 * max(1.0, np.abs(g_ref).max())
        assert np.abs(Jd[b] - J_ref).max() <= 1e-9 * np.abs(J_ref).max()
        assert np.abs(grad[b].cpu().numpy() - grad_ref).max() <= 1e-9 * max(1.0, np.abs(grad_ref).max())
        assert np.array_equal(lbg[b].cpu().numpy(), lb_ref) and np.array_equal(ubg[b].cpu().numpy(), ub_ref)
    assert np.array_equal(h.hess_diag(p)[0].cpu().numpy(), o.hess_diag(probs[0][1]))


@pytest.mark.parametrize("variant", ["throughput", "latency"])
@pytest.mark.parametrize("rn,kind,N", [("b2", "centroidal_acc", 6), ("b2g", "whole_body_acc", 5), ("go2", "centroidal_vel", 6), ("b2g", "centroidal_vel", 20)])
def test_sqp_iterations_match_oracle(robots, rn, kind, N, variant, monkeypatch):
    """Warm-started SQP iterations (sqp_data -> OSQP -> Armijo) through plm_sqp_step: dense base-integrator rows take the
    dense-coupling path of the factor / ADMM kernels."""
    from pino_locoman_b200.handle import Handle
    monkeypatch.setenv("PLM_ADMM_LATENCY_MAX_BATCH", "0" if variant == "throughput" else "1000000")
    prod, ora = robots
    rng = np.random.default_rng(41)
    B, iters = 2, 3
    ocps = [OracleOCP(ora[rn], kind, N, include_base=False) for _ in range(B)]
    xs, ps = zip(*[nominal_problem(o, rng, k) for o, k in zip(ocps, (0, 37))])
    sqps = [OracleSQP(o) for o in ocps]
    for s in sqps:
        s.init_solver()
    h = Handle(prod[rn], kind, N, max_batch=B, include_base=False)
    x = torch.tensor(np.stack(xs), device="cuda")
    p = torch.tensor(np.stack(ps), device="cuda")
    xr = [np.array(v) for v in xs]
    for it in range(iters):
        x, stats = h.sqp_step(x, p)
        stats = stats.cpu().numpy()
        for b in range(B):
            xr[b], info = sqps[b].solve(xr[b], ps[b])
            assert int(stats[b, 0]) == info["qp_iters"], (it, b)
            assert bool(stats[b, 2]) == info["accepted"] and int(stats[b, 4]) == info["trials"], (it, b)
            scale = max(1.0, np.abs(xr[b]).max())
            assert np.abs(x[b].cpu().numpy() - xr[b]).max() <= 1e-6 * scale, (it, b)
            assert abs(stats[b, 5] - info["f"]) <= 1e-6 * max(1.0, abs(info["f"]))
            assert abs(stats[b, 7] - info["violation_max"]) <= 1e-6 * max(1.0, info["violation_max"])


@pytest.mark.parametrize("rn,kind", [("b2", "centroidal_acc"), ("go2", "centroidal_vel")])
def test_plugin_surface_mpc_loop(robots, rn, kind):
    """make_ocp(..., include_base=False) in the receding-horizon loop of run_mpc.py:115-143 against the oracle."""
    from pino_locoman_b200.optimization import make_ocp
    prod, ora = robots
    N, B, loops, dt_min = 5, 2, 3, 0.01
    ocp = make_ocp(dynamics=kind, default_args={"include_base": False}, robot=prod[rn], nodes=N, solver="osqp", batch=B)
    oracles = [OracleOCP(ora[rn], kind, N, include_base=False) for _ in range(B)]
    sqps = [OracleSQP(o) for o in oracles]
    t0 = np.array([0.0, 0.21])

    def configure(o, x_init, t):
        o.set_time_params(dt_min, 0.08)
        o.set_swing_params(0.07, [0.1, -0.2])
        o.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]), np.zeros(3), np.zeros(3))
        o.update_initial_state(x_init)
        o.update_gait_sequence(t)

    x_init = np.stack([ocp.x_nom, ocp.x_nom])
    configure(ocp, x_init, t0)
    ocp.init_solver()
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
        for b in range(B):
            o = oracles[b]
            o.update_initial_state(xi_ref[b])
            o.update_gait_sequence(t[b])
            x_ref, info = sqps[b].solve(o.warm_start(), o.p_vector())
            o.retract_stacked_sol(x_ref)
            xi_ref[b] = o.dyn.state_integrate()(xi_ref[b], o.DX_prev[1])
            assert np.abs(sol[b] - x_ref).max() <= 1e-6 * max(1.0, np.abs(x_ref).max()), (k, b)
            assert int(ocp.stats[b, 0]) == info["qp_iters"]
            assert np.abs(x_init[b] - xi_ref[b]).max() <= 1e-6
    if kind == "centroidal_vel":      # retracted velocities carry the base part from base_vel (ocp_centroidal_vel.py:224-231)
        assert ocp.v_sol[0].shape == (B, ocp.nv)
        hq = np.concatenate((ocp.h_sol[-1], ocp.q_sol[-1]), 1)
        vb = oracles[0].dyn.base_vel_dynamics()(hq[0, :6], hq[0, 6:], ocp.v_sol[-1][0, 6:])
        assert np.abs(vb - ocp.v_sol[-1][0, :6]).max() <= 1e-9 * max(1.0, np.abs(vb).max())


def test_rnea_without_acceleration_inputs(robots, monkeypatch):
    """include_acc=False (ocp_whole_body_rnea.py:21-25,156,183-191): a_i = (v_{i+1} - v_i) / dt_i, no dv integrator rows; the
    RNEA rows touch dv_{i+1}, which the QP kernels handle on their general-coupling path.  Rows / Jacobian on random
    states, then warm-started SQP iterations on both instantiations of the ADMM kernel."""
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(42)
    for rn, N in (("b2", 6), ("b2g", 5)):
        o = OracleOCP(ora[rn], "whole_body_rnea", N, include_acc=False)
        h = Handle(prod[rn], "whole_body_rnea", N, max_batch=2, include_acc=False)
        assert (h.n, h.m, h.np) == (o.n, o.m, o.np_)
        probs = [random_problem(o, rng) for _ in range(2)]
        x = torch.tensor(np.stack([q[0] for q in probs]), device="cuda")
        p = torch.tensor(np.stack([q[1] for q in probs]), device="cuda")
        grad, J, g, lbg, ubg = h.sqp_data(x, p)
        Jd = h.jac_dense(J).cpu().numpy()
        for b, (xb, pb) in enumerate(probs):
            g_ref, lb_ref, ub_ref = o.g_data(xb, pb)
            J_ref = o.jac_g(xb, pb)
            assert np.abs(g[b].cpu().numpy() - g_ref).max() <= 1e-9 * max(1.0, np.abs(g_ref).max())
            assert np.abs(Jd[b] - J_ref).max() <= 1e-9 * np.abs(J_ref).max()
            assert np.array_equal(lbg[b].cpu().numpy(), lb_ref) and np.array_equal(ubg[b].cpu().numpy(), ub_ref)
    for variant in ("throughput", "latency"):
        monkeypatch.setenv("PLM_ADMM_LATENCY_MAX_BATCH", "0" if variant == "throughput" else "1000000")
        for rn, N in (("b2", 6), ("b2g", 20)):
            B, iters = 2, 3
            ocps = [OracleOCP(ora[rn], "whole_body_rnea", N, include_acc=False) for _ in range(B)]
            xs, ps = zip(*[nominal_problem(o, rng, k) for o, k in zip(ocps, (0, 37))])
            sqps = [OracleSQP(o) for o in ocps]
            for s in sqps:
                s.init_solver()
            h = Handle(prod[rn], "whole_body_rnea", N, max_batch=B, include_acc=False)
            x = torch.tensor(np.stack(xs), device="cuda")
            p = torch.tensor(np.stack(ps), device="cuda")
            xr = [np.array(v) for v in xs]
            for it in range(iters):
                x, stats = h.sqp_step(x, p)
                stats = stats.cpu().numpy()
                for b in range(B):
                    xr[b], info = sqps[b].solve(xr[b], ps[b])
                    assert int(stats[b, 0]) == info["qp_iters"], (variant, rn, it, b, stats[b], info["qp_iters"])
                    assert bool(stats[b, 2]) == info["accepted"] and int(stats[b, 4]) == info["trials"], (variant, rn, it, b)
                    assert np.abs(x[b].cpu().numpy() - xr[b]).max() <= 1e-6 * max(1.0, np.abs(xr[b]).max()), (variant, rn, it, b)
                    assert abs(stats[b, 5] - info["f"]) <= 1e-6 * max(1.0, abs(info["f"]))
