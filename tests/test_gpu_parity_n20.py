"""GPU parity against the oracle on BASELINE.json's own configurations at the full horizon N = 20 (round-1 parity
tests stopped at N = 4..6), through the C ABI (Handle -> plm_sqp_step / plm_sqp_data / plm_qp_*).

Tolerances (north-star): residual rows and Jacobian entries <= 1e-9 relative; SQP primal variables, cost and
violation <= 1e-6 with identical ADMM iteration counts, statuses, accepted step sizes and trial counts.
"""
import numpy as np
import pytest

from emu_util import random_problem
from oracle.model import OracleRobot
from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def nominal_problem(o, rng, k, base_vel=(0.2, 0, 0, 0, 0, 0), ext=5.0, arm=0.05):
    """The reference's own starting point (run_ocp.py:58-68): x_init = x_nom, x = opti.initial(), gait time k dt_min."""
    o.set_time_params(0.01, 0.08)
    o.set_swing_params(0.07, [0.1, -0.2])
    o.set_tracking_targets(list(base_vel), rng.uniform(-ext, ext, 3), rng.uniform(-arm, arm, 3))
    o.update_initial_state(o.x_nom)
    o.update_gait_sequence(k * 0.01)
    if o.kind == "whole_body_rnea":
        o.update_previous_torques(np.zeros(o.nj))
    return o.initial_guess(), o.p_vector()


def run_sqp_parity(prod_robot, ora_robot, kind, N, ks, iters, seed, gait=("trot", 0.8), tol=1e-6, base_vel=(0.2, 0, 0, 0, 0, 0)):
    """`iters` warm-started SQP iterations of len(ks) instances on the device and in the oracle, compared after every
    iteration.  Returns the per-iteration oracle infos of instance 0."""
    from pino_locoman_b200.handle import Handle
    rng = np.random.default_rng(seed)
    B = len(ks)
    prod_robot.set_gait_sequence(*gait)
    try:
        ocps = [OracleOCP(ora_robot, kind, N, gait_type=gait[0], gait_period=gait[1]) for _ in range(B)]
        xs, ps = zip(*[nominal_problem(o, rng, k, base_vel=base_vel) for o, k in zip(ocps, ks)])
        sqps = [OracleSQP(o) for o in ocps]
        for s in sqps:
            s.init_solver()
        h = Handle(prod_robot, kind, N, max_batch=B)
        assert (h.n, h.m, h.np) == (ocps[0].n, ocps[0].m, ocps[0].np_)
        x = torch.tensor(np.stack(xs), device="cuda")
        p = torch.tensor(np.stack(ps), device="cuda")
        xr = [np.array(v) for v in xs]
        infos = []
        for it in range(iters):
            x, stats = h.sqp_step(x, p)
            stats = stats.cpu().numpy()
            for b in range(B):
                xr[b], info = sqps[b].solve(xr[b], ps[b])
                if b == 0:
                    infos.append(info)
                assert int(stats[b, 0]) == info["qp_iters"], (it, b, stats[b], info["qp_iters"])
                assert bool(stats[b, 2]) == info["accepted"] and int(stats[b, 4]) == info["trials"], (it, b)
                if info["accepted"]:
                    assert stats[b, 3] == info["alpha"]
                scale = max(1.0, np.abs(xr[b]).max())
                err = np.abs(x[b].cpu().numpy() - xr[b]).max()
                assert err <= tol * scale, (it, b, err, scale)
                assert abs(stats[b, 5] - info["f"]) <= tol * max(1.0, abs(info["f"]))
                assert abs(stats[b, 7] - info["violation_max"]) <= tol * max(1.0, info["violation_max"])
        return infos
    finally:
        prod_robot.set_gait_sequence("trot", 0.8)


@pytest.mark.parametrize("variant", ["throughput", "latency"])
def test_b2g_rnea_n20_sqp_matches_oracle(robots, variant, monkeypatch):
    """BASELINE configs[4] formulation (the bench workload): 21 stages, sparse-coupling path, tau_nodes = 3; both
    instantiations of the ADMM kernel."""
    monkeypatch.setenv("PLM_ADMM_LATENCY_MAX_BATCH", "0" if variant == "throughput" else "1000000")
    prod, ora = robots
    run_sqp_parity(prod["b2g"], ora["b2g"], "whole_body_rnea", 20, (0, 33), 3, seed=21)


def test_b2g_aba_n20_sqp_matches_oracle(robots):
    """BASELINE configs[2] formulation: dense integrator rows (the dense-coupling branch of the factor / ADMM kernels)."""
    prod, ora = robots
    run_sqp_parity(prod["b2g"], ora["b2g"], "whole_body_aba", 20, (0, 57), 3, seed=22)


def test_go2_centroidal_vel_n20_sqp_matches_oracle(robots):
    """BASELINE configs[0] formulation through the OSQP path."""
    prod, ora = robots
    run_sqp_parity(prod["go2"], ora["go2"], "centroidal_vel", 20, (0, 41), 3, seed=23)


@pytest.mark.parametrize("rn,kind", [("b2", "centroidal_acc"), ("b2", "whole_body_acc")])
def test_other_formulations_n20_sqp_matches_oracle(robots, rn, kind):
    """BASELINE configs[3] formulation (centroidal_acc) and whole_body_acc through QP + line search at N = 20."""
    prod, ora = robots
    run_sqp_parity(prod[rn], ora[rn], kind, 20, (0, 17), 3, seed=24)


def test_b2_rnea_n20_single_instance_long_run(robots):
    """BASELINE configs[1]: one B2 whole_body_rnea instance from the reference's own start (run_ocp.py:58-68), iterated as
    far as the reference's loop goes.  The reference runs ONE SQP iteration per solve() (optimization/ocp.py:383) with
    OSQP truncated at max_iter = 100 / eps = 1e-3 and a fixed rho, so its iterates never reach a violation of 1e-6
    (the oracle stalls at 1e-2 .. 1e-1 after 40 iterations, and a single QP does not reach eps = 1e-9 within 20 000 ADMM
    iterations): "converged" is what this sequence reaches.  Twelve warm-started iterations are compared one by one (every
    iteration feeds the next one on both sides: primal variables, cost, violation, ADMM iteration counts, step sizes)."""
    prod, ora = robots
    infos = run_sqp_parity(prod["b2"], ora["b2"], "whole_body_rnea", 20, (0,), 12, seed=25)
    # the sequence does make progress: the violation drops by two orders of magnitude from the first iterate
    assert infos[-1]["violation_max"] < 0.05 * infos[0]["violation_max"]


@pytest.mark.parametrize("gait", [("walk", 0.8), ("stand", 0.8)])
def test_other_gaits_on_device(robots, gait):
    """SURVEY 8f rank 4 on the device (round 1 had host emulation only): walk / stand schedules, eval + SQP parity."""
    prod, ora = robots
    vel = (0.0, 0, 0, 0, 0, 0) if gait[0] == "stand" else (0.2, 0, 0, 0, 0, 0)
    run_sqp_parity(prod["b2g"], ora["b2g"], "whole_body_rnea", 20, (0, 29), 2, seed=26, gait=gait, base_vel=vel)
    run_sqp_parity(prod["go2"], ora["go2"], "centroidal_vel", 8, (5, 61), 2, seed=27, gait=gait, base_vel=vel)


@pytest.mark.parametrize("payload", ["front", "rear"])
def test_b2_payload_frames_on_device(payload):
    """B2(payload=...) external-force frame on the base body (utils/robot.py:70-76): residuals / Jacobian at a random
    point and SQP iterations from the nominal point, on the device."""
    from pino_locoman_b200.handle import Handle
    from pino_locoman_b200.utils.robot import B2
    prod = B2(payload=payload)
    prod.set_gait_sequence("trot", 0.8)
    ora = OracleRobot("b2", payload=payload)
    assert prod.nf == ora.nf == 15
    rng = np.random.default_rng(28)
    for kind in ("whole_body_rnea", "centroidal_acc", "whole_body_aba"):
        o = OracleOCP(ora, kind, 4)
        h = Handle(prod, kind, 4, max_batch=2)
        x, p = random_problem(o, rng)
        xd = torch.tensor(np.tile(x, (2, 1)), device="cuda")
        pd = torch.tensor(np.tile(p, (2, 1)), device="cuda")
        _, J, g, lbg, ubg = h.sqp_data(xd, pd)
        g_ref, lb_ref, ub_ref = o.g_data(x, p)
        J_ref = o.jac_g(x, p)
        assert np.abs(g[1].cpu().numpy() - g_ref).max() <= 1e-9 * max(1.0, np.abs(g_ref).max())
        assert np.abs(h.jac_dense(J)[1].cpu().numpy() - J_ref).max() <= 1e-9 * np.abs(J_ref).max()
        assert np.array_equal(lbg[0].cpu().numpy(), lb_ref) and np.array_equal(ubg[0].cpu().numpy(), ub_ref)
    run_sqp_parity(prod, ora, "whole_body_rnea", 20, (0, 47), 2, seed=29)


@pytest.mark.parametrize("rn,kind", [("b2g", "whole_body_rnea"), ("b2g", "whole_body_aba"), ("go2", "centroidal_vel"), ("b2", "whole_body_acc")])
def test_eval_matches_oracle_n20(robots, rn, kind):
    """sqp_data at N = 20 on random states (SURVEY 8d distributions): g, J_g, grad_f, bounds."""
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(30)
    o = OracleOCP(ora[rn], kind, 20)
    h = Handle(prod[rn], kind, 20, max_batch=2)
    probs = [random_problem(o, rng) for _ in range(2)]
    x = torch.tensor(np.stack([q[0] for q in probs]), device="cuda")
    p = torch.tensor(np.stack([q[1] for q in probs]), device="cuda")
    grad, J, g, lbg, ubg = h.sqp_data(x, p)
    Jd = h.jac_dense(J).cpu().numpy()
    for b, (xb, pb) in enumerate(probs):
        g_ref, lb_ref, ub_ref = o.g_data(xb, pb)
        J_ref = o.jac_g(xb, pb)
        _, grad_ref = o.f_data(xb, pb)
        assert np.abs(g[b].cpu().numpy() - g_ref).max() <= 1e-9 * max(1.0, np.abs(g_ref).max())
        assert np.abs(Jd[b] - J_ref).max() <= 1e-9 * np.abs(J_ref).max()
        assert np.abs(grad[b].cpu().numpy() - grad_ref).max() <= 1e-9 * np.abs(grad_ref).max()
        assert np.array_equal(lbg[b].cpu().numpy(), lb_ref) and np.array_equal(ubg[b].cpu().numpy(), ub_ref)
