"""GPU parity of the batched ADMM QP solver against the oracle's OSQP restatement (same iterate sequence)."""
import numpy as np
import pytest

from emu_util import random_problem
from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _nominal_problem(o, rng, k):
    o.set_time_params(0.01, 0.08)
    o.set_swing_params(0.07, [0.1, -0.2])
    o.set_tracking_targets([0.2, 0, 0, 0, 0, 0], rng.uniform(-5, 5, 3), rng.uniform(-0.05, 0.05, 3))
    o.update_initial_state(o.x_nom)
    o.update_gait_sequence(k * 0.01)
    if o.kind == "whole_body_rnea":
        o.update_previous_torques(np.zeros(o.nj))
    return o.initial_guess(), o.p_vector()


@pytest.mark.parametrize("variant", ["throughput", "latency"])
@pytest.mark.parametrize("rn,kind,N", [("b2", "whole_body_rnea", 6), ("b2g", "whole_body_rnea", 5), ("b2", "centroidal_acc", 6),
                                       ("go2", "centroidal_vel", 5), ("b2g", "whole_body_aba", 4), ("b2", "whole_body_acc", 5)])
def test_qp_matches_oracle_osqp(robots, rn, kind, N, variant, monkeypatch):
    """Both instantiations of the ADMM kernel: the 256-thread one used for large batches and the 512-thread one that
    small batches get (the library reads PLM_ADMM_LATENCY_MAX_BATCH at every solve)."""
    from pino_locoman_b200.handle import Handle
    monkeypatch.setenv("PLM_ADMM_LATENCY_MAX_BATCH", "0" if variant == "throughput" else "1000000")
    prod, ora = robots
    rng = np.random.default_rng(3)
    B = 3
    ocps = [OracleOCP(ora[rn], kind, N) for _ in range(B)]
    xs, ps = zip(*[_nominal_problem(o, rng, k) for o, k in zip(ocps, (0, 17, 41))])
    sqps = [OracleSQP(o) for o in ocps]
    for s in sqps:
        s.init_solver()
    h = Handle(prod[rn], kind, N, max_batch=B)
    x = torch.tensor(np.stack(xs), device="cuda")
    p = torch.tensor(np.stack(ps), device="cuda")
    hess = h.hess_diag(p)
    h.qp_setup(hess)
    for it in range(2):   # second round exercises the warm start and the re-scaling
        grad, J, g, lbg, ubg = h.sqp_data(x, p)
        h.qp_update(hess, grad, J, lbg - g, ubg - g)
        D, E, c = h.qp_get_scaling(B)
        dx, iters, status = h.qp_solve(B)
        xq, zq, yq = h.qp_get_iterates(B)
        new_x = []
        for b in range(B):
            o, s = ocps[b], sqps[b]
            xb = x[b].cpu().numpy()
            grad_r, J_r, g_r, lb_r, ub_r = o.sqp_data(xb, ps[b])
            s.osqp.update(q=grad_r, Ax=s.csc_values(J_r), l=lb_r - g_r, u=ub_r - g_r)
            dx_r = s.osqp.solve()
            Q = s.osqp
            assert np.abs(D[b].cpu().numpy() - Q.D).max() <= 1e-10 * np.abs(Q.D).max()
            assert np.abs(E[b].cpu().numpy() - Q.E).max() <= 1e-10 * np.abs(Q.E).max()
            assert abs(c[b].item() - Q.c) <= 1e-10 * Q.c
            assert iters[b].item() == Q.iters
            assert {1: "solved", 2: "solved inaccurate", -2: "maximum iterations reached"}[status[b].item()] == Q.status
            scale = max(1.0, np.abs(dx_r).max())
            assert np.abs(dx[b].cpu().numpy() - dx_r).max() <= 1e-6 * scale
            assert np.abs(xq[b].cpu().numpy() - Q.x).max() <= 1e-6 * max(1.0, np.abs(Q.x).max())
            assert np.abs(yq[b].cpu().numpy() - Q.y).max() <= 1e-6 * max(1.0, np.abs(Q.y).max())
            new_x.append(xb + dx_r)
        x = torch.tensor(np.stack(new_x), device="cuda")


def test_failure_codes(robots):
    """Numeric failures surface as per-instance status codes, never as exceptions."""
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(1)
    o = OracleOCP(ora["b2"], "centroidal_acc", 4)
    x0, p0 = _nominal_problem(o, rng, 3)
    h = Handle(prod["b2"], "centroidal_acc", 4, max_batch=3)
    x = torch.tensor(np.tile(x0, (3, 1)), device="cuda")
    p = torch.tensor(np.tile(p0, (3, 1)), device="cuda")
    hess = h.hess_diag(p)
    h.qp_setup(hess)
    grad, J, g, lbg, ubg = h.sqp_data(x, p)
    grad[1, 5] = float("nan")                 # NaN in the linear cost of instance 1
    bad_hess = hess.clone()
    bad_hess[2] = -1e9                        # indefinite P for instance 2: the stage Cholesky must report it
    h.qp_update(bad_hess, grad, J, lbg - g, ubg - g)
    dx, iters, status = h.qp_solve(3)
    status = status.cpu().numpy()
    assert status[0] in (1, 2, -2) and torch.isfinite(dx[0]).all()
    assert status[1] == -11
    assert status[2] == -10


def test_fp64_peak_microbenchmark():
    """plm_fp64_peak (the measured FP64 roofline denominator of bench.py) returns a plausible DFMA rate."""
    import ctypes
    from pino_locoman_b200 import _lib
    t = ctypes.c_double(0.0)
    assert _lib.load().plm_fp64_peak(ctypes.byref(t)) == 0
    assert 5.0 < t.value < 100.0, t.value          # B200: about 36-40 TFLOP/s
