"""GPU parity of the Armijo line search and of the fused SQP step against the oracle (OSQP + Armijo restatement)."""
import numpy as np
import pytest

from emu_util import random_problem
from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP, armijo_line_search
from test_gpu_qp import _nominal_problem

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rn,kind,N", [("b2", "whole_body_rnea", 5), ("b2g", "whole_body_aba", 4), ("go2", "centroidal_vel", 5)])
def test_line_search_matches_oracle(robots, rn, kind, N):
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(9)
    o = OracleOCP(ora[rn], kind, N)
    h = Handle(prod[rn], kind, N, max_batch=8)
    x0, p0 = random_problem(o, rng)
    # directions of very different quality: some accepted at a=1, some after back-tracking, some never
    g0, lb, ub = o.g_data(x0, p0)
    J = o.jac_g(x0, p0)
    viol = np.where(g0 < lb, lb - g0, np.where(g0 > ub, ub - g0, 0.0))
    gn = np.linalg.lstsq(J, viol, rcond=None)[0]           # Gauss-Newton feasibility step
    dirs = [gn, 3.0 * gn, 40.0 * gn, -gn, rng.normal(size=o.n), 1e-3 * rng.normal(size=o.n), 0.0 * gn, 600.0 * gn]
    B = len(dirs)
    x = torch.tensor(np.tile(x0, (B, 1)), device="cuda")
    p = torch.tensor(np.tile(p0, (B, 1)), device="cuda")
    dx = torch.tensor(np.stack(dirs), device="cuda")
    x_new, info = h.line_search(x, p, dx)
    x_new, info = x_new.cpu().numpy(), info.cpu().numpy()
    n_acc = 0
    for b in range(B):
        xr, ir = armijo_line_search(o, dirs[b], x0, p0)
        assert bool(info[b, 0]) == ir["accepted"], (b, info[b], ir)
        assert int(info[b, 2]) == ir["trials"]
        if ir["accepted"]:
            n_acc += 1
            assert info[b, 1] == ir["alpha"]
        assert np.abs(x_new[b] - xr).max() <= 1e-12 * max(1.0, np.abs(xr).max())
        assert abs(info[b, 3] - ir["g_metric"]) <= 1e-9 * max(1.0, ir["g_metric"])
    assert 0 < n_acc


@pytest.mark.parametrize("rn,kind,N,iters", [("b2", "whole_body_rnea", 6, 4), ("b2g", "whole_body_rnea", 5, 3), ("b2", "centroidal_acc", 6, 3),
                                             ("go2", "centroidal_vel", 5, 3), ("b2g", "whole_body_aba", 4, 3), ("b2", "whole_body_acc", 5, 3)])
def test_sqp_iterations_match_oracle(robots, rn, kind, N, iters):
    """Several SQP iterations (sqp_data -> OSQP -> Armijo) with warm-started QP iterates; primal variables and cost
    within the north-star tolerance 1e-6 of the reference path restated by the oracle."""
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(21)
    B = 2
    ocps = [OracleOCP(ora[rn], kind, N) for _ in range(B)]
    xs, ps = zip(*[_nominal_problem(o, rng, k) for o, k in zip(ocps, (0, 33))])
    sqps = [OracleSQP(o) for o in ocps]
    for s in sqps:
        s.init_solver()
    h = Handle(prod[rn], kind, N, max_batch=B)
    x = torch.tensor(np.stack(xs), device="cuda")
    p = torch.tensor(np.stack(ps), device="cuda")
    xr = [np.array(v) for v in xs]
    for it in range(iters):
        x, stats = h.sqp_step(x, p)
        stats = stats.cpu().numpy()
        for b in range(B):
            xr[b], info = sqps[b].solve(xr[b], ps[b])
            assert int(stats[b, 0]) == info["qp_iters"]
            assert bool(stats[b, 2]) == info["accepted"] and int(stats[b, 4]) == info["trials"]
            scale = max(1.0, np.abs(xr[b]).max())
            assert np.abs(x[b].cpu().numpy() - xr[b]).max() <= 1e-6 * scale, (it, b)
            assert abs(stats[b, 5] - info["f"]) <= 1e-6 * max(1.0, abs(info["f"]))
            assert abs(stats[b, 7] - info["violation_max"]) <= 1e-6 * max(1.0, info["violation_max"])
    ms = h.last_phase_ms()
    assert len(ms) == 4 and all(v >= 0 for v in ms)


def test_full_size_sqp_step_is_deterministic(robots, monkeypatch):
    """Bench workload (B2G whole_body_rnea, N=20) on 1332 instances = three full waves of the ADMM kernel: instances are
    independent, so copies of one problem at different batch positions (different CTAs, different co-resident
    neighbours, different waves) must produce bit-identical steps, ADMM iteration counts and statuses -- the
    lock-free panel pipeline of the ADMM kernel must not depend on timing -- and two runs must agree bit for bit."""
    from pino_locoman_b200.handle import Handle
    monkeypatch.delenv("PLM_ADMM_LATENCY_MAX_BATCH", raising=False)
    prod, ora = robots
    rng = np.random.default_rng(11)
    o = OracleOCP(ora["b2g"], "whole_body_rnea", 20)
    B, nbase = 1332, 6
    h = Handle(prod["b2g"], "whole_body_rnea", 20, max_batch=B)
    base = [random_problem(o, rng) for _ in range(nbase)]
    idx = rng.integers(0, nbase, B)
    x = torch.tensor(np.stack([base[i][0] for i in idx]), device="cuda")
    p = torch.tensor(np.stack([base[i][1] for i in idx]), device="cuda")
    h.qp_setup(h.hess_diag(p))
    x1, s1 = h.sqp_step(x, p)
    x1, s1 = x1.clone(), s1.clone()
    first = {int(i): int(np.argmax(idx == i)) for i in set(idx.tolist())}
    ref = torch.tensor([first[int(i)] for i in idx], device="cuda")
    assert torch.equal(x1, x1[ref])
    assert torch.equal(s1, s1[ref])
    assert torch.isfinite(x1).all()
    # a second handle from scratch (fresh workspaces, cold ADMM warm start) reproduces the step bit for bit
    h2 = Handle(prod["b2g"], "whole_body_rnea", 20, max_batch=B)
    h2.qp_setup(h2.hess_diag(p))
    x2, s2 = h2.sqp_step(x, p)
    assert torch.equal(x1, x2) and torch.equal(s1, s2)
    # every instance ran the ADMM loop at least to its first termination check
    assert (s1[:, 0] >= 25).all()


def test_lazy_qp_setup_covers_a_growing_batch(robots):
    """plm_sqp_step sets up the OSQP state (zero iterates, setup-time row scaling) lazily: a later call with a larger
    batch must set up the additional instances too, without touching the warm-started iterates of the earlier ones."""
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(5)
    o = OracleOCP(ora["b2"], "centroidal_acc", 5)
    probs = [_nominal_problem(o, rng, k) for k in (0, 11, 23, 37)]
    x = torch.tensor(np.stack([q[0] for q in probs]), device="cuda")
    p = torch.tensor(np.stack([q[1] for q in probs]), device="cuda")
    full = Handle(prod["b2"], "centroidal_acc", 5, max_batch=4)
    xf1, sf1 = full.sqp_step(x, p)
    xf2, sf2 = full.sqp_step(xf1, p)
    grow = Handle(prod["b2"], "centroidal_acc", 5, max_batch=4)
    xg1, sg1 = grow.sqp_step(x[:2].contiguous(), p[:2].contiguous())
    assert torch.equal(xg1, xf1[:2]) and torch.equal(sg1, sf1[:2])
    # batch grows to 4: instances 0, 1 continue warm-started (second iteration), instances 2, 3 run their first one
    xin = torch.cat([xg1, x[2:]])
    xg2, sg2 = grow.sqp_step(xin, p)
    assert torch.equal(xg2[:2], xf2[:2]) and torch.equal(sg2[:2], sf2[:2])
    assert torch.equal(xg2[2:], xf1[2:]) and torch.equal(sg2[2:], sf1[2:])
