"""Size-independent properties at BASELINE.json's full per-GPU size (8192 B2G whole_body_rnea instances, N = 20, the bench
workload with its synthetic state distributions), plus oracle spot checks of instances picked from the full batch."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

B_FULL = 8192


@pytest.fixture(scope="module")
def full_batch():
    import bench
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    from pino_locoman_b200.utils.robot import B2G
    robot = B2G()
    robot.set_gait_sequence("trot", 0.8)
    ocp = make_ocp(dynamics=bench.DYNAMICS, default_args=OCP_ARGS[bench.DYNAMICS], robot=robot, nodes=bench.NODES, solver="osqp",
                   batch=B_FULL, device="cuda:0")
    x_host, p_host = bench.synthetic_inputs(robot, ocp, B_FULL, 0)
    ocp.init_solver()
    return ocp, x_host, p_host, torch.from_numpy(x_host).cuda(), torch.from_numpy(p_host).cuda()


def test_jacobian_is_the_derivative_of_the_rows_for_every_instance(full_batch):
    """Directional central differences of g along a random direction against J d, for all 8192 instances at once."""
    ocp, _, _, x, p = full_batch
    h = ocp.handle
    gen = torch.Generator(device="cuda").manual_seed(5)
    d = torch.randn(x.shape, dtype=torch.float64, device="cuda", generator=gen)
    d[:, :h.ndx] = 0.0
    _, J, g, _, _ = h.sqp_data(x, p)
    eps = 1e-6
    gp, _, _ = h.g_data(x + eps * d, p, bounds=False)
    gm, _, _ = h.g_data(x - eps * d, p, bounds=False)
    fd = (gp - gm) / (2 * eps)
    rows = torch.as_tensor(h.jac_rows, dtype=torch.long, device="cuda")
    cols = torch.as_tensor(h.jac_cols, dtype=torch.long, device="cuda")
    Jd = torch.zeros_like(g)
    Jd.index_add_(1, rows, J * d[:, cols])
    scale = torch.maximum(fd.abs().amax(1), torch.ones(B_FULL, dtype=torch.float64, device="cuda"))
    err = ((fd - Jd).abs().amax(1) / scale).max().item()
    assert err < 1e-6, err            # (second-order truncation of the central difference: eps^2 |g'''| ~ 1e-8 relative)
    assert torch.isfinite(J).all() and torch.isfinite(g).all()


def test_oracle_spot_checks_inside_the_full_batch(full_batch, robots):
    """Instances picked from the full batch (first, last, middle of a wave) against the oracle: rows, Jacobian, gradient."""
    from oracle.ocp import OracleOCP
    ocp, x_host, p_host, x, p = full_batch
    _, ora = robots
    h = ocp.handle
    o = OracleOCP(ora["b2g"], "whole_body_rnea", 20)
    grad, J, g, lbg, ubg = h.sqp_data(x, p)
    rows, cols = h.jac_rows, h.jac_cols
    for b in (0, 4097, B_FULL - 1):
        g_ref, lb_ref, ub_ref = o.g_data(x_host[b], p_host[b])
        J_ref = o.jac_g(x_host[b], p_host[b])
        _, grad_ref = o.f_data(x_host[b], p_host[b])
        Jd = np.zeros((h.m, h.n))
        Jd[rows, cols] = J[b].cpu().numpy()
        assert np.abs(g[b].cpu().numpy() - g_ref).max() <= 1e-9 * max(1.0, np.abs(g_ref).max())
        assert np.abs(Jd - J_ref).max() <= 1e-9 * np.abs(J_ref).max()
        assert np.abs(grad[b].cpu().numpy() - grad_ref).max() <= 1e-9 * max(1.0, np.abs(grad_ref).max())
        assert np.array_equal(lbg[b].cpu().numpy(), lb_ref) and np.array_equal(ubg[b].cpu().numpy(), ub_ref)


def test_full_batch_sqp_step_properties(full_batch, robots):
    """One SQP iteration of all 8192 instances: every accepted step is x + alpha dx with alpha a power of 1/2, the reported
    cost equals f_data at the new point, the reported violation equals the one recomputed from g_data, the QP step of a
    'solved' instance satisfies the linearised constraints to OSQP's tolerance, and three instances match the oracle."""
    from oracle.ocp import OracleOCP
    from oracle.sqp import OracleSQP
    ocp, x_host, p_host, x, p = full_batch
    _, ora = robots
    h = ocp.handle
    x_new, stats = h.sqp_step(x, p)
    st = stats.cpu().numpy()
    acc = st[:, 2] != 0
    assert acc.mean() > 0.9
    # (a rejected line search reports f / g_metric of its last trial, as the reference's overwrite quirk does: with a
    # diverging QP step these may be huge or non-finite; the returned point is the current one)
    assert torch.isfinite(x_new).all() and np.isfinite(st[acc]).all(), np.argwhere(~np.isfinite(st))[:5]
    assert np.isfinite(st[:, [0, 1, 2, 4, 7]]).all()
    alpha = st[:, 3]
    assert np.all(np.log2(alpha[acc]) == np.round(np.log2(alpha[acc]))) and np.all((alpha[acc] <= 1.0) & (alpha[acc] > 1e-4))
    assert torch.equal(x_new[torch.as_tensor(~acc, device="cuda")], x[torch.as_tensor(~acc, device="cuda")])     # rejected: current_x returned
    f_new, _ = h.f_data(x_new, p)
    assert np.abs(f_new.cpu().numpy()[acc] - st[acc, 5]).max() <= 1e-9 * np.abs(st[acc, 5]).max()
    g_new, lbg, ubg = h.g_data(x_new, p)
    viol = torch.maximum(torch.clamp(lbg - g_new, min=0), torch.clamp(g_new - ubg, min=0)).amax(1).cpu().numpy()
    assert np.abs(viol - st[:, 7]).max() <= 1e-9 * max(1.0, viol.max())
    # OSQP statuses: solved / solved inaccurate / max iterations; a few random states give an (inaccurate) infeasibility
    # certificate: OSQP then returns a NaN step and the line search keeps the current point, as in the reference
    status = st[:, 1].astype(int)
    assert set(np.unique(status)) <= {1, 2, -2, 3, -3, 4, -4}
    infeasible = np.isin(status, (3, -3, 4, -4))
    assert infeasible.mean() < 0.01 and not acc[infeasible].any()
    assert np.all((st[:, 0] >= 25) & (st[:, 0] <= 100) & (st[:, 0] % 25 == 0))
    for b in (0, 5000, B_FULL - 1):
        o = OracleOCP(ora["b2g"], "whole_body_rnea", 20)
        for name, (off, sz) in o.p_layout.items():
            o.params[name][:] = p_host[b, off:off + sz]
        s = OracleSQP(o)
        s.init_solver()
        x_ref, info = s.solve(x_host[b].copy(), p_host[b])
        assert int(st[b, 0]) == info["qp_iters"] and bool(st[b, 2]) == info["accepted"] and int(st[b, 4]) == info["trials"]
        assert np.abs(x_new[b].cpu().numpy() - x_ref).max() <= 1e-6 * max(1.0, np.abs(x_ref).max())
        assert abs(st[b, 5] - info["f"]) <= 1e-6 * max(1.0, abs(info["f"]))
