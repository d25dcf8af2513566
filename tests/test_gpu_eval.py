"""GPU parity: sqp_data / g_data / f_data / hess_diag through the C ABI against the oracle (<= 1e-9 relative)."""
import numpy as np
import pytest

from emu_util import random_problem
from oracle.ocp import OracleOCP

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL = 1e-9

CASES = [("b2g", "whole_body_rnea", 6, 3), ("b2", "whole_body_rnea", 5, 2), ("b2", "centroidal_acc", 5, 3), ("b2g", "whole_body_acc", 4, 2),
         ("go2", "centroidal_vel", 5, 3), ("b2g", "whole_body_aba", 4, 2), ("go2", "whole_body_aba", 4, 2)]


@pytest.mark.parametrize("rn,kind,N,B", CASES)
def test_sqp_data_matches_oracle(robots, rn, kind, N, B):
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(11)
    o = OracleOCP(ora[rn], kind, N)
    h = Handle(prod[rn], kind, N, max_batch=B)
    assert (h.n, h.m, h.np) == (o.n, o.m, o.np_)
    xs, ps = zip(*[random_problem(o, rng) for _ in range(B)])
    x = torch.tensor(np.stack(xs), device="cuda")
    p = torch.tensor(np.stack(ps), device="cuda")
    grad, J, g, lbg, ubg = h.sqp_data(x, p)
    f, grad2 = h.f_data(x, p)
    g2, _, _ = h.g_data(x, p, bounds=False)
    hd = h.hess_diag(p)
    Jd = h.jac_dense(J).cpu().numpy()
    assert torch.equal(g, g2) and torch.equal(grad, grad2)
    for b in range(B):
        g_ref, lb_ref, ub_ref = o.g_data(xs[b], ps[b])
        J_ref = o.jac_g(xs[b], ps[b])
        f_ref, grad_ref = o.f_data(xs[b], ps[b])
        assert np.abs(g[b].cpu().numpy() - g_ref).max() <= TOL * max(1.0, np.abs(g_ref).max())
        assert np.abs(Jd[b] - J_ref).max() <= TOL * np.abs(J_ref).max()
        assert np.array_equal(lbg[b].cpu().numpy(), lb_ref) and np.array_equal(ubg[b].cpu().numpy(), ub_ref)
        assert abs(f[b].item() - f_ref) <= TOL * abs(f_ref)
        assert np.abs(grad[b].cpu().numpy() - grad_ref).max() <= TOL * np.abs(grad_ref).max()
        assert np.abs(hd[b].cpu().numpy() - o.hess_diag(ps[b])).max() == 0


def test_full_size_properties(robots):
    """BASELINE config 4 size (B2 centroidal_acc, N=20) on a larger batch: size-independent properties --
    instances are independent (permutation equivariance) and residual-only rows equal the Jacobian-mode rows."""
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(5)
    o = OracleOCP(ora["b2"], "centroidal_acc", 20)
    B = 256
    h = Handle(prod["b2"], "centroidal_acc", 20, max_batch=B)
    base = [random_problem(o, rng) for _ in range(8)]
    idx = rng.integers(0, 8, B)
    x = torch.tensor(np.stack([base[i][0] for i in idx]), device="cuda")
    p = torch.tensor(np.stack([base[i][1] for i in idx]), device="cuda")
    grad, J, g, _, _ = h.sqp_data(x, p, bounds=False)
    g2, _, _ = h.g_data(x, p, bounds=False)
    assert torch.equal(g, g2)
    first = {int(i): int(np.argmax(idx == i)) for i in set(idx.tolist())}
    for b in range(B):
        assert torch.equal(g[b], g[first[int(idx[b])]]) and torch.equal(J[b], J[first[int(idx[b])]])
    g_ref, _, _ = o.g_data(*base[int(idx[0])])
    assert np.abs(g[0].cpu().numpy() - g_ref).max() <= TOL * max(1.0, np.abs(g_ref).max())


def test_input_validation(robots):
    from pino_locoman_b200.handle import Handle
    prod, _ = robots
    h = Handle(prod["b2"], "whole_body_rnea", 4, max_batch=2)
    with pytest.raises(ValueError):
        h.g_data(torch.zeros(1, h.n + 1, dtype=torch.float64, device="cuda"), torch.zeros(1, h.np, dtype=torch.float64, device="cuda"))
    with pytest.raises(TypeError):
        h.g_data(torch.zeros(1, h.n, dtype=torch.float32, device="cuda"), torch.zeros(1, h.np, dtype=torch.float64, device="cuda"))
    with pytest.raises(RuntimeError):
        h.g_data(torch.zeros(3, h.n, dtype=torch.float64, device="cuda"), torch.zeros(3, h.np, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        Handle(prod["b2"], "whole_body_foo", 4, max_batch=2)


def test_handles_of_different_sizes_coexist(robots):
    """Kernel attributes (dynamic shared memory) are per function, not per handle: a small handle created after a large one
    must not break the large one."""
    from pino_locoman_b200.handle import Handle
    prod, ora = robots
    rng = np.random.default_rng(2)
    big = Handle(prod["b2g"], "whole_body_rnea", 6, max_batch=2)
    o = OracleOCP(ora["b2g"], "whole_body_rnea", 6)
    x0, p0 = random_problem(o, rng)
    x = torch.tensor(np.tile(x0, (2, 1)), device="cuda")
    p = torch.tensor(np.tile(p0, (2, 1)), device="cuda")
    x1, _ = big.sqp_step(x, p)
    small = Handle(prod["go2"], "centroidal_vel", 4, max_batch=1)
    assert small.n < big.n
    x2, _ = big.sqp_step(x, p)      # same inputs, but the QP iterates are warm now: only check that it runs and is finite
    torch.cuda.synchronize()
    assert torch.isfinite(x1).all() and torch.isfinite(x2).all()
