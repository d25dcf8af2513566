"""Kernel mathematics (device code compiled for the host, lanes looped) against the oracle: residual rows and
analytic Jacobian blocks of every formulation, <= 1e-9 relative (north-star tolerance)."""
import numpy as np
import pytest

from emu_util import Emu, random_problem
from oracle.ocp import OracleOCP

TOL = 1e-9
CASES = [("b2g", "whole_body_rnea", 5), ("b2", "whole_body_rnea", 4), ("go2", "whole_body_rnea", 4), ("b2", "centroidal_acc", 4),
         ("b2g", "centroidal_acc", 3), ("b2g", "whole_body_acc", 3), ("go2", "centroidal_vel", 4), ("b2g", "centroidal_vel", 3),
         ("b2g", "whole_body_aba", 3), ("b2", "whole_body_aba", 3)]


@pytest.mark.parametrize("rn,kind,N", CASES)
def test_rows_and_jacobian_match_oracle(robots, rn, kind, N):
    prod, ora = robots
    rng = np.random.default_rng(hash((rn, kind)) % 2**32)
    o = OracleOCP(ora[rn], kind, N)
    e = Emu(prod[rn], kind, N)
    assert (e.n, e.m, e.np_) == (o.n, o.m, o.np_)
    for trial in range(2):
        x, p = random_problem(o, rng)
        if trial == 1:   # first SQP iterate: DX = 0 exactly (series branch of Exp / Jr)
            for i in range(N + 1):
                x[o.x_off[i]:o.x_off[i] + o.ndx] = 0
        g_ref, _, _ = o.g_data(x, p)
        J_ref = o.jac_g(x, p)
        g, Jv = e.eval(x, p)
        assert np.abs(g[0] - g_ref).max() <= TOL * max(1.0, np.abs(g_ref).max())
        assert np.abs(e.dense(Jv[0]) - J_ref).max() <= TOL * np.abs(J_ref).max()
        g2, _ = e.eval(x, p, want_jac=False)     # residual-only mode (line search) gives the same rows
        assert np.array_equal(g2, g)


def test_pattern_is_sorted_csr(robots):
    e = Emu(robots[0]["b2g"], "whole_body_rnea", 6)
    key = e.rows.astype(np.int64) * e.n + e.cols
    assert np.all(np.diff(key) > 0)
    assert e.rows.max() == e.m - 1 and e.cols.max() == e.n - 1


def test_rnea_tau_nodes_variants(robots):
    prod, ora = robots
    rng = np.random.default_rng(5)
    for tau_nodes in (1, 2, 4):
        o = OracleOCP(ora["b2"], "whole_body_rnea", 4, tau_nodes=tau_nodes)
        e = Emu(prod["b2"], "whole_body_rnea", 4, tau_nodes=tau_nodes)
        assert (e.n, e.m) == (o.n, o.m)
        x, p = random_problem(o, rng)
        g, Jv = e.eval(x, p)
        assert np.abs(g[0] - o.g_data(x, p)[0]).max() < 1e-9 * 1e3
        assert np.abs(e.dense(Jv[0]) - o.jac_g(x, p)).max() < 1e-9 * 1e3
