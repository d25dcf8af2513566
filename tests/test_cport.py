"""The compiled CPU port (oracle/cport, bench.py's CPU baseline) against the numpy oracle: same rows / Jacobian, same
OSQP iterate sequence (scaling, iteration counts, statuses), same Armijo decisions, SQP step <= 1e-6."""
import numpy as np
import pytest

from emu_util import random_problem
from oracle.model import OracleRobot
from oracle.ocp import OracleOCP
from oracle.sqp import OracleSQP


@pytest.fixture(scope="module")
def cport():
    from oracle import cport as cp
    cp.build()
    return cp


def _nominal(o, rng, k):
    o.set_time_params(0.01, 0.08)
    o.set_swing_params(0.07, [0.1, -0.2])
    o.set_tracking_targets([0.2, 0, 0, 0, 0, 0], rng.uniform(-5, 5, 3), rng.uniform(-0.05, 0.05, 3))
    o.update_initial_state(o.x_nom)
    o.update_gait_sequence(k * 0.01)
    if o.kind == "whole_body_rnea":
        o.update_previous_torques(np.zeros(o.nj))
    return o.initial_guess(), o.p_vector()


@pytest.mark.parametrize("rn,kind,N", [("b2g", "whole_body_rnea", 5), ("go2", "centroidal_vel", 4), ("b2g", "whole_body_aba", 3)])
def test_cport_rows_and_jacobian(cport, rn, kind, N):
    rng = np.random.default_rng(7)
    o = OracleOCP(OracleRobot(rn), kind, N)
    x, p = random_problem(o, rng)
    s = cport.CPortSQP(o)
    g, Jv = s.node_eval(x, p)
    g_ref, _, _ = o.g_data(x, p)
    J_ref = o.jac_g(x, p)
    rows = np.zeros(s.nnz, dtype=np.int32)
    cols = np.zeros(s.nnz, dtype=np.int32)
    s.lib.cport_pattern(s.h, rows.ctypes.data_as(cport._ip), cols.ctypes.data_as(cport._ip))
    Jd = np.zeros((o.m, o.n))
    Jd[rows, cols] = Jv
    assert np.abs(g - g_ref).max() <= 1e-9 * max(1.0, np.abs(g_ref).max())
    assert np.abs(Jd - J_ref).max() <= 1e-9 * np.abs(J_ref).max()


@pytest.mark.parametrize("rn,kind,N,iters,kw", [("b2g", "whole_body_rnea", 5, 3, {}), ("b2", "centroidal_acc", 5, 2, {}), ("go2", "centroidal_vel", 4, 2, {}),
                                                ("b2", "centroidal_acc", 4, 2, {"include_base": False}),
                                                ("b2", "whole_body_rnea", 5, 3, {"include_acc": False})])
def test_cport_sqp_iterations_match_numpy_oracle(cport, rn, kind, N, iters, kw):
    rng = np.random.default_rng(8)
    o = OracleOCP(OracleRobot(rn), kind, N, **kw)
    x, p = _nominal(o, rng, 17)
    ref = OracleSQP(o)
    ref.init_solver()
    s = cport.CPortSQP(o)
    s.init_solver(p)
    xr, xc = np.array(x), np.array(x)
    for it in range(iters):
        xr, ir = ref.solve(xr, p)
        xc, ic = s.solve(xc, p)
        D, E, c = s.scaling()
        assert np.abs(D - ref.osqp.D).max() <= 1e-10 * np.abs(ref.osqp.D).max()
        assert np.abs(E - ref.osqp.E).max() <= 1e-10 * np.abs(ref.osqp.E).max()
        assert abs(c - ref.osqp.c) <= 1e-10 * ref.osqp.c
        assert ic["qp_iters"] == ir["qp_iters"] and ic["qp_status"] == ir["qp_status"], (it, ic["qp_iters"], ir["qp_iters"])
        assert ic["accepted"] == ir["accepted"] and ic["trials"] == ir["trials"]
        assert np.abs(ic["sol_dx"] - ir["sol_dx"]).max() <= 1e-6 * max(1.0, np.abs(ir["sol_dx"]).max())
        assert np.abs(xc - xr).max() <= 1e-6 * max(1.0, np.abs(xr).max())
        assert abs(ic["f"] - ir["f"]) <= 1e-6 * max(1.0, abs(ir["f"]))
        assert abs(ic["violation_max"] - ir["violation_max"]) <= 1e-6 * max(1.0, ir["violation_max"])
        xq, zq, yq = s.iterates()
        assert np.abs(xq - ref.osqp.x).max() <= 1e-6 * max(1.0, np.abs(ref.osqp.x).max())
        assert np.abs(yq - ref.osqp.y).max() <= 1e-6 * max(1.0, np.abs(ref.osqp.y).max())
