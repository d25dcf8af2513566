"""Host logic of the plugin surface (no GPU): parameter-vector assembly, initial guess and warm start of the batched
OCP classes against the oracle's restatement of the reference, factory / error behaviour."""
import numpy as np
import pytest

from oracle.ocp import OracleOCP
from pino_locoman_b200 import OCP_ARGS
from pino_locoman_b200.optimization import make_ocp

CASES = [("go2", "centroidal_vel"), ("b2", "whole_body_rnea"), ("b2g", "whole_body_aba"), ("b2", "centroidal_acc"),
         ("b2g", "whole_body_rnea"), ("b2g", "whole_body_acc")]


def _configure(ocp, x_init, t, ext, arm):
    ocp.set_time_params(0.01, 0.08)
    ocp.set_swing_params(0.07, [0.1, -0.2])
    ocp.set_tracking_targets(np.array([0.2, 0, 0, 0, 0, 0]), ext, arm)
    ocp.update_initial_state(x_init)
    ocp.update_gait_sequence(t)


@pytest.mark.parametrize("rn,kind", CASES)
def test_parameter_vector_and_guess_match_oracle(robots, rn, kind):
    prod, ora = robots
    N, B = 20, 3
    ocp = make_ocp(dynamics=kind, default_args=OCP_ARGS[kind], robot=prod[rn], nodes=N, solver="osqp", batch=B, device="layout")
    rng = np.random.default_rng(0)
    ts = np.array([0.0, 0.17, 0.79])
    x_inits = np.stack([ocp.x_nom + 0.01 * rng.normal(size=ocp.x_nom.size) for _ in range(B)])
    ext, arm = np.array([1.0, -2.0, 3.0]), np.array([0.05, 0.0, -0.02])
    _configure(ocp, x_inits, ts, ext, arm)
    if kind == "whole_body_rnea":
        ocp.update_previous_torques(np.arange(ocp.nj) * 0.1)
    for b in range(B):
        o = OracleOCP(ora[rn], kind, N)
        assert (ocp.n, ocp.m, ocp.handle.np) == (o.n, o.m, o.np_)
        _configure(o, x_inits[b], ts[b], ext, arm)
        if kind == "whole_body_rnea":
            o.update_previous_torques(np.arange(o.nj) * 0.1)
        assert np.array_equal(ocp._p[b], o.p_vector())            # bit-exact, incl. the gait schedules
        # forces in the guess carry the total mass: the two independent loaders sum link masses in different orders
        assert np.allclose(ocp.initial_guess()[b], o.initial_guess(), rtol=1e-14, atol=0)
        assert np.array_equal(np.asarray(ocp.dts), np.asarray(o.dts()))
    # warm start from a fake previous solution
    sol = rng.normal(size=(B, ocp.n))
    ocp.retract_stacked_sol(sol, retract_all=False)
    ocp.warm_start()
    for b in range(B):
        o = OracleOCP(ora[rn], kind, N)
        _configure(o, x_inits[b], ts[b], ext, arm)
        o.retract_stacked_sol(sol[b])
        assert np.allclose(ocp._x0[b], o.warm_start(), rtol=1e-14, atol=0)
    assert len(ocp.q_sol) == 1 and ocp.q_sol[0].shape == (B, ocp.nq)


def test_factory_errors(robots):
    prod, _ = robots
    with pytest.raises(ValueError, match="Unknown dynamics type"):
        make_ocp(dynamics="whole_body_foo", default_args={}, robot=prod["b2"], nodes=10, solver="osqp")
    ocp = make_ocp(dynamics="centroidal_acc", default_args=OCP_ARGS["centroidal_acc"], robot=prod["b2"], nodes=10, solver="ipopt",
                   device="layout")
    with pytest.raises(ValueError, match="not supported"):
        ocp.init_solver()


@pytest.mark.parametrize("rn,kind,kw", [("b2", "centroidal_acc", {"include_base": False}), ("b2g", "whole_body_acc", {"include_base": False}),
                                        ("go2", "centroidal_vel", {"include_base": False}), ("b2g", "centroidal_vel", {"include_base": False}),
                                        ("b2g", "whole_body_rnea", {"tau_nodes": 3, "include_acc": False})])
def test_layout_of_the_reduced_input_variants(robots, rn, kind, kw):
    """include_base=False / include_acc=False: inputs (v_j | a_j, f) / (f, tau_j), no gap rows / no dv
    integrator rows; sizes and initial guess equal the oracle's."""
    from oracle.ocp import OracleOCP
    prod, ora = robots
    ocp = make_ocp(dynamics=kind, default_args=kw, robot=prod[rn], nodes=8, solver="osqp", device="layout")
    o = OracleOCP(ora[rn], kind, 8, **kw)
    assert (ocp.n, ocp.m, ocp.handle.np) == (o.n, o.m, o.np_)
    assert list(ocp.handle.x_off[:9]) == [int(v) for v in o.x_off]
    assert ocp.nu_opt[0] == o.nu[0] and ocp.nu_opt[-1] == o.nu[-1]
    assert np.allclose(ocp.initial_guess()[0], o.initial_guess(), rtol=1e-12, atol=0)      # (two independent URDF loaders: masses agree to 1e-12)
    assert np.array_equal(ocp._get("R_diag")[0], o.params["R_diag"])


def test_retract_integrates_on_the_manifold(robots):
    from oracle.dynamics import DynamicsWholeBodyTorque
    prod, ora = robots
    ocp = make_ocp(dynamics="whole_body_rnea", default_args=OCP_ARGS["whole_body_rnea"], robot=prod["b2g"], nodes=6, solver="osqp",
                   batch=2, device="layout")
    rng = np.random.default_rng(1)
    x = np.stack([ocp.x_nom, ocp.x_nom])
    x[1, 3:7] = [0.1, -0.2, 0.3, 0.9]
    x[1, 3:7] /= np.linalg.norm(x[1, 3:7])
    dx = rng.normal(size=(2, ocp.ndx_opt)) * 0.4
    dx[0, 3:6] = 0
    got = ocp.state_integrate(x, dx)
    dyn = DynamicsWholeBodyTorque(ora["b2g"].model, ora["b2g"].mass, ora["b2g"].foot_frames)
    for b in range(2):
        assert np.abs(got[b] - dyn.state_integrate()(x[b], dx[b])).max() < 1e-12


@pytest.mark.parametrize("rn,kind,N", [("b2g", "whole_body_rnea", 20), ("b2", "centroidal_acc", 7), ("go2", "centroidal_vel", 5),
                                       ("b2g", "whole_body_aba", 6)])
def test_stored_qp_factor_size(robots, rn, kind, N):
    """kkt_factor_doubles (the nnz(F) of SURVEY 8d that bench.py reports) = inverse stage blocks S_i^-1 stored by cyclic
    diagonals (stages 0..N: s/2 + 1 rows of s doubles, the size of the packed triangle for odd s, 16-byte aligned) plus the
    back-substitution blocks B_i = S_i^-1 G_i^T (ndx columns of the stage size rounded up to even)."""
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    prod, _ = robots
    ocp = make_ocp(dynamics=kind, default_args=OCP_ARGS[kind], robot=prod[rn], nodes=N, solver="osqp", batch=1, device="layout")
    h = ocp.handle
    ndx = h.ndx
    sizes = [ndx + nu for nu in h.nu] + [ndx]
    packed = sum(((s // 2 + 1) * s + 1) & ~1 for s in sizes)
    back = sum(ndx * ((s + 1) & ~1) for s in sizes[:-1])
    assert h.dims.kkt_factor_doubles == packed + back
    if (rn, kind, N) == ("b2g", "whole_body_rnea", 20):
        assert (h.n, h.m) == (1842, 2665) and h.dims.kkt_factor_doubles == 170046


@pytest.mark.parametrize("rn,kind", [("go2", "centroidal_vel"), ("b2g", "whole_body_aba"), ("b2", "centroidal_acc"), ("b2g", "whole_body_rnea"),
                                     ("b2", "whole_body_acc")])
def test_bench_synthetic_inputs_cover_every_formulation(robots, rn, kind):
    """bench.py's synthetic states (SURVEY 8d distributions) for the legs on all BASELINE configs: finite, contact-masked forces."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from pino_locoman_b200 import OCP_ARGS
    from pino_locoman_b200.optimization import make_ocp
    prod, _ = robots
    B, N = 3, 20
    ocp = make_ocp(dynamics=kind, default_args=OCP_ARGS[kind], robot=prod[rn], nodes=N, solver="osqp", batch=B, device="layout")
    x, p = bench.synthetic_inputs(prod[rn], ocp, B, 0)
    assert x.shape == (B, ocp.handle.n) and p.shape == (B, ocp.handle.np)
    assert np.isfinite(x).all() and np.isfinite(p).all()
    contact = ocp._get("contact_schedule").reshape(B, N, 4)
    lead = ocp._lead()
    for i in range(N):
        o = ocp.handle.x_off[i] + ocp.ndx_opt + lead
        f = x[:, o:o + 12].reshape(B, 4, 3)
        assert np.all(f[contact[:, i] == 0] == 0.0)           # swing feet carry no force
        assert np.all(f[..., 2] >= 0.0)
    x2, p2 = bench.synthetic_inputs(prod[rn], ocp, B, 0)      # seeded: identical bits on every call (CPU and GPU arms)
    assert np.array_equal(x, x2) and np.array_equal(p, p2)


def test_cyclic_diagonal_storage_of_the_inverse_blocks():
    """S^-1 of a stage is stored by cyclic diagonals M[j][k] = S^-1[k][(k + j) mod s] (plm_qp_types.h): every pair (r >= c)
    has one slot, the array has the size of the packed triangle (plus s/2 zeros for even s), and the mask-free product of
    the ADMM kernel's sym_panel reproduces S^-1 v."""
    from emu_util import build_emu
    lib = build_emu()
    rng = np.random.default_rng(3)
    for s in (1, 2, 5, 48, 87, 105, 128):
        rows = lib.emu_sinv_rows(s)
        assert rows == s // 2 + 1
        S = rng.standard_normal((s, s))
        S = S + S.T
        M = np.zeros(rows * s)
        hit = np.zeros(rows * s, dtype=int)
        for r in range(s):
            for c in range(r + 1):
                k = lib.emu_sinv_index(s, r, c)
                assert 0 <= k < rows * s
                hit[k] += 1
                M[k] = S[r, c]
        assert hit.max() == 1
        free = np.flatnonzero(hit == 0)
        if s % 2:
            assert free.size == 0 and rows * s == s * (s + 1) // 2
        else:      # second half of the last diagonal: stored as zeros
            assert np.array_equal(free, (s // 2) * s + s // 2 + np.arange(s // 2))
        M = M.reshape(rows, s)
        v = rng.standard_normal(s)
        k = np.arange(s)
        out = M[0] * v
        for j in range(1, rows):
            out += M[j] * v[(k + j) % s] + M[j][(k - j) % s] * v[(k - j) % s]
        assert np.abs(out - S @ v).max() < 1e-12 * max(1.0, np.abs(S @ v).max())


@pytest.mark.parametrize("rn,kind,N", [("b2g", "whole_body_rnea", 20), ("go2", "centroidal_vel", 5), ("b2g", "whole_body_aba", 6)])
def test_admm_panel_schedule_covers_every_stored_double_once(robots, rn, kind, N):
    """The host-built panel schedules of one ADMM iteration (throughput and latency kernels): bulk copies are 16-byte
    aligned, fit the panel buffer, and the forward / backward steps together visit every row of every S_i^-1 array and
    every column of every B_i exactly once."""
    import ctypes
    from emu_util import Emu
    prod, _ = robots
    e = Emu(prod[rn], kind, N)
    fac_off = (ctypes.c_int * (N + 2))()
    bk_off = (ctypes.c_int * (N + 1))()
    pd = (ctypes.c_int * 3)()
    e.lib.emu_qp_factor_offsets(e.h, fac_off, bk_off, pd)
    for latency in (0, 1):
        buf = (ctypes.c_int * (8 * 4096))()
        ns = e.lib.emu_qp_schedule(e.h, latency, buf, len(buf))
        sched = np.array(buf[:8 * ns]).reshape(ns, 8)
        seen_rows = {i: [] for i in range(N + 1)}
        seen_cols = {i: [] for i in range(N)}
        seen_coup = {i: [] for i in range(1, N + 1)}
        for off, length, a, b, i, flags, start, ss in sched:
            s = ss & 255
            assert off % 2 == 0 and length % 2 == 0 and 0 < length <= pd[latency]
            if flags & 8:      # coupling panel (dense integrator rows): rows [a, b) of the block of node i - 1, leading dimension `start`
                assert kind in ("whole_body_aba", "centroidal_vel") and not (flags & 7) and i >= 1
                assert off == ((i - 1) * e.ndx + a) * start and length == (b - a) * start and start % 2 == 0
                assert bool(flags & 16) == (a == 0)
                seen_coup[i] += list(range(a, b))
                continue
            if flags & 1:      # backward: columns [a, b) of B_i, stride `start`
                assert off == bk_off[i] + a * start and length == (b - a) * start and start == (s + 1) & ~1
                seen_cols[i] += list(range(a, b))
            else:
                assert off == fac_off[i] + start and start <= a * s and a * s - start <= 1 and start + length >= b * s
                assert off + length <= fac_off[i + 1]
                seen_rows[i] += list(range(a, b))
        for i in range(N + 1):
            rows = sorted(seen_rows[i])
            assert rows == list(range(len(rows)))
        for i in range(N):
            assert sorted(seen_cols[i]) == list(range(e.ndx))
        if kind in ("whole_body_aba", "centroidal_vel"):      # every stage after the first is preceded by its coupling panels
            for i in range(1, N + 1):
                assert seen_coup[i] == list(range(e.ndx))
        # the forward steps of a stage cover rows 0 .. s/2 of its array
        for off, length, a, b, i, flags, start, ss in sched:
            if not (flags & 9) and (flags & 4):
                assert b == (ss & 255) // 2 + 1
