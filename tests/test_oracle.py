"""Oracle pinned against the reference's own gait code (golden vectors) and against algebraic identities."""
import json
import os

import numpy as np
import pytest

from oracle import rbd, spatial as sp
from oracle.dynamics import DynamicsCentroidalAcc, DynamicsWholeBodyTorque
from oracle.gait import GaitSequence, get_spline_vel_z, opti_dts
from oracle.model import OracleRobot
from oracle.ocp import OracleOCP

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gait_golden.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def _check_schedules(gold, make_seq, make_dts):
    for case in gold["schedules"]:
        dts = make_dts(case["dt_min"], case["dt_max"], case["nodes"])
        assert [float(d).hex() for d in dts] == case["dts"]
        seq = make_seq(case["gait"], case["period"])
        assert seq.n_contacts == case["n_contacts"] and float(seq.swing_period).hex() == case["swing_period"]
        contact, swing = seq.get_gait_schedule(case["k"] * case["dt_min"], dts, case["nodes"])
        assert contact.astype(int).tolist() == case["contact"]
        assert [[float(s).hex() for s in row] for row in swing] == case["swing"]


def test_oracle_gait_bit_exact_vs_reference(gold):
    _check_schedules(gold, GaitSequence, opti_dts)


def test_product_gait_bit_exact_vs_reference(gold):
    from pino_locoman_b200.utils.gait_sequence import GaitSequence as G, horizon_dts
    _check_schedules(gold, G, horizon_dts)
    # batched start times give the same bits as one call per start time
    seq = G("trot", 0.8)
    dts = horizon_dts(0.01, 0.08, 20)
    ks = np.arange(0, 80)
    cb, sb = seq.get_gait_schedule(ks * 0.01, dts, 20)
    for k in ks:
        c1, s1 = seq.get_gait_schedule(k * 0.01, dts, 20)
        assert np.array_equal(cb[k], c1) and np.array_equal(sb[k], s1)


def test_spline_bit_exact_vs_reference(gold):
    from pino_locoman_b200.utils.gait_sequence import get_spline_vel_z as prod_spline
    for s in gold["splines"]:
        args = (s["phase"], s["swing_period"], s["h_max"], s["v_liftoff"], s["v_touchdown"])
        assert float(get_spline_vel_z(*args)).hex() == s["vel_z"]
        assert float(prod_spline(*args)).hex() == s["vel_z"]


def test_unknown_gait_raises():
    from pino_locoman_b200.utils.gait_sequence import GaitSequence as G
    with pytest.raises(ValueError):
        GaitSequence("gallop", 0.5)
    with pytest.raises(ValueError):
        G("gallop", 0.5)


@pytest.mark.parametrize("name,mass,root_mass,nq,nv,nf", [("go2", 16.087, 7.279, 19, 18, 12), ("b2", 72.5803, 39.4371, 19, 18, 12),
                                                          ("b2g", 77.26826983, 40.210575, 25, 24, 15)])
def test_model_constants(name, mass, root_mass, nq, nv, nf):
    """Published constants of SURVEY.md 2.4 (total / merged root mass, dims, joint order)."""
    r = OracleRobot(name)
    m = r.model
    assert (m.nq, m.nv, r.nf) == (nq, nv, nf)
    assert abs(m.total_mass - mass) < 1e-6 and abs(m.mass[1] - root_mass) < 1e-6
    legs = [f"{l}_{j}_joint" for l in ("FL", "FR", "RL", "RR") for j in ("hip", "thigh", "calf")]
    arm = [f"joint{i}" for i in range(1, 7)] if name == "b2g" else []
    assert m.names[2:] == legs + arm
    assert [m.frames[f].name for f in r.foot_frames] == ["FR_foot", "FL_foot", "RR_foot", "RL_foot"]


def test_product_loader_matches_oracle_model(robots):
    prod, ora = robots
    for name in prod:
        t, m = prod[name].tables(), ora[name].model
        ine = np.array([[m.mass[i], *m.com[i], m.Ic[i][0, 0], m.Ic[i][0, 1], m.Ic[i][0, 2], m.Ic[i][1, 1], m.Ic[i][1, 2],
                         m.Ic[i][2, 2]] for i in range(1, m.njoints)])
        assert np.abs(ine - t["inertia"]).max() < 1e-12
        assert np.array_equal(t["parent"][1:], np.array(m.parents[2:]) - 1)
        assert np.abs(prod[name].q0 - ora[name].q0).max() == 0


def _random_state(r, rng):
    m = r.model
    q = r.q0.copy()
    q[7:] += rng.normal(0, 0.3, m.nq - 7)
    quat = rng.normal(size=4)
    q[3:7] = quat / np.linalg.norm(quat)
    q[:3] = rng.normal(size=3)
    return q, rng.normal(size=m.nv), rng.normal(size=m.nv), rng.normal(size=r.nf) * 20


@pytest.mark.parametrize("name", ["go2", "b2g"])
def test_rigid_body_identities(name):
    """EOM identity of run_ocp.py:106-161, ABA o RNEA = id, centroidal identities, integrate/difference."""
    r = OracleRobot(name)
    m = r.model
    rng = np.random.default_rng(3)
    q, v, a, f = _random_state(r, rng)
    dyn = DynamicsWholeBodyTorque(m, r.mass, r.foot_frames)
    ee = dyn._ee(r.ext_force_frame)
    kin = rbd.Kin(m, q)
    tau = dyn.rnea_dynamics(r.ext_force_frame)(q, v, a, f)
    fext = rbd.local_ext_forces(m, kin, ee, f)
    assert np.abs(rbd.aba(m, kin, v, tau, fext) - a).max() < 1e-10
    z = np.zeros(m.nv)
    nle = rbd.rnea(m, kin, v, z, {})
    M = np.stack([rbd.rnea(m, kin, z, e, {}) - rbd.rnea(m, kin, z, z, {}) for e in np.eye(m.nv)], -1)
    tau_ext = sum(np.stack([dyn.get_frame_velocity(fid)(q, e)[:3] for e in np.eye(m.nv)], -1).T @ f[3 * i:3 * i + 3]
                  for i, fid in enumerate(ee))
    assert np.abs(M @ a + nle - tau_ext - tau).max() < 1e-9
    assert np.abs(M - M.T).max() < 1e-10
    A = rbd.centroidal_map(m, kin)
    assert np.abs(A[:3, :3] / r.mass - kin.oR[1]).max() < 1e-12
    assert np.abs(A @ v - rbd.centroidal_momentum(m, kin, v)).max() < 1e-10
    gaps = DynamicsCentroidalAcc(m, r.mass, r.foot_frames).dynamics_gaps(r.ext_force_frame)(q, v, a, f)
    com = rbd.center_of_mass(m, kin)
    fw = sp.act_force(kin.oR[1], kin.op[1], tau[:6])
    assert np.abs(gaps - np.concatenate([fw[:3], fw[3:] - np.cross(com, fw[:3])])).max() < 1e-9
    x0 = np.concatenate([q, v])
    dx = rng.normal(size=2 * m.nv) * 0.3
    x1 = dyn.state_integrate()(x0, dx)
    assert np.abs(dyn.state_difference()(x0, x1) - dx).max() < 1e-12
    # d/dt (A v) along the motion
    eps = 1e-6

    def h(t):
        qq = rbd.integrate(m, q, v * t + 0.5 * a * t * t)
        return rbd.centroidal_momentum(m, rbd.Kin(m, qq), v + a * t)
    assert np.abs((h(eps) - h(-eps)) / (2 * eps) - rbd.centroidal_momentum_rate(m, kin, v, a)).max() < 1e-5


@pytest.mark.parametrize("rn,kind,n,m,np_", [("go2", "centroidal_vel", 1104, 1744, 258), ("b2", "whole_body_rnea", 1392, 2032, 318),
                                              ("b2g", "whole_body_aba", 1668, 2437, 309), ("b2", "centroidal_acc", 1356, 1960, 282),
                                              ("b2g", "whole_body_rnea", 1842, 2665, 369)])
def test_problem_sizes(robots, rn, kind, n, m, np_):
    """(n, m, np) of SURVEY.md 8(a) row 13 / row 15 at N=20."""
    o = OracleOCP(robots[1][rn], kind, 20)
    assert (o.n, o.m, o.np_) == (n, m, np_)


def test_unknown_dynamics_raises(robots):
    with pytest.raises(ValueError):
        OracleOCP(robots[1]["b2"], "whole_body_foo", 5)


def test_oracle_jacobian_vs_finite_differences(robots):
    from emu_util import random_problem
    rng = np.random.default_rng(7)
    o = OracleOCP(robots[1]["b2g"], "whole_body_rnea", 4)
    x, p = random_problem(o, rng)
    J = o.jac_g(x, p)
    d = rng.normal(size=x.size)
    eps = 1e-6
    gp, _, _ = o.g_data(x + eps * d, p)
    gm, _, _ = o.g_data(x - eps * d, p)
    assert np.abs((gp - gm) / (2 * eps) - J @ d).max() < 1e-6 * np.abs(J @ d).max()
    f0, grad = o.f_data(x, p)
    fp, _ = o.f_data(x + eps * d, p)
    fm, _ = o.f_data(x - eps * d, p)
    assert abs((fp - fm) / (2 * eps) - grad @ d) < 1e-6 * abs(grad @ d)
    assert np.allclose(np.diag(np.diag(o.hess_diag(p))) @ d * 0 + o.hess_diag(p) * d, (o.f_data(x + d, p)[1] - grad), rtol=1e-9, atol=1e-6)


@pytest.mark.parametrize("name", ["go2", "b2g"])
def test_rnea_against_lagrangian_mechanics(name):
    """Independent derivation: the oracle's RNEA against Lagrange's equations built from forward kinematics and the
    link inertial parameters only (no spatial-algebra recursion): kinetic energy T = 1/2 sum m |v_c|^2 + w^T I w from
    finite-differenced link poses, potential U = sum m g z_c.
      * v^T M(q) v = 2 T(q, v) with M taken from RNEA columns,
      * gravity torques rnea(q, 0, 0) = dU/dq along every tangent direction,
      * power balance: v^T (M a + nle) = d/dt (T + U) along the motion."""
    r = OracleRobot(name)
    m = r.model
    rng = np.random.default_rng(21)
    q, v, a, _ = _random_state(r, rng)
    z = np.zeros(m.nv)
    kin = rbd.Kin(m, q)

    def link_poses(qq):
        k = rbd.Kin(m, qq)
        return [(k.oR[i], k.op[i] + k.oR[i] @ m.com[i]) for i in range(1, m.njoints)]

    def inertia_world(i, R):
        return R @ np.asarray(m.Ic[i]) @ R.T

    def energies(qq, vv):
        """T from central differences of the link poses along vv; U from the CoM heights."""
        eps = 1e-6
        pp, pm, p0 = link_poses(rbd.integrate(m, qq, eps * vv)), link_poses(rbd.integrate(m, qq, -eps * vv)), link_poses(qq)
        T = U = 0.0
        for i in range(1, m.njoints):
            (Rp, cp), (Rm, cm), (R0, c0) = pp[i - 1], pm[i - 1], p0[i - 1]
            vc = (cp - cm) / (2 * eps)
            W = (Rp - Rm) / (2 * eps) @ R0.T                 # [w]x in the world frame
            w = np.array([W[2, 1], W[0, 2], W[1, 0]])
            T += 0.5 * m.mass[i] * vc @ vc + 0.5 * w @ inertia_world(i, R0) @ w
            U += m.mass[i] * 9.81 * c0[2]
        return T, U

    M = np.stack([rbd.rnea(m, kin, z, e, {}) - rbd.rnea(m, kin, z, z, {}) for e in np.eye(m.nv)], -1)
    T, _ = energies(q, v)
    assert abs(v @ M @ v - 2 * T) <= 1e-6 * abs(2 * T)
    grav = rbd.rnea(m, kin, z, z, {})
    eps = 1e-6
    for d in range(m.nv):
        e = np.zeros(m.nv)
        e[d] = 1.0
        dU = (energies(rbd.integrate(m, q, eps * e), z)[1] - energies(rbd.integrate(m, q, -eps * e), z)[1]) / (2 * eps)
        assert abs(grav[d] - dU) <= 1e-5 * max(1.0, np.abs(grav).max()), d
    # power balance along q(t) = q (+) (v t + a t^2 / 2), v(t) = v + a t
    tau = rbd.rnea(m, kin, v, a, {})
    h = 1e-4

    def total(t):
        Tt, Ut = energies(rbd.integrate(m, q, v * t + 0.5 * a * t * t), v + a * t)
        return Tt + Ut
    power = (total(h) - total(-h)) / (2 * h)
    assert abs(v @ tau - power) <= 1e-4 * max(1.0, abs(power))


def test_osqp_restatement_against_independent_solver():
    """The oracle's ADMM (Ruiz scaling, rho_vec, relaxation) against an independent method on random convex QPs
    min 1/2 x^T P x + q^T x, l <= A x <= u with equality, inequality and free rows: run to tight tolerances it must
    reach the optimum that scipy's trust-constr interior-point method finds, and the returned duals must satisfy the
    KKT conditions; an infeasible problem must be reported as such."""
    from scipy import sparse
    from scipy.optimize import Bounds, LinearConstraint, minimize
    from oracle.osqp_admm import OSQP
    rng = np.random.default_rng(17)
    for trial in range(3):
        n, m = 12, 18
        P = rng.uniform(0.5, 3.0, n)
        q = rng.normal(size=n)
        A = sparse.random(m, n, density=0.35, random_state=int(rng.integers(1 << 30)), data_rvs=rng.standard_normal).tocsc()
        A = (A + sparse.csc_matrix((np.full(n, 1e-3), (np.arange(n), np.arange(n))), shape=(m, n))).tocsc()
        x_feas = rng.normal(size=n)
        Ax = A @ x_feas
        l, u = Ax - rng.uniform(0.1, 1.0, m), Ax + rng.uniform(0.1, 1.0, m)
        l[:4] = u[:4] = Ax[:4]                      # equality rows
        l[4:7], u[4:7] = -np.inf, np.inf            # free rows
        l[7:9] = -np.inf                            # one-sided rows
        s = OSQP()
        s.setup(P, np.ones(n), sparse.csc_matrix((np.ones_like(A.data), A.indices, A.indptr), shape=A.shape), -np.ones(m), np.ones(m),
                max_iter=20000, alpha=1.4, rho=2e-2, eps_abs=1e-9, eps_rel=1e-9)
        s.update(q=q, Ax=A.data, l=l, u=u)
        xs = s.solve()
        assert s.status == "solved"
        ref = minimize(lambda x: 0.5 * x @ (P * x) + q @ x, x_feas, jac=lambda x: P * x + q, hess=lambda x: np.diag(P),
                       method="trust-constr", constraints=[LinearConstraint(A.toarray(), l, u)], bounds=Bounds(-np.inf, np.inf),
                       options={"gtol": 1e-10, "xtol": 1e-12, "maxiter": 3000})
        assert np.abs(xs - ref.x).max() < 1e-5 * max(1.0, np.abs(ref.x).max())
        # KKT with the ADMM duals: stationarity, primal feasibility, complementarity signs
        y = s.E * s.y / s.c                         # unscaled duals
        assert np.abs(P * xs + q + A.T @ y).max() < 1e-6
        Axs = A @ xs
        assert (Axs >= l - 1e-6).all() and (Axs <= u + 1e-6).all()
        assert (y[(Axs > l + 1e-5) & (Axs < u - 1e-5)].__abs__() < 1e-6).all()
    # primal infeasible: x0 >= 1 and x0 <= 0
    s = OSQP()
    Ai = sparse.csc_matrix(np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]]))
    s.setup(np.ones(2), np.zeros(2), Ai, np.array([1.0, -np.inf, -1.0]), np.array([np.inf, 0.0, 1.0]), max_iter=4000, alpha=1.4, rho=2e-2)
    out = s.solve()
    assert s.status == "primal infeasible" and np.isnan(out).all()


@pytest.mark.parametrize("name", ["go2", "b2g"])
def test_frame_velocity_against_finite_differences(name):
    """getFrameVelocity(LOCAL_WORLD_ALIGNED) of the oracle against the time derivative of the frame placement obtained
    by forward kinematics alone; base-relative variant (dynamics/dynamics.py:86-113) against its definition."""
    r = OracleRobot(name)
    m = r.model
    rng = np.random.default_rng(5)
    q, v, _, _ = _random_state(r, rng)
    dyn = DynamicsWholeBodyTorque(m, r.mass, r.foot_frames)
    eps = 1e-6
    frames = list(r.foot_frames) + ([r.arm_ee_frame] if r.arm_ee_frame else [])
    for fid in frames:
        Rp, pp = rbd.Kin(m, rbd.integrate(m, q, eps * v)).frame_placement(fid)
        Rm, pm = rbd.Kin(m, rbd.integrate(m, q, -eps * v)).frame_placement(fid)
        R0, _ = rbd.Kin(m, q).frame_placement(fid)
        W = (Rp - Rm) / (2 * eps) @ R0.T
        ref = np.concatenate([(pp - pm) / (2 * eps), [W[2, 1], W[0, 2], W[1, 0]]])
        got = dyn.get_frame_velocity(fid)(q, v)
        assert np.abs(got - ref).max() < 1e-6 * max(1.0, np.abs(ref).max())
    if r.arm_ee_frame:
        kin = rbd.Kin(m, q)
        vf = dyn.get_frame_velocity(r.arm_ee_frame)(q, v)
        vb = dyn.get_frame_velocity(dyn.base_frame)(q, v)
        Rb, pb = kin.frame_placement(dyn.base_frame)
        _, pf = kin.frame_placement(r.arm_ee_frame)
        lin = Rb.T @ (vf[:3] - vb[:3] - np.cross(vb[3:], pf - pb))
        ang = Rb.T @ (vf[3:] - vb[3:])
        ref = np.array([lin[0], lin[1], vf[2], ang[0], ang[1], vf[5]])
        got = dyn.get_frame_velocity(r.arm_ee_frame, relative_to_base=True)(q, v)
        assert np.abs(got - ref).max() < 1e-10
