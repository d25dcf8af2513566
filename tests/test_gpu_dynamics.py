"""GPU parity of the Dynamics* plugin Functions (values and analytic Jacobians) against the oracle's classes."""
import numpy as np
import pytest

from oracle import dynamics as odyn
from oracle import rbd

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL = 1e-9


def _states(r, rng, B):
    q = np.tile(r.q0, (B, 1))
    q[:, :3] = rng.normal(size=(B, 3))
    quat = rng.normal(size=(B, 4))
    q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    q[:, 7:] += rng.normal(0, 0.3, (B, r.nq - 7))
    return q, rng.normal(size=(B, r.nv)), rng.normal(size=(B, r.nv)), rng.normal(size=(B, r.nf)) * 20


def _cstep(fun, args, idx, model=None):
    h, x = 1e-30, args[idx]
    n = model.nv if model is not None else x.size
    cols = []
    for d in range(n):
        e = np.zeros(n, complex)
        e[d] = 1j * h
        a2 = list(args)
        a2[idx] = rbd.integrate(model, x.astype(complex), e) if model is not None else x + e
        cols.append(np.imag(fun(*a2)) / h)
    return np.stack(cols, -1)


def _t(a):
    return torch.tensor(np.ascontiguousarray(a), device="cuda")


@pytest.mark.parametrize("rn", ["b2g", "go2"])
def test_rnea_aba_gaps_match_oracle(robots, rn):
    from pino_locoman_b200 import dynamics as pdyn
    prod, ora = robots
    r, o = prod[rn], ora[rn]
    rng = np.random.default_rng(4)
    B = 3
    q, v, a, f = _states(o, rng, B)
    ext = r.ext_force_frame
    d = pdyn.DynamicsWholeBodyTorque(r.model, r.mass, r.foot_frames, max_batch=B)
    od = odyn.DynamicsWholeBodyTorque(o.model, o.mass, o.foot_frames, o.base_frame)
    tau, jac = d.rnea_dynamics(ext).jacobian(_t(q), _t(v), _t(a), _t(f))
    tau_j = rng.normal(size=(B, r.nj)) * 10
    acc, ajac = d.aba_dynamics(ext).jacobian(_t(q), _t(v), _t(tau_j), _t(f))
    fr = od.rnea_dynamics(o.ext_force_frame)
    fa = od.aba_dynamics(o.ext_force_frame)
    for b in range(B):
        ref = fr(q[b], v[b], a[b], f[b])
        assert np.abs(tau[b].cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
        Jr = np.concatenate([_cstep(fr, [q[b], v[b], a[b], f[b]], 0, o.model)] + [_cstep(fr, [q[b], v[b], a[b], f[b]], k) for k in (1, 2, 3)], 1)
        assert np.abs(jac[b].cpu().numpy() - Jr).max() <= TOL * np.abs(Jr).max()
        ref = fa(q[b], v[b], tau_j[b], f[b])
        assert np.abs(acc[b].cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
        Ja = np.concatenate([_cstep(fa, [q[b], v[b], tau_j[b], f[b]], 0, o.model)] + [_cstep(fa, [q[b], v[b], tau_j[b], f[b]], k) for k in (1, 2, 3)], 1)
        assert np.abs(ajac[b].cpu().numpy() - Ja).max() <= TOL * np.abs(Ja).max()
    # integrate / difference round trip and against the oracle
    x0 = np.concatenate([q, v], 1)
    dx = rng.normal(size=(B, 2 * r.nv)) * 0.3
    x1 = d.state_integrate()(_t(x0), _t(dx))
    back = d.state_difference()(_t(x0), x1)
    assert np.abs(back.cpu().numpy() - dx).max() < 1e-12
    for b in range(B):
        assert np.abs(x1[b].cpu().numpy() - od.state_integrate()(x0[b], dx[b])).max() < 1e-12
    # frame velocities: all six components (LOCAL_WORLD_ALIGNED / the base-relative variant) and positions, any frame;
    # the three components the OCP rows use also through the node kernel's own rows
    has_base = "base_link" in r.model.frames           # (Go2's root link is called "base": the reference's base_frame id is invalid there)
    frames = list(zip(r.foot_frames, o.foot_frames)) + [("FR_thigh", o.model.getFrameId("FR_thigh"))]
    if has_base:
        frames.append(("base_link", o.model.getFrameId("base_link")))
    if r.arm_ee_frame:
        frames.append((r.arm_ee_frame, o.arm_ee_frame))
    for fname, fid in frames:
        for rel in ((False, True) if has_base else (False,)):
            vel = d.get_frame_velocity(fname, relative_to_base=rel)(_t(q), _t(v)).cpu().numpy()
            assert vel.shape == (B, 6)
            for b in range(B):
                ref = od.get_frame_velocity(fid, relative_to_base=rel)(q[b], v[b])
                assert np.abs(vel[b] - ref).max() <= TOL * max(1.0, np.abs(ref).max()), (fname, rel)
        pos = d.get_frame_position(fname)(_t(q)).cpu().numpy()
        for b in range(B):
            ref = od.get_frame_position(fid)(q[b])
            assert np.abs(pos[b] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), fname
    for k, fname in enumerate(r.foot_frames):
        vel3 = d.frame_velocity_rows(fname)(_t(q), _t(v)).cpu().numpy()
        vel6 = d.get_frame_velocity(fname)(_t(q), _t(v)).cpu().numpy()
        assert np.abs(vel3 - vel6[:, :3]).max() <= 1e-12 * max(1.0, np.abs(vel6).max())
    if r.arm_ee_frame:
        vel3 = d.frame_velocity_rows(r.arm_ee_frame, relative_to_base=True)(_t(q), _t(v)).cpu().numpy()
        vel6 = d.get_frame_velocity(r.arm_ee_frame, relative_to_base=True)(_t(q), _t(v)).cpu().numpy()
        assert np.abs(vel3 - vel6[:, :3]).max() <= 1e-12 * max(1.0, np.abs(vel6).max())
    with pytest.raises(KeyError):
        d.get_frame_position("no_such_frame")


def test_centroidal_functions_match_oracle(robots):
    from pino_locoman_b200 import dynamics as pdyn
    prod, ora = robots
    r, o = prod["b2g"], ora["b2g"]
    rng = np.random.default_rng(8)
    B = 2
    q, v, a, f = _states(o, rng, B)
    ext = r.ext_force_frame
    dca = pdyn.DynamicsCentroidalAcc(r.model, r.mass, r.foot_frames, max_batch=B)
    dwa = pdyn.DynamicsWholeBodyAcc(r.model, r.mass, r.foot_frames, max_batch=B)
    dcv = pdyn.DynamicsCentroidalVel(r.model, r.mass, r.foot_frames, max_batch=B)
    oca = odyn.DynamicsCentroidalAcc(o.model, o.mass, o.foot_frames, o.base_frame).dynamics_gaps(o.ext_force_frame)
    owa = odyn.DynamicsWholeBodyAcc(o.model, o.mass, o.foot_frames, o.base_frame).dynamics_gaps(o.ext_force_frame)
    ocv = odyn.DynamicsCentroidalVel(o.model, o.mass, o.foot_frames, o.base_frame)
    g1, j1 = dca.dynamics_gaps(ext).jacobian(_t(q), _t(v), _t(a), _t(f))
    g2, j2 = dwa.dynamics_gaps(ext).jacobian(_t(q), _t(v), _t(a), _t(f))
    hst = rng.normal(size=(B, 6))
    g3 = dcv.dynamics_gaps()(_t(hst), _t(q), _t(v))
    g4 = dcv.com_dynamics(ext)(_t(q), _t(f))
    for b in range(B):
        for got, jac, fn in ((g1, j1, oca), (g2, j2, owa)):
            ref = fn(q[b], v[b], a[b], f[b])
            assert np.abs(got[b].cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
            Jr = np.concatenate([_cstep(fn, [q[b], v[b], a[b], f[b]], 0, o.model)] + [_cstep(fn, [q[b], v[b], a[b], f[b]], k) for k in (1, 2, 3)], 1)
            assert np.abs(jac[b].cpu().numpy() - Jr).max() <= TOL * np.abs(Jr).max()
        ref = ocv.dynamics_gaps()(hst[b], q[b], v[b])
        assert np.abs(g3[b].cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
        ref = ocv.com_dynamics(o.ext_force_frame)(q[b], f[b])
        assert np.abs(g4[b].cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
    # base_vel / base_acc (the 6x6 solves of the base rows), interleaved with the gap Jacobians above on the same probes
    oCA = odyn.DynamicsCentroidalAcc(o.model, o.mass, o.foot_frames, o.base_frame)
    oWA = odyn.DynamicsWholeBodyAcc(o.model, o.mass, o.foot_frames, o.base_frame)
    a_j, v_j = a[:, 6:], v[:, 6:]
    ab1 = dca.base_acc_dynamics(ext)(_t(q), _t(v), _t(a_j), _t(f)).cpu().numpy()
    ab2 = dwa.base_acc_dynamics(ext)(_t(q), _t(v), _t(a_j), _t(f)).cpu().numpy()
    ab3 = dcv.base_acc_dynamics(ext)(_t(q), _t(v), _t(a_j), _t(f)).cpu().numpy()
    vb = dcv.base_vel_dynamics()(_t(hst), _t(q), _t(v_j)).cpu().numpy()
    g1b, j1b = dca.dynamics_gaps(ext).jacobian(_t(q), _t(v), _t(a), _t(f))       # map rebuilt after the solve
    assert torch.equal(g1, g1b) and torch.equal(j1, j1b)
    for b in range(B):
        ref = oCA.base_acc_dynamics(o.ext_force_frame)(q[b], v[b], a_j[b], f[b])
        assert np.abs(ab1[b] - ref).max() <= TOL * np.abs(ref).max()
        assert np.abs(ab3[b] - ref).max() <= TOL * np.abs(ref).max()
        ref = oWA.base_acc_dynamics(o.ext_force_frame)(q[b], v[b], a_j[b], f[b])
        assert np.abs(ab2[b] - ref).max() <= TOL * np.abs(ref).max()
        ref = ocv.base_vel_dynamics()(hst[b], q[b], v_j[b])
        assert np.abs(vb[b] - ref).max() <= TOL * np.abs(ref).max()
        # consistency: the solved base part zeroes the gaps
        full_a = np.concatenate([ab1[b], a_j[b]])
        assert np.abs(oca(q[b], v[b], full_a, f[b])).max() <= 1e-8 * max(1.0, np.abs(f[b]).max())
