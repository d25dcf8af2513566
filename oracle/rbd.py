"""Rigid-body algorithms in pinocchio's local-frame formulation (oracle; test infrastructure only).

Restates the pinocchio calls at the reference call sites listed in SURVEY.md section 2.2
(dynamics/dynamics.py:48-63,71-72,83-90; dynamics_whole_body_torque.py:54-69,86-101;
dynamics_centroidal_acc.py:97-117; dynamics_centroidal_vel.py:54-69,144-146) following the
published algorithms [3P] summarised in SURVEY.md appendix A.3-A.6.  Everything broadcasts over
leading batch dimensions and accepts complex inputs (complex-step derivatives).
"""
import numpy as np

from . import spatial as sp


def _axis_rot(axis, ang):
    """Rodrigues rotation about a constant unit axis."""
    K = sp.skew(np.asarray(axis, dtype=float))
    s, c = np.sin(ang)[..., None, None], np.cos(ang)[..., None, None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def integrate(model, q, dq):
    """pin.integrate: base M * exp6(dq[:6]), joints additive (appendix A.2)."""
    R0 = sp.quat_to_R(q[..., 3:7])
    Re, pe = sp.exp6(dq[..., :6])
    p = q[..., :3] + sp.mv(R0, pe)
    quat = sp.quat_mul(q[..., 3:7], sp.exp3_quat(dq[..., 3:6]))
    return np.concatenate([p, quat, q[..., 7:] + dq[..., 6:]], -1)


def difference(model, q0, q1):
    """pin.difference for real, non-batched q (only ever applied to parameters)."""
    R0, R1 = sp.quat_to_R(q0[3:7]), sp.quat_to_R(q1[3:7])
    nu = sp.log6(R0.T @ R1, R0.T @ (q1[:3] - q0[:3]))
    return np.concatenate([nu, q1[7:] - q0[7:]])


class Kin:
    """Forward kinematics: liMi and oMi for every joint."""

    def __init__(self, model, q):
        n = model.njoints
        self.model = model
        self.lR, self.lp, self.oR, self.op = [None] * n, [None] * n, [None] * n, [None] * n
        Rb = sp.quat_to_R(q[..., 3:7])
        self.lR[1], self.lp[1] = Rb, q[..., :3]
        self.oR[1], self.op[1] = Rb, q[..., :3]
        for i in range(2, n):
            Rj = _axis_rot(model.axis[i], q[..., model.idx_q[i]])
            self.lR[i] = model.placement_R[i] @ Rj
            self.lp[i] = np.broadcast_to(model.placement_p[i], Rj.shape[:-2] + (3,))
            par = model.parents[i]
            self.oR[i] = sp.mm(self.oR[par], self.lR[i])
            self.op[i] = self.op[par] + sp.mv(self.oR[par], self.lp[i])

    def frame_placement(self, frame_id):
        f = self.model.frames[frame_id]
        R = sp.mm(self.oR[f.parent], f.R)
        p = self.op[f.parent] + sp.mv(self.oR[f.parent], f.p)
        return R, p


def _S(model, i):
    """Joint motion subspace columns (list of 6-vectors) in the joint frame."""
    if i == 1:
        return [np.eye(6)[k] for k in range(6)]
    return [np.concatenate([np.zeros(3), model.axis[i]])]


def _vj(model, i, v):
    """S_i * qdot_i."""
    if i == 1:
        return v[..., :6]
    a = np.concatenate([np.zeros(3), model.axis[i]])
    return v[..., model.idx_v[i], None] * a


def body_velocities(model, kin, v):
    vel = [None] * model.njoints
    vel[0] = np.zeros(v.shape[:-1] + (6,), dtype=v.dtype)
    for i in range(1, model.njoints):
        vel[i] = sp.actinv_motion(kin.lR[i], kin.lp[i], vel[model.parents[i]]) + _vj(model, i, v)
    return vel


def local_ext_forces(model, kin, ee_frames, forces):
    """f_ext[parentJoint] = [R_wj^T f; r x (R_wj^T f)] (dynamics_whole_body_torque.py:56-66), assignment."""
    fext = {}
    for idx, fid in enumerate(ee_frames):
        fr = model.frames[fid]
        f_lin = sp.mtv(kin.oR[fr.parent], forces[..., 3 * idx:3 * idx + 3])
        fext[fr.parent] = np.concatenate([f_lin, sp.cross(fr.p, f_lin)], -1)
    return fext


def rnea(model, kin, v, a, fext):
    """pin.rnea(model, data, q, v, a, fext) -> tau (appendix A.3)."""
    n = model.njoints
    batch = np.broadcast(v[..., 0], a[..., 0], kin.oR[1][..., 0, 0], *[f_[..., 0] for f_ in fext.values()]).shape
    dt = np.result_type(v.dtype, a.dtype, kin.oR[1].dtype, *[f_.dtype for f_ in fext.values()])
    vel, acc, f = [None] * n, [None] * n, [None] * n
    vel[0] = np.zeros(batch + (6,), dtype=dt)
    acc[0] = np.zeros(batch + (6,), dtype=dt) + np.concatenate([-model.gravity, np.zeros(3)])
    for i in range(1, n):
        par = model.parents[i]
        vj = _vj(model, i, v)
        vel[i] = sp.actinv_motion(kin.lR[i], kin.lp[i], vel[par]) + vj
        acc[i] = sp.actinv_motion(kin.lR[i], kin.lp[i], acc[par]) + _vj(model, i, a) + sp.cross_mm(vel[i], vj)
        h = sp.inertia_mul(model.mass[i], model.com[i], model.Ic[i], vel[i])
        f[i] = sp.inertia_mul(model.mass[i], model.com[i], model.Ic[i], acc[i]) + sp.cross_mf(vel[i], h)
        if i in fext:
            f[i] = f[i] - fext[i]
    tau = np.zeros(batch + (model.nv,), dtype=dt)
    for i in range(n - 1, 0, -1):
        if i == 1:
            tau[..., :6] = f[i]
        else:
            tau[..., model.idx_v[i]] = np.sum(f[i][..., 3:] * model.axis[i], -1)
            par = model.parents[i]
            f[par] = f[par] + sp.act_force(kin.lR[i], kin.lp[i], f[i])
    return tau


def aba(model, kin, v, tau, fext):
    """pin.aba(model, data, q, v, tau, fext) -> a (Featherstone ABA, appendix A.4)."""
    n = model.njoints
    batch = np.broadcast(v[..., 0], tau[..., 0], kin.oR[1][..., 0, 0], *[f_[..., 0] for f_ in fext.values()]).shape
    dt = np.result_type(v.dtype, tau.dtype, kin.oR[1].dtype, *[f_.dtype for f_ in fext.values()])
    vel, c, IA, pA = [None] * n, [None] * n, [None] * n, [None] * n
    vel[0] = np.zeros(batch + (6,), dtype=dt)
    for i in range(1, n):
        par = model.parents[i]
        vj = _vj(model, i, v)
        vel[i] = sp.actinv_motion(kin.lR[i], kin.lp[i], vel[par]) + vj
        c[i] = sp.cross_mm(vel[i], vj)
        Y = sp.inertia_matrix(model.mass[i], model.com[i], model.Ic[i])
        IA[i] = np.zeros(batch + (6, 6), dtype=dt) + Y
        pA[i] = sp.cross_mf(vel[i], sp.mv(Y, vel[i]))
        if i in fext:
            pA[i] = pA[i] - fext[i]
    U, Dinv, u = [None] * n, [None] * n, [None] * n
    for i in range(n - 1, 0, -1):
        if i == 1:
            U[i] = IA[i]
            Dinv[i] = np.linalg.inv(IA[i])
            u[i] = tau[..., :6] - pA[i]
        else:
            s = np.concatenate([np.zeros(3), model.axis[i]])
            U[i] = sp.mv(IA[i], s)[..., None]                                # [...,6,1]
            Dinv[i] = 1.0 / np.sum(U[i][..., 0] * s, -1)[..., None, None]    # [...,1,1]
            u[i] = (tau[..., model.idx_v[i]] - np.sum(pA[i] * s, -1))[..., None]
            Ia = IA[i] - sp.mm(sp.mm(U[i], Dinv[i]), np.swapaxes(U[i], -1, -2))
            pa = pA[i] + sp.mv(Ia, c[i]) + sp.mv(sp.mm(U[i], Dinv[i]), u[i])
            par = model.parents[i]
            Xf = sp.force_xform(kin.lR[i], kin.lp[i])            # child force -> parent
            IA[par] = IA[par] + sp.mm(sp.mm(Xf, Ia), np.swapaxes(Xf, -1, -2))
            pA[par] = pA[par] + sp.mv(Xf, pa)
    acc = [None] * n
    acc[0] = np.zeros(batch + (6,), dtype=dt) + np.concatenate([-model.gravity, np.zeros(3)])
    out = np.zeros(batch + (model.nv,), dtype=dt)
    for i in range(1, n):
        par = model.parents[i]
        ai = sp.actinv_motion(kin.lR[i], kin.lp[i], acc[par]) + c[i]
        if i == 1:
            qdd = sp.mv(Dinv[i], u[i] - sp.mtv(U[i], ai))
            out[..., :6] = qdd
            acc[i] = ai + qdd
        else:
            s = np.concatenate([np.zeros(3), model.axis[i]])
            qdd = sp.mv(Dinv[i], u[i] - sp.mtv(U[i], ai))
            out[..., model.idx_v[i]] = qdd[..., 0]
            acc[i] = ai + qdd * s
    return out


def center_of_mass(model, kin):
    c = 0
    for i in range(1, model.njoints):
        c = c + model.mass[i] * (kin.op[i] + sp.mv(kin.oR[i], model.com[i]))
    return c / model.total_mass


def _to_centroidal(kin, i, f_local, com):
    """Local force of joint i -> world-aligned axes about the CoM."""
    fw = sp.act_force(kin.oR[i], kin.op[i], f_local)
    return np.concatenate([fw[..., :3], fw[..., 3:] - sp.cross(com, fw[..., :3])], -1)


def centroidal_momentum(model, kin, v, com=None):
    """A_g(q) v = [m cdot; L_c] (computeCentroidalMap applied to v, appendix A.5)."""
    if com is None:
        com = center_of_mass(model, kin)
    vel = body_velocities(model, kin, v)
    h = 0
    for i in range(1, model.njoints):
        hi = sp.inertia_mul(model.mass[i], model.com[i], model.Ic[i], vel[i])
        h = h + _to_centroidal(kin, i, hi, com)
    return h


def centroidal_map(model, kin):
    """A_g(q) (6 x nv), column c = momentum for v = e_c."""
    com = center_of_mass(model, kin)
    batch = kin.oR[1].shape[:-2]
    cols = []
    for c in range(model.nv):
        e = np.zeros(batch + (model.nv,))
        e[..., c] = 1.0
        cols.append(centroidal_momentum(model, kin, e, com))
    return np.stack(cols, -1)


def centroidal_momentum_rate(model, kin, v, a):
    """d/dt (A_g v) = A_g a + Adot_g v: momentum rate without gravity, about the CoM."""
    n = model.njoints
    com = center_of_mass(model, kin)
    batch = np.broadcast(v[..., 0], a[..., 0]).shape
    dt = np.result_type(v.dtype, a.dtype, kin.oR[1].dtype)
    vel, acc = [None] * n, [None] * n
    vel[0] = np.zeros(batch + (6,), dtype=dt)
    acc[0] = np.zeros(batch + (6,), dtype=dt)
    out = 0
    for i in range(1, n):
        par = model.parents[i]
        vj = _vj(model, i, v)
        vel[i] = sp.actinv_motion(kin.lR[i], kin.lp[i], vel[par]) + vj
        acc[i] = sp.actinv_motion(kin.lR[i], kin.lp[i], acc[par]) + _vj(model, i, a) + sp.cross_mm(vel[i], vj)
        h = sp.inertia_mul(model.mass[i], model.com[i], model.Ic[i], vel[i])
        f = sp.inertia_mul(model.mass[i], model.com[i], model.Ic[i], acc[i]) + sp.cross_mf(vel[i], h)
        out = out + _to_centroidal(kin, i, f, com)
    return out


def dccrba_times_v(model, kin, v):
    """Adot_g(q, v) v (the only use of cpin.dccrba in the reference)."""
    return centroidal_momentum_rate(model, kin, v, np.zeros_like(v))


def frame_velocity_lwa(model, kin, vel, frame_id):
    """getFrameVelocity(..., LOCAL_WORLD_ALIGNED).vector (appendix A.6)."""
    f = model.frames[frame_id]
    vj = vel[f.parent]
    lin = vj[..., :3] + sp.cross(vj[..., 3:], f.p)
    R = kin.oR[f.parent]
    return np.concatenate([sp.mv(R, lin), sp.mv(R, vj[..., 3:])], -1)
