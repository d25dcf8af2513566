"""Dynamics* function factories with the reference's signatures (oracle; test infrastructure only).

Mirrors dynamics/dynamics.py:6-118, dynamics_whole_body_torque.py:8-103,
dynamics_whole_body_acc.py:8-126, dynamics_centroidal_acc.py:8-119 and
dynamics_centroidal_vel.py:8-148: each factory returns a plain callable with the
casadi.Function argument order of the reference.  Inputs may carry leading batch
dimensions and may be complex (complex-step differentiation).
"""
import numpy as np

from . import rbd, spatial as sp


class Dynamics:
    def __init__(self, model, mass, foot_frames, base_frame=None):
        self.model = model
        self.mass = mass
        self.foot_frames = list(foot_frames)
        self.base_frame = model.getFrameId("base_link") if base_frame is None else base_frame
        self.nq, self.nv, self.nj = model.nq, model.nv, model.nq - 7

    def _ee(self, ext_force_frame):
        ee = self.foot_frames.copy()
        if ext_force_frame:
            ee.append(ext_force_frame)
        return ee

    # dynamics.py:33-65 (== dynamics_whole_body_torque.py:42-71)
    def rnea_dynamics(self, ext_force_frame=None):
        ee = self._ee(ext_force_frame)

        def rnea_dyn(q, v, a, forces):
            kin = rbd.Kin(self.model, q)
            return rbd.rnea(self.model, kin, v, a, rbd.local_ext_forces(self.model, kin, ee, forces))
        return rnea_dyn

    # dynamics.py:67-75
    def get_frame_position(self, frame_id):
        def frame_pos(q):
            return rbd.Kin(self.model, q).frame_placement(frame_id)[1]
        return frame_pos

    # dynamics.py:77-118
    def get_frame_velocity(self, frame_id, relative_to_base=False):
        def frame_vel(q, v):
            kin = rbd.Kin(self.model, q)
            vel = rbd.body_velocities(self.model, kin, v)
            fv = rbd.frame_velocity_lwa(self.model, kin, vel, frame_id)
            if not relative_to_base:
                return fv
            bv = rbd.frame_velocity_lwa(self.model, kin, vel, self.base_frame)
            Rb, pb = kin.frame_placement(self.base_frame)
            _, pf = kin.frame_placement(frame_id)
            rel = pf - pb
            lin_w = fv[..., :3] - bv[..., :3] - sp.cross(bv[..., 3:], rel)
            ang_w = fv[..., 3:] - bv[..., 3:]
            lin_b, ang_b = sp.mtv(Rb, lin_w), sp.mtv(Rb, ang_w)
            # z components stay in the world frame (dynamics.py:108-113)
            return np.concatenate([lin_b[..., :2], fv[..., 2:3], ang_b[..., :2], fv[..., 5:6]], -1)
        return frame_vel

    # shared by the three (q, v) state classes
    def _integrate_qv(self, x, dx):
        nq, nv = self.nq, self.nv
        q_next = rbd.integrate(self.model, x[..., :nq], dx[..., :nv])
        return np.concatenate([q_next, x[..., nq:] + dx[..., nv:]], -1)

    def _difference_qv(self, x0, x1):
        nq = self.nq
        return np.concatenate([rbd.difference(self.model, x0[:nq], x1[:nq]), x1[nq:] - x0[nq:]])

    def _com_wrench(self, kin, com, ee, forces):
        """[sum f + m g ; sum (p_k - c) x f_k] (dynamics_centroidal_acc.py:103-114)."""
        dp = 0
        dl = 0
        for idx, fid in enumerate(ee):
            f = forces[..., 3 * idx:3 * idx + 3]
            dp = dp + f
            dl = dl + sp.cross(kin.frame_placement(fid)[1] - com, f)
        dp = dp + np.array([0, 0, -9.81 * self.mass])
        return np.concatenate([dp, dl], -1)


class DynamicsWholeBodyTorque(Dynamics):
    def state_integrate(self):
        return self._integrate_qv

    def state_difference(self):
        return self._difference_qv

    # dynamics_whole_body_torque.py:73-103
    def aba_dynamics(self, ext_force_frame=None):
        ee = self._ee(ext_force_frame)

        def aba_dyn(q, v, tau_j, forces):
            kin = rbd.Kin(self.model, q)
            tau = np.concatenate([np.zeros(tau_j.shape[:-1] + (6,), dtype=tau_j.dtype), tau_j], -1)
            return rbd.aba(self.model, kin, v, tau, rbd.local_ext_forces(self.model, kin, ee, forces))
        return aba_dyn


class DynamicsWholeBodyAcc(Dynamics):
    def state_integrate(self):
        return self._integrate_qv

    def state_difference(self):
        return self._difference_qv

    # dynamics_whole_body_acc.py:85-126
    def dynamics_gaps(self, ext_force_frame=None):
        full = self.rnea_dynamics(ext_force_frame)

        def dyn_gaps(q, v, a, forces):
            return full(q, v, a, forces)[..., :6]
        return dyn_gaps

    # dynamics_whole_body_acc.py:43-83: a_b = M_bb^-1 (-nle_b - M_bj a_j + sum J_c^T f); M columns from RNEA differences
    def base_acc_dynamics(self, ext_force_frame=None):
        ee = self._ee(ext_force_frame)

        def base_acc(q, v, a_j, forces):
            kin = rbd.Kin(self.model, q)
            zero_f = rbd.local_ext_forces(self.model, kin, ee, np.zeros_like(forces))
            fext = rbd.local_ext_forces(self.model, kin, ee, forces)
            zv = np.zeros_like(v)
            grav = rbd.rnea(self.model, kin, zv, zv, zero_f)
            M_b = []
            for c in range(self.nv):
                e = np.zeros_like(v)
                e[..., c] = 1.0
                M_b.append((rbd.rnea(self.model, kin, zv, e, zero_f) - grav)[..., :6])
            M_b = np.stack(M_b, -1)                                   # [.., 6, nv] = M[:6, :]
            rhs = -rbd.rnea(self.model, kin, v, zv, fext)[..., :6] - np.einsum("...ij,...j->...i", M_b[..., 6:], a_j)
            return np.linalg.solve(M_b[..., :6], rhs[..., None])[..., 0]
        return base_acc


class DynamicsCentroidalAcc(Dynamics):
    def state_integrate(self):
        return self._integrate_qv

    def state_difference(self):
        return self._difference_qv

    # dynamics_centroidal_acc.py:84-119
    def dynamics_gaps(self, ext_force_frame=None):
        ee = self._ee(ext_force_frame)

        def dyn_gaps(q, v, a, forces):
            kin = rbd.Kin(self.model, q)
            com = rbd.center_of_mass(self.model, kin)
            dh = self._com_wrench(kin, com, ee, forces)
            return rbd.centroidal_momentum_rate(self.model, kin, v, a) - dh
        return dyn_gaps

    # dynamics_centroidal_acc.py:43-82: a_b = A_b^-1 (dh - Adot v - A_j a_j)
    def base_acc_dynamics(self, ext_force_frame=None):
        return _centroidal_base_acc(self, ext_force_frame)


class DynamicsCentroidalVel(Dynamics):
    # dynamics_centroidal_vel.py:12-27
    def state_integrate(self):
        def integrate(x, dx):
            q_next = rbd.integrate(self.model, x[..., 6:], dx[..., 6:])
            return np.concatenate([x[..., :6] + dx[..., :6], q_next], -1)
        return integrate

    # dynamics_centroidal_vel.py:29-41
    def state_difference(self):
        def difference(x0, x1):
            return np.concatenate([x1[:6] - x0[:6], rbd.difference(self.model, x0[6:], x1[6:])])
        return difference

    # dynamics_centroidal_vel.py:43-71
    def com_dynamics(self, ext_force_frame=None):
        ee = self._ee(ext_force_frame)

        def com_dyn(q, forces):
            kin = rbd.Kin(self.model, q)
            com = rbd.center_of_mass(self.model, kin)
            return self._com_wrench(kin, com, ee, forces) / self.mass
        return com_dyn

    # dynamics_centroidal_vel.py:136-148
    def dynamics_gaps(self):
        def dyn_gaps(h, q, v):
            kin = rbd.Kin(self.model, q)
            return rbd.centroidal_momentum(self.model, kin, v) - h * self.mass
        return dyn_gaps

    # dynamics_centroidal_vel.py:73-89: v_b = A_b^-1 (m h - A_j v_j)
    def base_vel_dynamics(self):
        def base_vel(h, q, v_j):
            kin = rbd.Kin(self.model, q)
            A = rbd.centroidal_map(self.model, kin)
            rhs = h * self.mass - np.einsum("...ij,...j->...i", A[..., 6:], v_j)
            return np.linalg.solve(A[..., :6], rhs[..., None])[..., 0]
        return base_vel

    # dynamics_centroidal_vel.py:91-134
    def base_acc_dynamics(self, ext_force_frame=None):
        return _centroidal_base_acc(self, ext_force_frame)


def _centroidal_base_acc(dyn, ext_force_frame):
    ee = dyn._ee(ext_force_frame)

    def base_acc(q, v, a_j, forces):
        kin = rbd.Kin(dyn.model, q)
        com = rbd.center_of_mass(dyn.model, kin)
        dh = dyn._com_wrench(kin, com, ee, forces)
        A = rbd.centroidal_map(dyn.model, kin)
        rhs = dh - rbd.dccrba_times_v(dyn.model, kin, v) - np.einsum("...ij,...j->...i", A[..., 6:], a_j)
        return np.linalg.solve(A[..., :6], rhs[..., None])[..., 0]
    return base_acc
