"""Dynamics* plugin classes: the reference's casadi.Function factories (dynamics/*.py) as batched GPU callables.

``Dynamics*(model, mass, foot_frames)`` keep the reference constructor (``model`` is ``robot.model``); every factory
method returns a callable with the reference Function's argument order that takes ``torch.float64`` CUDA tensors with
a leading batch dimension.  Where the reference differentiates these Functions with casadi AD, the callables offer
``.jacobian(...)`` returning the analytic Jacobian computed by the same kernel:
columns = [local tangent of q (nv) | v (nv) | a (nv) or tau_j (nj) | forces (nf)].
"""
import ctypes

import torch

from ..handle import Handle, _check_in, _ptr


class _Fn:
    def __init__(self, name, value, jac=None):
        self._name, self._value, self._jac = name, value, jac

    def name(self):
        return self._name

    def __call__(self, *args):
        return self._value(*args)

    def jacobian(self, *args):
        if self._jac is None:
            raise NotImplementedError(f"{self._name}: no Jacobian output")
        return self._jac(*args)


class Dynamics:
    dynamics = "whole_body_rnea"    # formulation whose state layout integrate / difference follow

    def __init__(self, model, mass, foot_frames, max_batch=1024, device=None):
        robot = getattr(model, "robot", None)
        if robot is None:
            raise TypeError("model must be robot.model of a pino_locoman_b200 Robot")
        self.robot, self.model, self.mass, self.foot_frames = robot, model, mass, list(foot_frames)
        self.nq, self.nv, self.nj = robot.nq, robot.nv, robot.nq - 7
        self.handle = Handle(robot, self.dynamics, 2, max_batch, tau_nodes=2, device=device)
        self.nf = self.handle.nf

    # -- helpers ------------------------------------------------------------------------------------------
    def _forces(self, forces, ext_force_frame):
        """Pad the force vector when the handle's robot carries an external-force frame the caller left out."""
        if ext_force_frame and ext_force_frame != self.robot.ext_force_frame:
            raise ValueError(f"ext_force_frame {ext_force_frame!r} is not the robot's external-force frame")
        n_in = 12 + (3 if ext_force_frame else 0)
        _check_in(forces, (n_in,), "forces")
        if n_in == self.nf:
            return forces
        pad = torch.zeros(forces.shape[0], self.nf - n_in, dtype=torch.float64, device=forces.device)
        return torch.cat([forces, pad], 1).contiguous()

    def _call(self, fn, outs, *ins):
        h = self.handle
        rc = fn(h._h, *[_ptr(t) for t in ins], ins[0].shape[0], *[_ptr(t) for t in outs], h._stream())
        h._rc(rc)

    # -- state manifold (e.g. dynamics_whole_body_torque.py:11-40) ---------------------------------------------
    def state_integrate(self):
        h = self.handle

        def integrate(x, dx):
            _check_in(x, (h.nx,), "x")
            _check_in(dx, (h.ndx,), "dx")
            out = torch.empty_like(x)
            self._call(h.lib.plm_state_integrate, [out], x, dx)
            return out
        return _Fn("integrate", integrate)

    def state_difference(self):
        h = self.handle

        def difference(x0, x1):
            _check_in(x0, (h.nx,), "x0")
            _check_in(x1, (h.nx,), "x1")
            out = torch.empty(x0.shape[0], h.ndx, dtype=torch.float64, device=x0.device)
            self._call(h.lib.plm_state_difference, [out], x0, x1)
            return out
        return _Fn("difference", difference)

    # -- dynamics.py:33-65 -----------------------------------------------------------------------------------
    def rnea_dynamics(self, ext_force_frame=None):
        h = self.handle

        def run(q, v, a, forces, want_jac):
            f = self._forces(forces, ext_force_frame)
            B = _check_in(q, (self.nq,), "q").shape[0]
            _check_in(v, (self.nv,), "v")
            _check_in(a, (self.nv,), "a")
            tau = torch.empty(B, self.nv, dtype=torch.float64, device=q.device)
            jac = torch.empty(B, self.nv, 3 * self.nv + self.nf, dtype=torch.float64, device=q.device) if want_jac else None
            h._rc(h.lib.plm_rnea_dyn(h._h, _ptr(q), _ptr(v), _ptr(a), _ptr(f), B, _ptr(tau), _ptr(jac), h._stream()))
            return (tau, jac) if want_jac else tau
        return _Fn("rnea_dyn", lambda q, v, a, forces: run(q, v, a, forces, False), lambda q, v, a, forces: run(q, v, a, forces, True))

    # -- dynamics.py:67-118 ----------------------------------------------------------------------------------
    def _frame_args(self, frame_id):
        """(parent body, placement R|p as 12 doubles) of a frame name of the robot's kinematic tree."""
        import numpy as np
        body, T, _ = self.model.frame(frame_id)
        plc = np.concatenate((np.asarray(T[:3, :3], dtype=np.float64).reshape(9), np.asarray(T[:3, 3], dtype=np.float64)))
        return int(body), (ctypes.c_double * 12)(*plc)

    def get_frame_velocity(self, frame_id, relative_to_base=False):
        """frame_vel(q, v) -> [linear; angular] (6) in LOCAL_WORLD_ALIGNED axes, or the base-relative variant of
        dynamics.py:86-113, for any frame of the model."""
        h = self.handle
        body, plc = self._frame_args(frame_id)
        base_body, base_plc = self._frame_args("base_link") if "base_link" in self.model.frames else (0, None)

        def frame_vel(q, v):
            B = _check_in(q, (self.nq,), "q").shape[0]
            _check_in(v, (self.nv,), "v")
            out = torch.empty(B, 6, dtype=torch.float64, device=q.device)
            h._rc(h.lib.plm_frame_kinematics(h._h, body, plc, base_body, base_plc, int(relative_to_base), _ptr(q), _ptr(v), B, None, _ptr(out),
                                             h._stream()))
            return out
        return _Fn("frame_vel", frame_vel)

    def get_frame_position(self, frame_id):
        """frame_pos(q) -> oMf.translation (3) for any frame of the model (dynamics.py:67-75)."""
        h = self.handle
        body, plc = self._frame_args(frame_id)

        def frame_pos(q):
            B = _check_in(q, (self.nq,), "q").shape[0]
            out = torch.empty(B, 3, dtype=torch.float64, device=q.device)
            h._rc(h.lib.plm_frame_kinematics(h._h, body, plc, 0, None, 0, _ptr(q), None, B, _ptr(out), None, h._stream()))
            return out
        return _Fn("frame_pos", frame_pos)

    def frame_velocity_rows(self, frame_id, relative_to_base=False):
        """The three velocity components the OCP rows use (optimization/ocp.py:143,177), read off the node kernel's own rows:
        foot frames (world aligned) and the arm frame (base relative)."""
        h = self.handle
        if not relative_to_base and frame_id in self.foot_frames:
            contact = self.foot_frames.index(frame_id)
        elif relative_to_base and frame_id == self.robot.arm_ee_frame:
            contact = -1
        else:
            raise ValueError("frame_velocity_rows: foot frames (world aligned) and the arm frame (base relative) only")

        def frame_vel(q, v):
            B = _check_in(q, (self.nq,), "q").shape[0]
            _check_in(v, (self.nv,), "v")
            out = torch.empty(B, 3, dtype=torch.float64, device=q.device)
            h._rc(h.lib.plm_frame_vel(h._h, contact, int(relative_to_base), _ptr(q), _ptr(v), B, _ptr(out), h._stream()))
            return out
        return _Fn("frame_vel", frame_vel)

    def _gaps(self, dyn_id, ext_force_frame):
        h = self.handle

        def run(q, v, a, forces, want_jac):
            f = self._forces(forces, ext_force_frame)
            B = _check_in(q, (self.nq,), "q").shape[0]
            gaps = torch.empty(B, 6, dtype=torch.float64, device=q.device)
            jac = torch.empty(B, 6, 3 * self.nv + self.nf, dtype=torch.float64, device=q.device) if want_jac else None
            h._rc(h.lib.plm_dyn_gaps(h._h, dyn_id, _ptr(q), _ptr(v), _ptr(a), _ptr(f), B, _ptr(gaps), _ptr(jac), h._stream()))
            return (gaps, jac) if want_jac else gaps
        return _Fn("dyn_gaps", lambda q, v, a, forces: run(q, v, a, forces, False), lambda q, v, a, forces: run(q, v, a, forces, True))


    def _base_solve(self, name, dyn_id, ext_force_frame):
        """base_acc(q, v, a_j, forces) -> a_b: the 6x6 solve of the base rows of formulation ``dyn_id``."""
        h = self.handle

        def base_acc(q, v, a_j, forces):
            f = self._forces(forces, ext_force_frame)
            B = _check_in(q, (self.nq,), "q").shape[0]
            _check_in(v, (self.nv,), "v")
            _check_in(a_j, (self.nj,), "a_j")
            out = torch.empty(B, 6, dtype=torch.float64, device=q.device)
            h._rc(h.lib.plm_base_solve(h._h, dyn_id, None, _ptr(q), _ptr(v), _ptr(a_j), _ptr(f), B, _ptr(out), h._stream()))
            return out
        return _Fn(name, base_acc)


class DynamicsWholeBodyTorque(Dynamics):
    """dynamics_whole_body_torque.py: rnea_dyn (inherited) and aba_dyn."""

    def aba_dynamics(self, ext_force_frame=None):
        h = self.handle

        def run(q, v, tau_j, forces, want_jac):
            f = self._forces(forces, ext_force_frame)
            B = _check_in(q, (self.nq,), "q").shape[0]
            _check_in(tau_j, (self.nj,), "tau_j")
            a = torch.empty(B, self.nv, dtype=torch.float64, device=q.device)
            jac = torch.empty(B, self.nv, 2 * self.nv + self.nj + self.nf, dtype=torch.float64, device=q.device) if want_jac else None
            h._rc(h.lib.plm_aba_dyn(h._h, _ptr(q), _ptr(v), _ptr(tau_j), _ptr(f), B, _ptr(a), _ptr(jac), h._stream()))
            return (a, jac) if want_jac else a
        return _Fn("aba_dyn", lambda q, v, t, forces: run(q, v, t, forces, False), lambda q, v, t, forces: run(q, v, t, forces, True))


class DynamicsWholeBodyAcc(Dynamics):
    """dynamics_whole_body_acc.py: dyn_gaps = rnea[:6]."""
    dynamics = "whole_body_acc"

    def dynamics_gaps(self, ext_force_frame=None):
        return self._gaps(2, ext_force_frame)

    def base_acc_dynamics(self, ext_force_frame=None):
        """dynamics_whole_body_acc.py:43-83."""
        return self._base_solve("base_acc_dyn", 2, ext_force_frame)


class DynamicsCentroidalAcc(Dynamics):
    """dynamics_centroidal_acc.py: dyn_gaps = A a + Adot v - [sum f + m g; sum (r - c) x f]."""
    dynamics = "centroidal_acc"

    def dynamics_gaps(self, ext_force_frame=None):
        return self._gaps(1, ext_force_frame)

    def base_acc_dynamics(self, ext_force_frame=None):
        """dynamics_centroidal_acc.py:43-82."""
        return self._base_solve("base_acc", 1, ext_force_frame)


class DynamicsCentroidalVel(Dynamics):
    """dynamics_centroidal_vel.py: com_dyn(q, forces), dyn_gaps(h, q, v); state = [h, q]."""
    dynamics = "centroidal_vel"

    def com_dynamics(self, ext_force_frame=None):
        h = self.handle

        def com_dyn(q, forces):
            f = self._forces(forces, ext_force_frame)
            B = _check_in(q, (self.nq,), "q").shape[0]
            out = torch.empty(B, 6, dtype=torch.float64, device=q.device)
            h._rc(h.lib.plm_com_dyn(h._h, _ptr(q), _ptr(f), B, _ptr(out), h._stream()))
            return out
        return _Fn("com_dyn", com_dyn)

    def dynamics_gaps(self):
        h = self.handle

        def dyn_gaps(hh, q, v):
            B = _check_in(q, (self.nq,), "q").shape[0]
            _check_in(hh, (6,), "h")
            _check_in(v, (self.nv,), "v")
            out = torch.empty(B, 6, dtype=torch.float64, device=q.device)
            h._rc(h.lib.plm_centroidal_vel_gaps(h._h, _ptr(hh), _ptr(q), _ptr(v), B, _ptr(out), h._stream()))
            return out
        return _Fn("dyn_gaps", dyn_gaps)

    def base_vel_dynamics(self):
        """dynamics_centroidal_vel.py:73-89: v_b = A_b^-1 (m h - A_j v_j)."""
        h = self.handle

        def base_vel(hh, q, v_j):
            B = _check_in(q, (self.nq,), "q").shape[0]
            _check_in(hh, (6,), "h")
            _check_in(v_j, (self.nj,), "v_j")
            out = torch.empty(B, 6, dtype=torch.float64, device=q.device)
            h._rc(h.lib.plm_base_solve(h._h, 0, _ptr(hh), _ptr(q), None, _ptr(v_j), None, B, _ptr(out), h._stream()))
            return out
        return _Fn("base_vel", base_vel)

    def base_acc_dynamics(self, ext_force_frame=None):
        """dynamics_centroidal_vel.py:91-134 (the centroidal_acc base rows)."""
        return self._base_solve("base_acc", 1, ext_force_frame)
