"""Batched OCP front end: the reference's ``optimization/ocp.py`` plugin surface over the CUDA library.

Mirrors ``OCP`` (optimization/ocp.py:11-496): same method names and argument meaning (``setup_problem``,
``set_weights``, ``set_time_params``, ``set_swing_params``, ``set_tracking_targets``, ``update_initial_state``,
``update_gait_sequence``, ``warm_start``, ``init_solver``, ``solve``, ``retract_stacked_sol``) and the same
attributes (``x_nom``, ``dts``, ``DX_prev``, ``U_prev``, ``hess_diag``, ``solve_time``, ``sqp_data`` / ``f_data`` /
``g_data`` / ``hess_data``), generalised to a batch of independent MPC instances: every setter accepts either
the reference's per-problem value or an array with a leading batch dimension.  Nothing is symbolic: the casadi
Functions become calls into libpinolocoman_b200.so on batched ``torch.float64`` CUDA tensors; the parameter
vector ``p`` (creation order of optimization/ocp.py:54-69) is assembled on the host.
"""
import time

import numpy as np
import torch

from ..handle import Handle
from ..utils.gait_sequence import horizon_dts


def _se3_exp(nu):
    """Host-side exp6 for retraction (x_init (+) dx), batched over the leading dimension."""
    rho, w = nu[..., :3], nu[..., 3:]
    t2 = np.sum(w * w, -1)
    small = t2 < 1e-2
    t = np.sqrt(np.where(small, 1.0, t2))
    A = np.where(small, 1 - t2 / 6 + t2**2 / 120 - t2**3 / 5040 + t2**4 / 362880, np.sin(t) / t)
    B = np.where(small, 0.5 - t2 / 24 + t2**2 / 720 - t2**3 / 40320 + t2**4 / 3628800, (1 - np.cos(t)) / np.where(small, 1.0, t2))
    C = np.where(small, 1 / 6 - t2 / 120 + t2**2 / 5040 - t2**3 / 362880 + t2**4 / 39916800, (t - np.sin(t)) / (np.where(small, 1.0, t2) * t))
    wxr = np.cross(w, rho)
    p = rho + B[..., None] * wxr + C[..., None] * np.cross(w, wxr)
    half = 0.5 * np.sqrt(t2)
    s = np.where(small, 0.5 * (1 - t2 / 24 + t2**2 / 1920 - t2**3 / 322560), np.sin(half) / t)
    quat = np.concatenate([s[..., None] * w, np.cos(half)[..., None]], -1)
    return p, quat, A


def _quat_mul(a, b):
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def _quat_rotate(q, v):
    u, w = q[..., :3], q[..., 3:4]
    return v + 2 * np.cross(u, np.cross(u, v) + w * v)


def integrate_configuration(q, dq):
    """pin.integrate(model, q, dq) for a free-flyer + revolute joints: base M * exp6(dq[:6]), joints additive."""
    p, quat, _ = _se3_exp(dq[..., :6])
    out = np.empty(np.broadcast_shapes(q.shape[:-1], dq.shape[:-1]) + (q.shape[-1],))
    out[..., :3] = q[..., :3] + _quat_rotate(q[..., 3:7], p)
    qn = _quat_mul(q[..., 3:7], quat)
    out[..., 3:7] = qn / np.linalg.norm(qn, axis=-1, keepdims=True)
    out[..., 7:] = q[..., 7:] + dq[..., 6:]
    return out


class _BatchedFunction:
    """Callable standing in for a casadi.Function with signature (x, p) -> outputs, batched."""

    def __init__(self, name, fn, names_out):
        self._name, self._fn, self._names_out = name, fn, names_out

    def name(self):
        return self._name

    def name_out(self):
        return list(self._names_out)

    def __call__(self, x, p):
        return self._fn(x, p)


class OCP:
    dynamics = None  # set by subclasses

    def __init__(self, robot, solver, nodes, batch=1, device=None):
        if robot.gait_sequence is None:
            raise ValueError("robot.set_gait_sequence(gait_type, gait_period) must be called before building the OCP")
        self.robot = robot
        self.model = robot.model
        self.gait_sequence = robot.gait_sequence
        self.foot_frames = robot.foot_frames
        self.ext_force_frame = robot.ext_force_frame
        self.arm_ee_frame = robot.arm_ee_frame
        self.n_feet = len(self.foot_frames)
        self.nq, self.nv, self.nf, self.nj = robot.nq, robot.nv, robot.nf, robot.nj
        self.solver = solver
        self.nodes = nodes
        self.mass = robot.mass
        self.batch = int(batch)
        self.device = device
        self.tau_nodes = 0
        self.q_sol, self.v_sol, self.a_sol, self.forces_sol = [], [], [], []
        self.DX_prev = self.U_prev = self.lam_g = None
        self.solve_time = 0.0
        self.stats = None

    # ------------------------------------------------------------------ problem setup
    def setup_problem(self):
        self.setup_variables()
        self.setup_parameters()

    def setup_variables(self):
        """Create the library handle: fixes the x layout [DX_0, U_0, ..., DX_N] and the row layout of g."""
        self.handle = Handle(self.robot, self.dynamics, self.nodes, 0 if self.device == "layout" else self.batch,
                             tau_nodes=max(self.tau_nodes, 1), device=None if self.device == "layout" else self.device,
                             include_base=getattr(self, "include_base", True), include_acc=getattr(self, "include_acc", True))
        h = self.handle
        self.nx, self.ndx_opt, self.nu_opt = h.nx, h.ndx, list(h.nu)
        self.n, self.m = h.n, h.m

    def setup_parameters(self):
        """Host image of the Opti parameter vector (optimization/ocp.py:50-74), one row per instance."""
        h = self.handle
        self._p = np.zeros((self.batch, h.np))
        self._p_dirty = True
        self._set("n_contacts", self.gait_sequence.n_contacts)      # optimization/ocp.py:160
        self._set("swing_period", self.gait_sequence.swing_period)
        self._x0 = None

    def _slot(self, name, size=None):
        off = self.handle.p_off[name]
        if off < 0:
            raise KeyError(f"parameter {name} does not exist for {self.dynamics}")
        sizes = {"x_init": self.nx, "dt_min": 1, "dt_max": 1, "contact_schedule": 4 * self.nodes, "swing_schedule": 4 * self.nodes,
                 "n_contacts": 1, "swing_period": 1, "swing_height": 1, "swing_vel_limits": 2, "Q_diag": self.ndx_opt,
                 "R_diag": self.nu_opt[0], "base_vel_des": 6, "ext_force_des": 3, "arm_vel_des": 3, "tau_prev": self.nj,
                 "W_diag": self.nj}
        return off, sizes[name]

    def _set(self, name, value):
        off, size = self._slot(name)
        v = np.asarray(value, dtype=np.float64)
        if v.ndim == 0:
            v = v.reshape(1)
        if v.shape[-1] != size:
            raise ValueError(f"{name}: expected {size} values, got shape {v.shape}")
        self._p[:, off:off + size] = v      # broadcasts a per-problem value over the batch
        self._p_dirty = True

    def _get(self, name):
        off, size = self._slot(name)
        return self._p[:, off:off + size]

    def set_weights(self):
        raise NotImplementedError

    def set_time_params(self, dt_min, dt_max):
        self._set("dt_min", dt_min)
        self._set("dt_max", dt_max)
        self.dts = horizon_dts(float(np.ravel(dt_min)[0]), float(np.ravel(dt_max)[0]), self.nodes)

    def set_swing_params(self, swing_height, swing_vel_limits):
        self._set("swing_height", swing_height)
        self._set("swing_vel_limits", swing_vel_limits)

    def set_tracking_targets(self, base_vel_des, ext_force_des=None, arm_vel_des=None):
        self._set("base_vel_des", base_vel_des)
        if self.ext_force_frame:
            self._set("ext_force_des", ext_force_des)
        if self.arm_ee_frame:
            self._set("arm_vel_des", arm_vel_des)

    def update_initial_state(self, x_init):
        self._set("x_init", x_init)

    def update_gait_sequence(self, t_current):
        """Contact / swing schedules for the horizon; ``t_current`` scalar or one start time per instance."""
        t = np.broadcast_to(np.asarray(t_current, dtype=np.float64), (self.batch,))
        contact, swing = self.gait_sequence.get_gait_schedule(t, self.dts, self.nodes)      # [B, 4, N]
        self._set("contact_schedule", contact.transpose(0, 2, 1).reshape(self.batch, -1))   # column-major (4, N)
        self._set("swing_schedule", swing.transpose(0, 2, 1).reshape(self.batch, -1))
        self._set("n_contacts", self.gait_sequence.n_contacts)
        self._set("swing_period", self.gait_sequence.swing_period)

    # ------------------------------------------------------------------ initial guess / warm start
    def f_des(self):
        """Desired contact forces of setup_targets: 0.8/1.2 m g / n_contacts on z for front/rear feet."""
        nc = self._get("n_contacts")[:, 0]
        fg = 9.81 * self.mass
        f = np.zeros((self.batch, self.nf))
        f[:, 2] = f[:, 5] = 0.8 * fg / nc
        f[:, 8] = f[:, 11] = 1.2 * fg / nc
        return f

    def _lead(self):
        return self.handle.nu[-1] - self.nf      # leading input block (a, v or tau_j)

    def initial_guess(self):
        """opti.initial(): DX = 0, U_i = u_des[:nu_i] (optimization/ocp.py:159-163,193)."""
        x = np.zeros((self.batch, self.n))
        lead, f = self._lead(), self.f_des()
        for i in range(self.nodes):
            o = self.handle.x_off[i] + self.ndx_opt + lead
            x[:, o:o + self.nf] = f
        return x

    def warm_start(self):
        """Previous solution for DX and the leading inputs (un-shifted), contact-masked f_des for the forces."""
        x = self.initial_guess()
        if self.DX_prev is None:
            self._x0 = x
            return
        lead, h = self._lead(), self.handle
        contact = self._get("contact_schedule").reshape(self.batch, self.nodes, 4)
        f_des = self.f_des()
        for i in range(self.nodes + 1):
            o = h.x_off[i]
            x[:, o:o + self.ndx_opt] = self.DX_prev[i]
        for i in range(self.nodes):
            o = h.x_off[i] + self.ndx_opt
            u_prev = self.U_prev[i]
            f = f_des.copy()
            for j in range(self.n_feet):
                f[:, 3 * j:3 * j + 3] *= (contact[:, i, j] != 0)[:, None]
            x[:, o:o + lead] = u_prev[:, :lead]
            x[:, o + lead:o + lead + self.nf] = f
            if self.nu_opt[i] > lead + self.nf:
                x[:, o + lead + self.nf:o + self.nu_opt[i]] = u_prev[:, lead + self.nf:]
        self._x0 = x

    def set_initial(self, x):
        """opti.set_initial(opti.x, x): stacked starting point [batch, n] (or [n]) of the following solve() calls."""
        x = np.asarray(x, dtype=np.float64)
        self._x0 = np.broadcast_to(x, (self.batch, self.n)).copy()

    # ------------------------------------------------------------------ solver
    def _p_device(self):
        if self._p_dirty or getattr(self, "_p_dev", None) is None:
            if getattr(self, "_pin_p", None) is None:
                self._pin_p = torch.empty(self.batch, self.handle.np, dtype=torch.float64)
                if self.handle.device.type == "cuda":
                    self._pin_p = self._pin_p.pin_memory()
            self._pin_p.numpy()[:] = self._p
            self._p_dev = self._pin_p.to(self.handle.device, non_blocking=True)
            self._p_dirty = False
        return self._p_dev

    def init_solver(self):
        if self.solver == "osqp":
            h = self.handle
            self.osqp_opts = {"max_iter": 100, "alpha": 1.4, "rho": 2e-2, "warm_start": True, "adaptive_rho": False}
            self.sqp_data = _BatchedFunction("sqp_data", lambda x, p: h.sqp_data(x, p), ["grad_f", "J_g", "g", "lbg", "ubg"])
            self.f_data = _BatchedFunction("f_data", lambda x, p: h.f_data(x, p), ["f", "grad_f"])
            self.g_data = _BatchedFunction("g_data", lambda x, p: h.g_data(x, p), ["g", "lbg", "ubg"])
            self.hess_data = _BatchedFunction("hess_data", lambda x, p: (h.hess_diag(p),), ["hess_f"])
            hess = h.hess_diag(self._p_device())
            self.hess_diag = hess[0].cpu().numpy()       # constant diagonal (optimization/ocp.py:293-296)
            h.qp_setup(hess)                             # osqp setup with dummy data (optimization/ocp.py:305-313)
        elif self.solver == "fatrop":
            raise NotImplementedError("the Fatrop interior-point path is CPU-only in the reference and out of scope here")
        else:
            raise ValueError(f"Solver {self.solver} not supported")

    def solve(self, retract_all=True):
        """One SQP iteration for every instance: sqp_data -> OSQP update/solve -> Armijo (optimization/ocp.py:375-422).

        Returns the stacked solution [batch, n] (the reference returns nothing): a view of a pinned host buffer that the
        call after next reuses; ``DX_prev`` / ``U_prev`` are views of the same buffer (valid until the call after next),
        the solution history (``q_sol``, ``a_sol``, ``forces_sol``, ...) holds copies.  The starting point is
        ``opti.initial()``: only ``warm_start()`` moves it."""
        if self.solver != "osqp":
            raise ValueError(f"Solver {self.solver} not supported")
        h = self.handle
        # opti.initial(): the point of the last warm_start() call, else DX = 0, U = u_des (optimization/ocp.py:159-163,
        # 193,379); solve() itself never moves it (the reference's retract_stacked_sol does not call set_initial)
        current_x = self.initial_guess() if self._x0 is None else self._x0
        start_time = time.time()
        if getattr(self, "_pin_x", None) is None:     # pinned staging buffers for the host <-> device copies
            self._pin_x = torch.empty(self.batch, self.n, dtype=torch.float64).pin_memory()
            # two result buffers, used alternately: the solution returned by one call (and the DX_prev / U_prev views
            # of it) stays valid while the next call writes the other one
            self._pin_out = [torch.empty(self.batch, self.n, dtype=torch.float64).pin_memory() for _ in range(2)]
            self._pin_stats = torch.empty(self.batch, 8, dtype=torch.float64).pin_memory()
            self._pin_sel = 0
        self._pin_x.copy_(torch.from_numpy(np.ascontiguousarray(current_x)))      # (multi-threaded host copy into pinned memory)
        xd = self._pin_x.to(h.device, non_blocking=True)
        pd = self._p_device()
        x_new, stats = h.sqp_step(xd, pd)
        pin_out = self._pin_out[self._pin_sel]
        self._pin_sel ^= 1
        pin_out.copy_(x_new, non_blocking=True)
        self._pin_stats.copy_(stats, non_blocking=True)
        torch.cuda.current_stream(h.device).synchronize()
        sol_x = pin_out.numpy()
        self.stats = self._pin_stats.numpy().copy()
        self.solve_time = time.time() - start_time
        self.retract_stacked_sol(sol_x, retract_all)
        return sol_x

    # ------------------------------------------------------------------ solution unpacking
    def state_integrate(self, x, dx):
        """integrate(x, dx) of the formulation's Dynamics class (host copy used for retraction)."""
        nq, nv = self.nq, self.nv
        return np.concatenate([integrate_configuration(x[..., :nq], dx[..., :nv]), x[..., nq:] + dx[..., nv:]], -1)

    def _unpack(self, u):
        """Split an input vector into (a, forces, tau) lists entries for the solution history."""
        raise NotImplementedError

    def retract_stacked_sol(self, sol_x, retract_all=True):
        h = self.handle
        sol_x = np.atleast_2d(sol_x)
        self.DX_prev, self.U_prev = [], []
        x_init = self._get("x_init")
        for i in range(self.nodes):
            o = h.x_off[i]
            dx_sol = sol_x[:, o:o + self.ndx_opt]
            u_sol = sol_x[:, o + self.ndx_opt:o + self.ndx_opt + self.nu_opt[i]]
            self.DX_prev.append(dx_sol)      # views of the solution buffer (reused by the call after next)
            self.U_prev.append(u_sol)
            if i == 0 or retract_all:
                self._append_solution(self.state_integrate(x_init, dx_sol), u_sol)
        dx_last = sol_x[:, h.x_off[self.nodes]:]
        self.DX_prev.append(dx_last)
        if retract_all:
            self._append_state(self.state_integrate(x_init, dx_last))

    def _append_state(self, x_sol):
        self.q_sol.append(x_sol[:, :self.nq])
        self.v_sol.append(x_sol[:, self.nq:])

    def _append_solution(self, x_sol, u_sol):
        self._append_state(x_sol)
        lead = self._lead()
        self.a_sol.append(u_sol[:, :lead].copy())      # copies, as the reference's np.array(...) (history outlives the buffer)
        self.forces_sol.append(u_sol[:, lead:lead + self.nf].copy())

    # violation metrics of optimization/ocp.py:482-496 (host versions, for users of g_data)
    @staticmethod
    def _constraint_violation_metric(g, lbg, ubg):
        v = np.concatenate((np.maximum(0, lbg - g), np.maximum(0, g - ubg)), -1)
        return np.linalg.norm(v, axis=-1)

    @staticmethod
    def _constraint_violation_max(g, lbg, ubg):
        v = np.concatenate((np.maximum(0, lbg - g), np.maximum(0, g - ubg)), -1)
        return np.max(np.abs(v), axis=-1)
