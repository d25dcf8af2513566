"""OCP assembly: x/p layouts, g/lbg/ubg, f/grad_f/hess_diag, J_g (oracle; test infrastructure only).

Restates optimization/ocp.py:38-198 (variables, parameters, shared constraints), :265-296
(sqp_data / f_data / g_data / hess_diag) and the five optimization/ocp_*.py formulations
(set_weights, setup_variables, setup_parameters, setup_targets, setup_objective,
setup_dynamics_constraints, warm_start, retract_stacked_sol) plus ocp_args.py defaults.
casadi.Opti semantics [3P] per SURVEY.md 3.4 / appendix A.7: x and p are concatenations in creation
order (matrices column-major), g rows in subject_to order with the canonical forms listed there.
The reference obtains J_g from casadi AD; here it comes from complex-step differentiation of the
same residual code, one shooting node at a time (rows of node i only touch DX_i, U_i, DX_{i+1}).
"""
import numpy as np

from . import dynamics as dyn_mod
from .gait import GaitSequence, get_spline_vel_z, opti_dts

OCP_ARGS = {  # ocp_args.py:2-19
    "centroidal_vel": {"include_base": True},
    "centroidal_acc": {"include_base": True},
    "whole_body_acc": {"include_base": True},
    "whole_body_aba": {},
    "whole_body_rnea": {"tau_nodes": 3, "include_acc": True},
}
INF = np.inf
MU = 0.7  # ocp.py:103


class OracleOCP:
    def __init__(self, robot, dynamics, nodes, gait_type="trot", gait_period=0.8, **kwargs):
        if dynamics not in OCP_ARGS:
            raise ValueError(f"Unknown dynamics type: {dynamics}")  # ocp_factory.py:17-18
        args = dict(OCP_ARGS[dynamics])
        args.update(kwargs)
        # inputs without the base part (v_b / a_b follow from the dynamics, e.g.
        # ocp_centroidal_vel.py:19-23,104-120) and whole_body_rnea without accelerations (finite differences,
        # ocp_whole_body_rnea.py:21-25,156,183-191)
        self.include_base = bool(args.get("include_base", True))
        self.include_acc = bool(args.get("include_acc", True))
        self.robot, self.kind, self.nodes = robot, dynamics, nodes
        self.model = robot.model
        self.gait = GaitSequence(gait_type, gait_period)
        self.foot_frames = robot.foot_frames
        self.ext_force_frame, self.arm_ee_frame = robot.ext_force_frame, robot.arm_ee_frame
        self.nq, self.nv, self.nj, self.nf = robot.nq, robot.nv, robot.nj, robot.nf
        self.mass = robot.mass
        nq, nv, nj, nf, N = self.nq, self.nv, self.nj, self.nf, nodes
        cls = {"centroidal_vel": dyn_mod.DynamicsCentroidalVel, "centroidal_acc": dyn_mod.DynamicsCentroidalAcc,
               "whole_body_acc": dyn_mod.DynamicsWholeBodyAcc, "whole_body_aba": dyn_mod.DynamicsWholeBodyTorque,
               "whole_body_rnea": dyn_mod.DynamicsWholeBodyTorque}[dynamics]
        self.dyn = cls(self.model, self.mass, self.foot_frames, robot.base_frame)
        self.tau_nodes = args.get("tau_nodes", 0) if dynamics == "whole_body_rnea" else 0
        # --- setup_variables of each ocp_*.py
        self.n_lead = nv        # leading input block: v / a (nv), without the base part nj, tau_j (nj), nothing
        if dynamics in ("centroidal_vel", "centroidal_acc", "whole_body_acc") and not self.include_base:
            self.n_lead = nj
        if dynamics == "whole_body_aba":
            self.n_lead = nj
        if dynamics == "whole_body_rnea" and not self.include_acc:
            self.n_lead = 0
        if dynamics == "centroidal_vel":
            self.nx, self.ndx = 6 + nq, 6 + nv
            self.x_nom = np.concatenate((np.zeros(6), robot.q0))
            self.nu = [self.n_lead + nf] * N
            self.f_idx = self.n_lead
        else:
            self.nx, self.ndx = nq + nv, 2 * nv
            self.x_nom = np.concatenate((robot.q0, np.zeros(nv)))
            if dynamics == "whole_body_aba":
                self.nu = [nj + nf] * N
                self.f_idx = nj
            elif dynamics == "whole_body_rnea":
                na = self.n_lead
                self.nu = [na + nf + nj] * self.tau_nodes + [na + nf] * (N - self.tau_nodes)
                self.f_idx = na
                self.tau_idx = na + nf
            else:
                self.nu = [self.n_lead + nf] * N
                self.f_idx = self.n_lead
        self.x_off = np.concatenate(([0], np.cumsum([self.ndx + u for u in self.nu])))  # stage offsets
        self.n = int(self.x_off[-1]) + self.ndx
        # --- parameter layout, creation order of ocp.py:54-69 (+ rnea :88-89)
        lay = [("x_init", self.nx), ("dt_min", 1), ("dt_max", 1), ("contact_schedule", 4 * N),
               ("swing_schedule", 4 * N), ("n_contacts", 1), ("swing_period", 1), ("swing_height", 1),
               ("swing_vel_limits", 2), ("Q_diag", self.ndx), ("R_diag", self.nu[0]), ("base_vel_des", 6),
               ("ext_force_des", 3), ("arm_vel_des", 3)]
        if dynamics == "whole_body_rnea":
            lay += [("tau_prev", nj), ("W_diag", nj)]
        self.p_layout, off = {}, 0
        for name, sz in lay:
            self.p_layout[name] = (off, sz)
            off += sz
        self.np_ = off
        self.params = {name: np.zeros(sz) for name, sz in lay}
        self.params["n_contacts"][:] = self.gait.n_contacts  # ocp.py:160
        self.set_weights()
        # rows per node
        self.first_node_skip = dynamics != "centroidal_vel"   # ocp.py:137,170
        self.row_off = [self.ndx]
        for i in range(N):
            self.row_off.append(self.row_off[-1] + self._node_row_count(i))
        self.m = self.row_off[-1]
        self.DX_prev = self.U_prev = None

    # ------------------------------------------------------------------ parameters
    def set_weights(self):
        nj, nf = self.nj, self.nf
        Q_base = [0, 0, 1000, 10000, 10000, 0]
        Q_joint = np.tile([1000, 500, 500], 4)
        if self.arm_ee_frame:
            Q_joint = np.concatenate((Q_joint, [100] * 6))
        Q_vel = np.concatenate(([2000, 2000, 1000, 1000, 1000, 2000], [1] * nj))
        if self.kind == "centroidal_vel":   # ocp_centroidal_vel.py:25-49
            Q = np.concatenate(([1000] * 6, Q_base, Q_joint))
            R = np.concatenate(([1] * self.n_lead, [1e-3] * nf))
        else:
            Q = np.concatenate((Q_base, Q_joint, Q_vel))
            if self.kind == "whole_body_aba":   # ocp_whole_body_aba.py:44-47
                R = np.concatenate(([1e-3] * nj, [1e-3] * nf))
            elif self.kind == "whole_body_rnea":  # ocp_whole_body_rnea.py:50-59
                R = np.concatenate(([1e-3] * self.n_lead, [1e-3] * nf, [1e-4] * nj))
                self.params["W_diag"][:] = 0
            else:
                R = np.concatenate(([1e-3] * self.n_lead, [1e-3] * nf))
        self.params["Q_diag"][:] = Q
        self.params["R_diag"][:] = R

    def set_time_params(self, dt_min, dt_max):
        self.params["dt_min"][:] = dt_min
        self.params["dt_max"][:] = dt_max

    def set_swing_params(self, swing_height, swing_vel_limits):
        self.params["swing_height"][:] = swing_height
        self.params["swing_vel_limits"][:] = swing_vel_limits

    def set_tracking_targets(self, base_vel_des, ext_force_des=None, arm_vel_des=None):
        self.params["base_vel_des"][:] = base_vel_des
        if self.ext_force_frame:
            self.params["ext_force_des"][:] = ext_force_des
        if self.arm_ee_frame:
            self.params["arm_vel_des"][:] = arm_vel_des

    def update_initial_state(self, x_init):
        self.params["x_init"][:] = x_init

    def update_previous_torques(self, tau_prev):
        self.params["tau_prev"][:] = tau_prev

    def dts(self, p=None):
        P = self.params if p is None else self.unpack_p(p)
        return opti_dts(float(P["dt_min"][0]), float(P["dt_max"][0]), self.nodes)

    def update_gait_sequence(self, t_current):  # ocp.py:234-242
        contact, swing = self.gait.get_gait_schedule(t_current, self.dts(), self.nodes)
        self.params["contact_schedule"][:] = contact.flatten(order="F")
        self.params["swing_schedule"][:] = swing.flatten(order="F")
        self.params["n_contacts"][:] = self.gait.n_contacts
        self.params["swing_period"][:] = self.gait.swing_period

    def p_vector(self):
        p = np.zeros(self.np_)
        for name, (off, sz) in self.p_layout.items():
            p[off:off + sz] = self.params[name]
        return p

    def unpack_p(self, p):
        return {name: p[off:off + sz] for name, (off, sz) in self.p_layout.items()}

    # ------------------------------------------------------------------ targets / initial guess
    def targets(self, P):
        """dx_des, u_des of setup_targets (e.g. ocp_whole_body_rnea.py:91-106)."""
        q0 = self.robot.q0
        if self.kind == "centroidal_vel":
            x_des = np.concatenate((P["base_vel_des"], q0))
        else:
            x_des = np.concatenate((q0, P["base_vel_des"], np.zeros(self.nj)))
        dx_des = self.dyn.state_difference()(np.asarray(P["x_init"], dtype=float), x_des)
        fg = 9.81 * self.mass
        nc = float(P["n_contacts"][0])
        f_des = np.zeros(self.nf)
        f_des[2] = f_des[5] = 0.8 * fg / nc
        f_des[8] = f_des[11] = 1.2 * fg / nc
        u_des = np.concatenate((np.zeros(self.n_lead), f_des))
        if self.kind == "whole_body_rnea":
            u_des = np.concatenate((u_des, np.zeros(self.nj)))
        return dx_des, u_des, f_des

    def initial_guess(self):
        """opti.initial(): DX = 0, U_i = u_des[:nu_i] (ocp.py:159-163,193)."""
        _, u_des, _ = self.targets(self.params)
        x = np.zeros(self.n)
        for i in range(self.nodes):
            o = self.x_off[i] + self.ndx
            x[o:o + self.nu[i]] = u_des[:self.nu[i]]
        return x

    def split(self, x, i):
        o = self.x_off[i]
        dx = x[..., o:o + self.ndx]
        if i == self.nodes:
            return dx, None
        return dx, x[..., o + self.ndx:o + self.ndx + self.nu[i]]

    # ------------------------------------------------------------------ objective
    def _weights(self, P):
        w = np.zeros(self.n)
        for i in range(self.nodes + 1):
            o = self.x_off[i]
            w[o:o + self.ndx] = P["Q_diag"]
            if i < self.nodes:
                w[o + self.ndx:o + self.ndx + self.nu[i]] = P["R_diag"][:self.nu[i]]
        return w

    def _targets_stacked(self, P):
        dx_des, u_des, _ = self.targets(P)
        t = np.zeros(self.n)
        for i in range(self.nodes + 1):
            o = self.x_off[i]
            t[o:o + self.ndx] = dx_des
            if i < self.nodes:
                t[o + self.ndx:o + self.ndx + self.nu[i]] = u_des[:self.nu[i]]
        return t

    def f_data(self, x, p):
        """f_data(x,p) -> [f, grad_f] (ocp.py:289)."""
        P = self.unpack_p(p)
        w, t = self._weights(P), self._targets_stacked(P)
        e = x - t
        f = float(np.sum(w * e * e))
        grad = 2 * w * e
        if self.kind == "whole_body_rnea" and self.tau_nodes > 0:  # ocp_whole_body_rnea.py:125-129
            o = self.x_off[0] + self.ndx + self.tau_idx
            et = x[o:o + self.nj] - P["tau_prev"]
            f += float(np.sum(P["W_diag"] * et * et))
            grad[o:o + self.nj] += 2 * P["W_diag"] * et
        return f, grad

    def hess_diag(self, p):
        """diag(hess_data(x,p)) (ocp.py:293-296): constant, diagonal."""
        P = self.unpack_p(p)
        h = 2 * self._weights(P)
        if self.kind == "whole_body_rnea" and self.tau_nodes > 0:
            o = self.x_off[0] + self.ndx + self.tau_idx
            h[o:o + self.nj] += 2 * P["W_diag"]
        return h

    # ------------------------------------------------------------------ constraints
    def _node_row_count(self, i):
        nv, nj = self.nv, self.nj
        k = self.kind
        if k == "centroidal_vel":
            rows = 6 + nv + (6 if self.include_base else 0)
        elif k == "whole_body_aba":
            rows = 2 * nv
        elif k == "whole_body_rnea":
            rows = (2 * nv if self.include_acc else nv) + 6 + (2 * nj if i < self.tau_nodes else 0)
        else:
            rows = 2 * nv + (6 if self.include_base else 0)
        skip = i == 0 and self.first_node_skip
        rows += 4 * (5 + (0 if skip else 3))
        if self.ext_force_frame:
            rows += 3
        if not skip:
            if self.arm_ee_frame:
                rows += 3
            rows += 2 * nj
        return rows

    def node_rows(self, i, dx, u, dx_next, P):
        """Rows of shooting node i in subject_to order (ocp.py:111-190) -> (g, lb, ub)."""
        nq, nv, nj, nf = self.nq, self.nv, self.nj, self.nf
        x_init = np.asarray(P["x_init"])
        dt = self.dts_cache[i]
        x = self.dyn.state_integrate()(x_init, dx)
        g, lb, ub = [], [], []

        def eq(val, rhs=0.0):
            g.append(val)
            lb.append(np.zeros(val.shape[-1]) + rhs)
            ub.append(np.zeros(val.shape[-1]) + rhs)

        forces = u[..., self.f_idx:self.f_idx + nf]
        if self.kind == "centroidal_vel":   # ocp_centroidal_vel.py:85-107
            h, q = x[..., :6], x[..., 6:]
            if self.include_base:
                v = u[..., :nv]
            else:   # ocp_centroidal_vel.py:109-120: base velocity from the momentum
                v_j = u[..., :nj]
                v = np.concatenate((self.dyn.base_vel_dynamics()(h, q, v_j), v_j), -1)
            h_dot = self.dyn.com_dynamics(self.ext_force_frame)(q, forces)
            eq(dx_next[..., :6] - (dx[..., :6] + h_dot * dt))
            eq(dx_next[..., 6:] - (dx[..., 6:] + v * dt))
            if self.include_base:
                eq(self.dyn.dynamics_gaps()(h, q, v))
        else:
            q, v = x[..., :nq], x[..., nq:]
            dq, dv = dx[..., :nv], dx[..., nv:]
            if self.kind == "whole_body_aba":   # ocp_whole_body_aba.py:86-106
                tau_j = u[..., :nj]
                a = self.dyn.aba_dynamics(self.ext_force_frame)(q, v, tau_j, forces)
            elif self.kind == "whole_body_rnea" and not self.include_acc:   # ocp_whole_body_rnea.py:183-191
                v_next = self.dyn.state_integrate()(x_init, dx_next)[..., nq:]
                a = (v_next - v) / dt
            elif not self.include_base:   # ocp_centroidal_acc.py:129-140, ocp_whole_body_acc.py:130-141
                a_j = u[..., :nj]
                a = np.concatenate((self.dyn.base_acc_dynamics(self.ext_force_frame)(q, v, a_j, forces), a_j), -1)
            else:
                a = u[..., :nv]
            eq(dx_next[..., :nv] - (dq + v * dt))
            if self.include_acc:
                eq(dx_next[..., nv:] - (dv + a * dt))
            if self.kind == "whole_body_rnea":   # ocp_whole_body_rnea.py:160-171
                tau = self.dyn.rnea_dynamics(self.ext_force_frame)(q, v, a, forces)
                eq(tau[..., :6])
                if i < self.tau_nodes:
                    tau_j = u[..., self.tau_idx:]
                    eq(tau[..., 6:] - tau_j)
                    g.append(tau_j)
                    lb.append(-self.robot.joint_torque_max)
                    ub.append(self.robot.joint_torque_max)
            elif self.kind in ("whole_body_acc", "centroidal_acc") and self.include_base:
                eq(self.dyn.dynamics_gaps(self.ext_force_frame)(q, v, a, forces))

        skip = i == 0 and self.first_node_skip
        contact = np.asarray(P["contact_schedule"]).reshape(self.nodes, 4)[i]   # column-major (4, N)
        swing = np.asarray(P["swing_schedule"]).reshape(self.nodes, 4)[i]
        for idx, fid in enumerate(self.foot_frames):   # ocp.py:121-157
            f_e = forces[..., 3 * idx:3 * idx + 3]
            c = contact[idx]
            g.append(c * f_e[..., 2:3])
            lb.append(np.zeros(1))
            ub.append(np.full(1, INF))
            g.append(c * (f_e[..., 0:1] ** 2 + f_e[..., 1:2] ** 2) - c * MU ** 2 * f_e[..., 2:3] ** 2)
            lb.append(np.full(1, -INF))
            ub.append(np.zeros(1))
            eq((1 - c) * f_e)
            if skip:
                continue
            vel = self.dyn.get_frame_velocity(fid, relative_to_base=False)(q, v)
            eq(c * vel[..., :2])
            vz_des = get_spline_vel_z(swing[idx], float(P["swing_period"][0]), float(P["swing_height"][0]),
                                      float(P["swing_vel_limits"][0]), float(P["swing_vel_limits"][1]))
            eq(c * vel[..., 2:3] + (1 - c) * (vel[..., 2:3] - vz_des))
        if self.ext_force_frame:   # ocp.py:166-168
            eq(forces[..., 12:15], np.asarray(P["ext_force_des"], dtype=float))
        if not skip:
            if self.arm_ee_frame:   # ocp.py:176-180
                vel = self.dyn.get_frame_velocity(self.arm_ee_frame, relative_to_base=True)(q, v)
                eq(vel[..., :3] - np.asarray(P["arm_vel_des"], dtype=float))
            g.append(q[..., 7:])   # ocp.py:183-190
            lb.append(self.robot.joint_pos_min)
            ub.append(self.robot.joint_pos_max)
            g.append(v[..., 6:])
            lb.append(-self.robot.joint_vel_max)
            ub.append(self.robot.joint_vel_max)
        return np.concatenate(g, -1), np.concatenate(lb), np.concatenate(ub)

    def g_data(self, x, p):
        """g_data(x,p) -> [g, lbg, ubg] (ocp.py:290)."""
        P = self.unpack_p(p)
        self.dts_cache = self.dts(p)
        g = [x[:self.ndx]]
        lb = [np.zeros(self.ndx)]
        ub = [np.zeros(self.ndx)]
        for i in range(self.nodes):
            dx, u = self.split(x, i)
            dxn, _ = self.split(x, i + 1)
            gi, li, ui = self.node_rows(i, dx, u, dxn, P)
            g.append(gi)
            lb.append(li)
            ub.append(ui)
        return np.concatenate(g), np.concatenate(lb), np.concatenate(ub)

    def jac_g(self, x, p, h=1e-30):
        """Dense J_g (m x n) by complex step, node by node."""
        P = self.unpack_p(p)
        self.dts_cache = self.dts(p)
        J = np.zeros((self.m, self.n))
        J[:self.ndx, :self.ndx] = np.eye(self.ndx)
        for i in range(self.nodes):
            dx, u = self.split(x, i)
            dxn, _ = self.split(x, i + 1)
            z = np.concatenate((dx, u, dxn))
            K = z.size
            Z = np.tile(z.astype(complex), (K, 1)) + 1j * h * np.eye(K)
            gi, _, _ = self.node_rows(i, Z[:, :self.ndx], Z[:, self.ndx:self.ndx + self.nu[i]],
                                      Z[:, self.ndx + self.nu[i]:], P)
            Ji = np.imag(gi).T / h          # rows x K
            r0, o = self.row_off[i], self.x_off[i]
            J[r0:r0 + Ji.shape[0], o:o + K] = Ji
        return J

    def sqp_data(self, x, p):
        """sqp_data(x,p) -> [grad_f, J_g, g, lbg, ubg] (ocp.py:287), J_g dense."""
        _, grad = self.f_data(x, p)
        g, lb, ub = self.g_data(x, p)
        return grad, self.jac_g(x, p), g, lb, ub

    # ------------------------------------------------------------------ warm start / retract
    def retract_stacked_sol(self, sol_x):
        """DX_prev / U_prev of retract_stacked_sol (e.g. ocp_whole_body_rnea.py:293-324)."""
        self.DX_prev = [np.array(self.split(sol_x, i)[0]) for i in range(self.nodes + 1)]
        self.U_prev = [np.array(self.split(sol_x, i)[1]) for i in range(self.nodes)]

    def warm_start(self):
        """Initial guess from the previous solution (e.g. ocp_whole_body_rnea.py:207-235)."""
        x = self.initial_guess()
        if self.DX_prev is None:
            return x
        _, _, f_des0 = self.targets(self.params)
        contact = self.params["contact_schedule"].reshape(self.nodes, 4)
        for i in range(self.nodes + 1):
            o = self.x_off[i]
            x[o:o + self.ndx] = self.DX_prev[i]
        for i in range(self.nodes):
            f_des = f_des0.copy()
            for j in range(4):
                if contact[i, j] == 0:
                    f_des[3 * j:3 * j + 3] = 0
            u_prev = self.U_prev[i]
            lead = u_prev[:self.f_idx]
            u = np.concatenate((lead, f_des))
            if self.kind == "whole_body_rnea" and i < self.tau_nodes:
                u = np.concatenate((u, u_prev[self.tau_idx:]))
            o = self.x_off[i] + self.ndx
            x[o:o + self.nu[i]] = u
        return x
