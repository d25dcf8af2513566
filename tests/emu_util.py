"""Helpers for the CPU tests: host emulation of the node kernel + oracle problem builders."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "host_emu")
DYN_ID = {"centroidal_vel": 0, "centroidal_acc": 1, "whole_body_acc": 2, "whole_body_aba": 3, "whole_body_rnea": 4}


class OcpDesc(ctypes.Structure):
    _fields_ = [("dynamics", ctypes.c_int32), ("nodes", ctypes.c_int32), ("tau_nodes", ctypes.c_int32),
                ("mu", ctypes.c_double), ("osqp_max_iter", ctypes.c_int32), ("osqp_check_termination", ctypes.c_int32),
                ("osqp_scaling", ctypes.c_int32), ("osqp_rho", ctypes.c_double), ("osqp_sigma", ctypes.c_double),
                ("osqp_alpha", ctypes.c_double), ("osqp_eps_abs", ctypes.c_double), ("osqp_eps_rel", ctypes.c_double),
                ("osqp_eps_prim_inf", ctypes.c_double), ("osqp_eps_dual_inf", ctypes.c_double),
                ("include_base", ctypes.c_int32), ("include_acc", ctypes.c_int32)]


def default_ocp_desc(dynamics, nodes, tau_nodes=3, include_base=True, include_acc=True):
    return OcpDesc(DYN_ID[dynamics], nodes, tau_nodes, 0.7, 100, 25, 10, 2e-2, 1e-6, 1.4, 1e-3, 1e-3, 1e-4, 1e-4, int(include_base), int(include_acc))


def build_emu():
    so = os.path.join(EMU_DIR, "libplm_emu.so")
    srcs = [os.path.join(EMU_DIR, "plm_emu.cpp"), os.path.join(ROOT, "pino_locoman_b200", "csrc", "plm_host.cpp")]
    deps = srcs + [os.path.join(ROOT, "pino_locoman_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "pino_locoman_b200", "csrc"))
                   if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so] + srcs)
    lib = ctypes.CDLL(so)
    lib.emu_create.restype = ctypes.c_void_p
    return lib


class Emu:
    def __init__(self, robot, dynamics, nodes, tau_nodes=3, include_base=True, include_acc=True):
        from pino_locoman_b200.utils.robot import robot_desc
        self.lib = build_emu()
        self.rd = robot_desc(robot)
        self.od = default_ocp_desc(dynamics, nodes, tau_nodes, include_base, include_acc)
        err = ctypes.create_string_buffer(256)
        self.h = self.lib.emu_create(ctypes.byref(self.rd), ctypes.byref(self.od), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.h = ctypes.c_void_p(self.h)
        d = (ctypes.c_int * 8)()
        self.lib.emu_dims(self.h, d)
        self.n, self.m, self.np_, self.nnz, self.ndx, self.nx, self.nf, self.nodes = list(d)
        rows = np.zeros(self.nnz, dtype=np.int32)
        cols = np.zeros(self.nnz, dtype=np.int32)
        self.lib.emu_pattern(self.h, rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p))
        self.rows, self.cols = rows, cols

    def eval(self, x, p, want_jac=True):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        B = x.shape[0]
        g = np.zeros((B, self.m))
        J = np.zeros((B, self.nnz))
        self.lib.emu_node_eval(self.h, x.ctypes.data_as(ctypes.c_void_p), p.ctypes.data_as(ctypes.c_void_p), B,
                               g.ctypes.data_as(ctypes.c_void_p), J.ctypes.data_as(ctypes.c_void_p), int(want_jac))
        return g, J

    def dense(self, Jv):
        D = np.zeros((self.m, self.n))
        D[self.rows, self.cols] = Jv
        return D


def make_robots():
    from pino_locoman_b200.utils.robot import Go2, B2, B2G
    from oracle.model import OracleRobot
    prod = {"go2": Go2(), "b2": B2(), "b2g": B2G()}
    for r in prod.values():
        r.set_gait_sequence("trot", 0.8)
    ora = {k: OracleRobot(k) for k in prod}
    return prod, ora


def random_problem(oocp, rng, t_current=None, ext=True):
    """Fill an OracleOCP with the synthetic distributions of SURVEY.md 8(d); returns (x, p)."""
    o = oocp
    r = o.robot
    o.set_time_params(0.01, 0.08)
    o.set_swing_params(0.07, [0.1, -0.2])
    o.set_tracking_targets([0.2, 0, 0, 0, 0, 0], rng.uniform(-20, 20, 3) if ext else np.zeros(3),
                           rng.uniform(-0.2, 0.2, 3) if ext else np.zeros(3))
    q = r.q0.copy()
    q[0:2] = rng.uniform(-1, 1, 2)
    q[2] = rng.uniform(0.45, 0.65) if r.name != "go2" else rng.uniform(0.28, 0.40)
    yaw, roll, pitch = rng.uniform(-np.pi, np.pi), rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3)
    from oracle.spatial import rpy_to_R, R_to_quat
    q[3:7] = R_to_quat(rpy_to_R(roll, pitch, yaw))
    lo, hi = r.joint_pos_min, r.joint_pos_max
    qj = rng.uniform(lo, hi)
    q[7:] = r.q0[7:] + 0.9 * (qj - r.q0[7:])
    v = np.concatenate((rng.uniform(-1, 1, 6), rng.uniform(-0.25, 0.25, r.nj) * r.joint_vel_max))
    if o.kind == "centroidal_vel":
        x_init = np.concatenate((rng.uniform(-0.5, 0.5, 6), q))
    else:
        x_init = np.concatenate((q, v))
    o.update_initial_state(x_init)
    k = int(rng.integers(0, 80)) if t_current is None else t_current
    o.update_gait_sequence(k * 0.01)
    if o.kind == "whole_body_rnea":
        o.update_previous_torques(rng.uniform(-0.1, 0.1, r.nj) * r.joint_torque_max)
    p = o.p_vector()
    x = o.initial_guess()
    contact = o.params["contact_schedule"].reshape(o.nodes, 4)
    mg = o.mass * 9.81
    for i in range(o.nodes + 1):
        off = o.x_off[i]
        if i > 0:
            x[off:off + o.ndx] = rng.normal(0, 0.05, o.ndx)
        if i == o.nodes:
            break
        u = x[off + o.ndx:off + o.ndx + o.nu[i]]
        lead = o.f_idx
        if o.kind == "whole_body_aba":
            u[:lead] = rng.uniform(-0.5, 0.5, r.nj) * r.joint_torque_max
        elif o.kind == "centroidal_vel":
            u[:lead] = np.concatenate((rng.uniform(-1, 1, 6), rng.uniform(-0.25, 0.25, r.nj) * r.joint_vel_max))[-lead:]
        elif lead:
            u[:lead] = rng.normal(0, 5, lead)
        for kf in range(4):
            fz = rng.uniform(0, mg)
            u[lead + 3 * kf:lead + 3 * kf + 3] = np.array([0.7 * fz * rng.uniform(-0.5, 0.5), 0.7 * fz * rng.uniform(-0.5, 0.5), fz]) * contact[i, kf]
        if r.nf > 12:
            u[lead + 12:lead + 15] = rng.uniform(-20, 20, 3)
        if o.kind == "whole_body_rnea" and i < o.tau_nodes:
            u[o.tau_idx:] = rng.uniform(-0.5, 0.5, r.nj) * r.joint_torque_max
    return x, p
