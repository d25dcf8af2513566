"""Compiled CPU port of one SQP iteration of the reference (TEST / BASELINE INFRASTRUCTURE, see oracle/__init__.py).

``plm_cport.cpp`` restates optimization/ocp.py:383-406 (sqp_data -> OSQP update / solve -> Armijo) in C++; this module
builds it (``make``), describes the robot to it from the ORACLE's model (oracle/model.py: the product package is not
imported) and wraps one instance.  bench.py times it as the CPU baseline ("port (C++)", one process per core);
tests/test_cport.py pins it to the numpy oracle (same ADMM iterate sequence, SQP step <= 1e-6).
"""
import ctypes
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libplm_cport.so")
DYN_ID = {"centroidal_vel": 0, "centroidal_acc": 1, "whole_body_acc": 2, "whole_body_aba": 3, "whole_body_rnea": 4}
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


class RobotDesc(ctypes.Structure):      # plm_robot_desc (include/pino_locoman_b200.h)
    _fields_ = [("nbody", ctypes.c_int32), ("parent", _ip), ("placement", _dp), ("axis", _dp), ("inertia", _dp),
                ("nfeet", ctypes.c_int32), ("has_ext_force", ctypes.c_int32), ("contact_body", _ip), ("contact_offset", _dp),
                ("arm_body", ctypes.c_int32), ("arm_offset", ctypes.c_double * 3), ("joint_pos_min", _dp), ("joint_pos_max", _dp),
                ("joint_vel_max", _dp), ("joint_torque_max", _dp), ("q0", _dp)]


class OcpDesc(ctypes.Structure):        # plm_ocp_desc
    _fields_ = [("dynamics", ctypes.c_int32), ("nodes", ctypes.c_int32), ("tau_nodes", ctypes.c_int32),
                ("mu", ctypes.c_double), ("osqp_max_iter", ctypes.c_int32), ("osqp_check_termination", ctypes.c_int32),
                ("osqp_scaling", ctypes.c_int32), ("osqp_rho", ctypes.c_double), ("osqp_sigma", ctypes.c_double),
                ("osqp_alpha", ctypes.c_double), ("osqp_eps_abs", ctypes.c_double), ("osqp_eps_rel", ctypes.c_double),
                ("osqp_eps_prim_inf", ctypes.c_double), ("osqp_eps_dual_inf", ctypes.c_double),
                ("include_base", ctypes.c_int32), ("include_acc", ctypes.c_int32)]


def build(portable=False):
    """Compile libplm_cport.so if it is missing or older than its sources."""
    subprocess.check_call(["make", "-C", HERE, "-s"] + (["portable"] if portable else []))
    return SO


_LIB = None


def load():
    """The compiled library; raises if it has not been built (``build()`` / ``__graft_entry__.build()``)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            raise FileNotFoundError(SO)
        lib = ctypes.CDLL(SO)
        lib.cport_create.restype = ctypes.c_void_p
        lib.cport_eval_seconds.restype = ctypes.c_double
        _LIB = lib
    return _LIB


def robot_desc_from_oracle(robot):
    """plm_robot_desc from an oracle.model.OracleRobot (body 0 = free-flyer root, bodies 1.. = revolute joints)."""
    m = robot.model
    nb = m.njoints - 1                      # drop the universe
    keep = {}

    def dptr(name, a):
        keep[name] = np.ascontiguousarray(a, dtype=np.float64)
        return keep[name].ctypes.data_as(_dp)

    def iptr(name, a):
        keep[name] = np.ascontiguousarray(a, dtype=np.int32)
        return keep[name].ctypes.data_as(_ip)

    parent = [-1] + [m.parents[j] - 1 for j in range(2, m.njoints)]
    placement = np.zeros((nb, 12))
    axis = np.zeros((nb, 3))
    inertia = np.zeros((nb, 10))
    for b in range(nb):
        j = b + 1
        placement[b, :9] = np.asarray(m.placement_R[j]).reshape(9)
        placement[b, 9:] = m.placement_p[j]
        if b > 0:
            axis[b] = m.axis[j]
        Ic = m.Ic[j]
        inertia[b] = [m.mass[j], *m.com[j], Ic[0, 0], Ic[0, 1], Ic[0, 2], Ic[1, 1], Ic[1, 2], Ic[2, 2]]
    frames = list(robot.foot_frames) + ([robot.ext_force_frame] if robot.ext_force_frame else [])
    d = RobotDesc()
    d.nbody = nb
    d.parent = iptr("parent", parent)
    d.placement, d.axis, d.inertia = dptr("placement", placement), dptr("axis", axis), dptr("inertia", inertia)
    d.nfeet, d.has_ext_force = 4, int(bool(robot.ext_force_frame))
    d.contact_body = iptr("contact_body", [m.frames[f].parent - 1 for f in frames])
    d.contact_offset = dptr("contact_offset", np.array([m.frames[f].p for f in frames]))
    if robot.arm_ee_frame:
        fr = m.frames[robot.arm_ee_frame]
        d.arm_body = fr.parent - 1
        d.arm_offset = (ctypes.c_double * 3)(*fr.p)
    else:
        d.arm_body = -1
        d.arm_offset = (ctypes.c_double * 3)(0.0, 0.0, 0.0)
    d.joint_pos_min, d.joint_pos_max = dptr("pmin", robot.joint_pos_min), dptr("pmax", robot.joint_pos_max)
    d.joint_vel_max, d.joint_torque_max = dptr("vmax", robot.joint_vel_max), dptr("tmax", robot.joint_torque_max)
    d.q0 = dptr("q0", robot.q0)
    d._keep = keep
    return d


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CPortSQP:
    """One MPC instance: mirror of oracle.sqp.OracleSQP on the compiled port."""

    def __init__(self, oocp, **osqp_opts):
        self.o = oocp
        self.lib = load()
        self._rd = robot_desc_from_oracle(oocp.robot)
        od = OcpDesc(DYN_ID[oocp.kind], oocp.nodes, max(oocp.tau_nodes, 1), 0.7, 100, 25, 10, 2e-2, 1e-6, 1.4, 1e-3, 1e-3, 1e-4, 1e-4,
                     int(getattr(oocp, "include_base", True)), int(getattr(oocp, "include_acc", True)))
        for k, v in osqp_opts.items():
            setattr(od, "osqp_" + k, v)
        err = ctypes.create_string_buffer(256)
        h = self.lib.cport_create(ctypes.byref(self._rd), ctypes.byref(od), err, 256)
        if not h:
            raise RuntimeError(err.value.decode())
        self.h = ctypes.c_void_p(h)
        d = (ctypes.c_int * 6)()
        self.lib.cport_dims(self.h, d)
        self.n, self.m, self.np_, self.nnz, self.ndx, self.nodes = list(d)
        assert (self.n, self.m, self.np_) == (oocp.n, oocp.m, oocp.np_), "layout mismatch between the port and the oracle"

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.cport_destroy(self.h)
            self.h = None

    def objective_terms(self, p):
        """hess_diag and the two weighted-quadratic terms of f (x-independent; oracle/ocp.py f_data / hess_diag)."""
        o = self.o
        P = o.unpack_p(p)
        w1, t1 = o._weights(P), o._targets_stacked(P)
        w2, t2 = np.zeros(o.n), np.zeros(o.n)
        if o.kind == "whole_body_rnea" and o.tau_nodes > 0:
            off = o.x_off[0] + o.ndx + o.tau_idx
            w2[off:off + o.nj] = P["W_diag"]
            t2[off:off + o.nj] = P["tau_prev"]
        return 2 * (w1 + w2), w1, t1, w2, t2

    def init_solver(self, p=None):
        p = self.o.p_vector() if p is None else p
        hess = self.objective_terms(p)[0]
        self.lib.cport_init_solver(self.h, _p(np.ascontiguousarray(hess)))

    def prepare(self, x, p):
        """Everything of one iteration that does not depend on x (bounds, objective terms)."""
        _, lbg, ubg = self.o.g_data(x, p)
        return tuple(np.ascontiguousarray(a, dtype=np.float64) for a in (*self.objective_terms(p), lbg, ubg))

    def solve(self, x, p, prepared=None):
        hess, w1, t1, w2, t2, lbg, ubg = self.prepare(x, p) if prepared is None else prepared
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        x_new, dx, info = np.zeros(self.n), np.zeros(self.n), np.zeros(8)
        self.lib.cport_sqp_iteration(self.h, _p(x), _p(p), _p(hess), _p(w1), _p(t1), _p(w2), _p(t2), _p(lbg), _p(ubg),
                                     _p(x_new), _p(dx), _p(info))
        status = {1: "solved", 2: "solved inaccurate", -2: "maximum iterations reached", 3: "primal infeasible inaccurate",
                  -3: "primal infeasible", 4: "dual infeasible inaccurate", -4: "dual infeasible"}.get(int(info[1]), str(int(info[1])))
        return x_new, dict(sol_dx=dx, qp_iters=int(info[0]), qp_status=status, accepted=bool(info[2]), alpha=info[3],
                           trials=int(info[4]), f=info[5], g_metric=info[6], violation_max=info[7])

    def node_eval(self, x, p, want_jac=True):
        g, J = np.zeros(self.m), np.zeros(self.nnz)
        self.lib.cport_node_eval(self.h, _p(np.ascontiguousarray(x)), _p(np.ascontiguousarray(p)), _p(g), _p(J), int(want_jac))
        return g, J

    def eval_seconds(self, reset=True):
        return float(self.lib.cport_eval_seconds(self.h, int(reset)))

    def scaling(self):
        D, E, c = np.zeros(self.n), np.zeros(self.m), ctypes.c_double()
        self.lib.cport_get_scaling(self.h, _p(D), _p(E), ctypes.byref(c))
        return D, E, c.value

    def iterates(self):
        x, z, y = np.zeros(self.n), np.zeros(self.m), np.zeros(self.m)
        self.lib.cport_get_iterates(self.h, _p(x), _p(z), _p(y))
        return x, z, y


def _worker(args):
    robot, dynamics, nodes, xs, ps = args
    from oracle.model import OracleRobot
    from oracle.ocp import OracleOCP
    o = OracleOCP(OracleRobot(robot), dynamics, nodes)
    t_sqp = t_eval = 0.0
    for x, p in zip(xs, ps):
        for name, (off, sz) in o.p_layout.items():
            o.params[name][:] = p[off:off + sz]
        s = CPortSQP(o)
        s.init_solver(p)
        prep = s.prepare(x, p)
        s.eval_seconds(reset=True)
        t0 = time.perf_counter()
        s.solve(x, p, prep)
        t_sqp += time.perf_counter() - t0
        t_eval += s.eval_seconds()
    return t_sqp, t_eval


def time_sqp(robot, dynamics, nodes, x, p, procs):
    """One SQP iteration per instance on `procs` processes (one per core).  Returns (wall seconds of the slowest
    process for its share = the time the whole sample takes with all cores busy, node-evaluation seconds summed)."""
    n_inst = len(x)
    shares = [list(range(k, n_inst, procs)) for k in range(procs)]
    jobs = [(robot, dynamics, nodes, [x[i] for i in sh], [p[i] for i in sh]) for sh in shares if sh]
    if len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(len(jobs)) as pool:
            res = pool.map(_worker, jobs)
    else:
        res = [_worker(j) for j in jobs]
    return max(r[0] for r in res), sum(r[1] for r in res)
