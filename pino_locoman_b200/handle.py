"""Thin object wrapper over the C-ABI handle; tensors are handed over as raw device pointers.

PyTorch is used for device memory and streams only.  Every method takes/returns ``torch.float64`` CUDA tensors with a
leading batch dimension; the arithmetic happens in libpinolocoman_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .utils.robot import robot_desc


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_in(t, shape_tail, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise TypeError(f"{name}: expected a contiguous torch.float64 CUDA tensor")
    if tuple(t.shape[1:]) != tuple(shape_tail):
        raise ValueError(f"{name}: expected shape [batch, {', '.join(map(str, shape_tail))}], got {tuple(t.shape)}")
    return t


class Handle:
    """One problem formulation (robot x dynamics x horizon) on one GPU."""

    def __init__(self, robot, dynamics, nodes, max_batch, tau_nodes=3, device=None, include_base=True, include_acc=True, **osqp_opts):
        if dynamics not in _lib.DYNAMICS_ID:
            raise ValueError(f"Unknown dynamics type: {dynamics}")
        self.layout_only = int(max_batch) == 0     # layout queries only (host-logic tests); no compute
        if not self.layout_only and not torch.cuda.is_available():
            raise RuntimeError("pino_locoman_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        if self.layout_only:
            self.device = torch.device("cpu")
        else:
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            torch.cuda.set_device(self.device)
        self._rd = robot_desc(robot)
        od = _lib.OcpDesc()
        self.lib.plm_fill_default_ocp_desc(ctypes.byref(od), _lib.DYNAMICS_ID[dynamics], nodes)
        od.tau_nodes = tau_nodes
        od.include_base = int(bool(include_base))
        od.include_acc = int(bool(include_acc))
        for k, v in osqp_opts.items():
            setattr(od, "osqp_" + k, v)
        self.ocp_desc = od
        self.max_batch = int(max_batch)
        h = ctypes.c_void_p()
        rc = self.lib.plm_create(ctypes.byref(self._rd), ctypes.byref(od), self.max_batch, ctypes.byref(h))
        self._h = h
        if rc != 0:
            msg = self.lib.plm_last_error(h).decode() if h else "allocation failed"
            if h:
                self.lib.plm_destroy(h)
            self._h = None
            raise (ValueError if rc == 2 else _lib.PlmError)(msg)
        d = _lib.Dims()
        self.lib.plm_get_dims(h, ctypes.byref(d))
        self.dims = d
        self.n, self.m, self.np, self.nnz, self.nodes = d.n, d.m, d.np, d.nnz, d.nodes
        self.ndx, self.nx, self.nq, self.nv, self.nj, self.nf = d.ndx, d.nx, d.nq, d.nv, d.nj, d.nf
        x_off = (ctypes.c_int32 * (nodes + 2))()
        nu = (ctypes.c_int32 * nodes)()
        row_off = (ctypes.c_int32 * (nodes + 2))()
        self.lib.plm_stage_offsets(h, x_off, nu, row_off)
        self.x_off, self.nu, self.row_off = list(x_off), list(nu), list(row_off)
        po = (ctypes.c_int32 * 16)()
        self.lib.plm_param_offsets(h, po)
        names = ["x_init", "dt_min", "dt_max", "contact_schedule", "swing_schedule", "n_contacts", "swing_period",
                 "swing_height", "swing_vel_limits", "Q_diag", "R_diag", "base_vel_des", "ext_force_des", "arm_vel_des",
                 "tau_prev", "W_diag"]
        self.p_off = {k: int(v) for k, v in zip(names, po)}
        rows = np.zeros(self.nnz, dtype=np.int32)
        cols = np.zeros(self.nnz, dtype=np.int32)
        self.lib.plm_jac_pattern(h, rows.ctypes.data_as(_lib.c_int32_p), cols.ctypes.data_as(_lib.c_int32_p))
        self.jac_rows, self.jac_cols = rows, cols

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib.plm_destroy(self._h)
            self._h = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _rc(self, rc):
        if rc != 0:
            raise _lib.PlmError(self.lib.plm_last_error(self._h).decode())

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def launch_count(self):
        return int(self.lib.plm_launch_count(self._h))

    # ------------------------------------------------------------------ casadi Function equivalents
    def sqp_data(self, x, p, bounds=True):
        """sqp_data(x, p) -> grad_f [B,n], J values [B,nnz], g [B,m], lbg, ubg   (optimization/ocp.py:287)."""
        B = _check_in(x, (self.n,), "x").shape[0]
        _check_in(p, (self.np,), "p")
        grad, J, g = self._new(B, self.n), self._new(B, self.nnz), self._new(B, self.m)
        lbg = self._new(B, self.m) if bounds else None
        ubg = self._new(B, self.m) if bounds else None
        self._rc(self.lib.plm_sqp_data(self._h, _ptr(x), _ptr(p), B, _ptr(grad), _ptr(J), _ptr(g), _ptr(lbg), _ptr(ubg), self._stream()))
        return grad, J, g, lbg, ubg

    def g_data(self, x, p, bounds=True):
        """g_data(x, p) -> g, lbg, ubg   (optimization/ocp.py:290)."""
        B = _check_in(x, (self.n,), "x").shape[0]
        _check_in(p, (self.np,), "p")
        g = self._new(B, self.m)
        lbg = self._new(B, self.m) if bounds else None
        ubg = self._new(B, self.m) if bounds else None
        self._rc(self.lib.plm_g_data(self._h, _ptr(x), _ptr(p), B, _ptr(g), _ptr(lbg), _ptr(ubg), self._stream()))
        return g, lbg, ubg

    def f_data(self, x, p):
        """f_data(x, p) -> f [B], grad_f [B,n]   (optimization/ocp.py:289)."""
        B = _check_in(x, (self.n,), "x").shape[0]
        _check_in(p, (self.np,), "p")
        f, grad = self._new(B), self._new(B, self.n)
        self._rc(self.lib.plm_f_data(self._h, _ptr(x), _ptr(p), B, _ptr(f), _ptr(grad), self._stream()))
        return f, grad

    def hess_diag(self, p):
        """diag(hess_data(x, p)) [B,n]   (optimization/ocp.py:288,293-296)."""
        B = _check_in(p, (self.np,), "p").shape[0]
        h = self._new(B, self.n)
        self._rc(self.lib.plm_hess_diag(self._h, _ptr(p), B, _ptr(h), self._stream()))
        return h

    def jac_dense(self, Jvals):
        """Scatter J values [B,nnz] into dense [B,m,n] (tests / small problems only)."""
        B = Jvals.shape[0]
        D = torch.zeros(B, self.m, self.n, dtype=torch.float64, device=Jvals.device)
        r = torch.as_tensor(self.jac_rows, dtype=torch.long, device=Jvals.device)
        c = torch.as_tensor(self.jac_cols, dtype=torch.long, device=Jvals.device)
        D[:, r, c] = Jvals
        return D

    # ------------------------------------------------------------------ OSQP equivalents
    def qp_setup(self, hess):
        """osqp setup(): zero iterates, record the setup-time row scaling (optimization/ocp.py:305-313)."""
        B = _check_in(hess, (self.n,), "hess").shape[0]
        self._rc(self.lib.plm_qp_setup(self._h, B, _ptr(hess), self._stream()))

    def qp_get_scaling(self, batch):
        D, E, c = self._new(batch, self.n), self._new(batch, self.m), self._new(batch)
        self._rc(self.lib.plm_qp_get_scaling(self._h, batch, _ptr(D), _ptr(E), _ptr(c), self._stream()))
        return D, E, c

    def qp_update(self, hess, q, J, l, u):
        B = _check_in(q, (self.n,), "q").shape[0]
        _check_in(hess, (self.n,), "hess")
        _check_in(J, (self.nnz,), "J")
        _check_in(l, (self.m,), "l")
        _check_in(u, (self.m,), "u")
        self._rc(self.lib.plm_qp_update(self._h, B, _ptr(hess), _ptr(q), _ptr(J), _ptr(l), _ptr(u), self._stream()))

    def qp_solve(self, batch):
        dx = self._new(batch, self.n)
        iters = self._new(batch, dtype=torch.int32)
        status = self._new(batch, dtype=torch.int32)
        self._rc(self.lib.plm_qp_solve(self._h, batch, _ptr(dx), _ptr(iters), _ptr(status), self._stream()))
        return dx, iters, status

    def qp_get_iterates(self, batch):
        x, z, y = self._new(batch, self.n), self._new(batch, self.m), self._new(batch, self.m)
        self._rc(self.lib.plm_qp_get_iterates(self._h, batch, _ptr(x), _ptr(z), _ptr(y), self._stream()))
        return x, z, y

    def qp_set_iterates(self, x, z, y):
        B = x.shape[0]
        self._rc(self.lib.plm_qp_set_iterates(self._h, B, _ptr(x), _ptr(z), _ptr(y), self._stream()))

    # ------------------------------------------------------------------ line search / SQP step
    def line_search(self, x, p, dx):
        B = _check_in(x, (self.n,), "x").shape[0]
        _check_in(dx, (self.n,), "dx")
        x_new, info = self._new(B, self.n), self._new(B, 4)
        self._rc(self.lib.plm_line_search(self._h, _ptr(x), _ptr(p), _ptr(dx), B, _ptr(x_new), _ptr(info), self._stream()))
        return x_new, info

    def sqp_step(self, x, p, x_new=None, stats=None):
        B = _check_in(x, (self.n,), "x").shape[0]
        _check_in(p, (self.np,), "p")
        x_new = self._new(B, self.n) if x_new is None else x_new
        stats = self._new(B, 8) if stats is None else stats
        self._rc(self.lib.plm_sqp_step(self._h, _ptr(x), _ptr(p), B, _ptr(x_new), _ptr(stats), self._stream()))
        return x_new, stats

    def last_phase_ms(self):
        ms = (ctypes.c_double * 4)()
        self._rc(self.lib.plm_last_phase_ms(self._h, ms))
        return list(ms)
