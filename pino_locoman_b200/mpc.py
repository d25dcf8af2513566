"""Batched receding-horizon driver, device resident.

``BatchedMPC`` runs the generic loop of the reference's ``run_mpc.py:115-143`` for every instance of an OCP at once and
keeps the decision vector, the parameter vector and the initial states on the GPU between steps: per step one launch
sequence ``plm_mpc_step`` = update_gait_sequence(k dt_min) -> warm_start() -> solve() -> x_init = integrate(x_init,
DX_prev[1]).  As in the reference's generic branch the previous torques ``tau_prev`` stay at their initial value.
With ``warm_start=False`` every step starts from ``opti.initial()`` (DX = 0, U = u_des), as ``run_mpc.py`` does when
its ``warm_start`` flag is off.

    ocp = make_ocp(..., batch=B, device="cuda:0"); ocp.set_time_params(...); ...; ocp.update_initial_state(x_init)
    mpc = BatchedMPC(ocp, warm_start=True)
    for k in range(mpc_loops):
        stats = mpc.step()            # torch [B, 8] on the device (see plm_sqp_step)
    x_init = mpc.x_init()             # torch [B, nx]
"""
import ctypes

import numpy as np
import torch

from .handle import _ptr

_GAIT_ID = {"trot": 0, "walk": 1, "stand": 2}


class BatchedMPC:
    def __init__(self, ocp, warm_start=True, t0=None, update_tau_prev=False):
        if ocp.solver != "osqp":
            raise ValueError(f"Solver {ocp.solver} not supported")
        if getattr(ocp, "hess_diag", None) is None:
            ocp.init_solver()
        self.ocp, self.h = ocp, ocp.handle
        self.warm_start = bool(warm_start)
        self.update_tau_prev = bool(update_tau_prev)     # run_mpc.py:108-111 (compiled-solver branch of the loop)
        self.k = 0
        dev = self.h.device
        gs = ocp.gait_sequence
        self._gait, self._period = _GAIT_ID[gs.gait_type], float(gs.gait_period)
        self._dts = (ctypes.c_double * ocp.nodes)(*[float(d) for d in ocp.dts])
        self._dt_min = float(ocp._get("dt_min")[0, 0])
        self._t0 = None if t0 is None else torch.as_tensor(np.broadcast_to(np.asarray(t0, dtype=np.float64), (ocp.batch,)).copy(), device=dev)
        self.x = torch.from_numpy(ocp.initial_guess() if ocp._x0 is None else np.ascontiguousarray(ocp._x0)).to(dev)
        self.x_new = torch.empty_like(self.x)
        self.p = ocp._p_device().clone()
        self.stats = torch.empty(ocp.batch, 8, dtype=torch.float64, device=dev)

    def step(self):
        """One MPC step for every instance; returns the device tensor of SQP statistics [B, 8]."""
        h = self.h
        rc = h.lib.plm_mpc_step(h._h, _ptr(self.x), _ptr(self.p), _ptr(self._t0), self.k * self._dt_min, self._gait, self._period,
                                self._dts, 2 if self.k == 0 else int(self.warm_start), int(self.update_tau_prev), self.ocp.batch,
                                _ptr(self.x_new), _ptr(self.stats), h._stream())
        h._rc(rc)
        self.x, self.x_new = self.x_new, self.x
        self.k += 1
        return self.stats

    def x_init(self):
        o = self.h.p_off["x_init"]
        return self.p[:, o:o + self.ocp.nx]

    def solution(self):
        """Stacked decision vector of the last step [B, n] (device)."""
        return self.x
