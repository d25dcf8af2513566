"""centroidal_vel: state [h_com/m, q], U_i = (v, f); dh_next == dh + hdot/m dt, path constraint A v - m h == 0
(reference optimization/ocp_centroidal_vel.py)."""
import numpy as np

from . import _weights as W
from .ocp import OCP, integrate_configuration


class OCPCentroidalVel(OCP):
    dynamics = "centroidal_vel"

    def __init__(self, robot, solver, nodes, include_base=False, batch=1, device=None):
        super().__init__(robot, solver, nodes, batch=batch, device=device)
        if not include_base:
            raise NotImplementedError("include_base=False (base velocity from the dynamics) is not available yet")
        self.include_base = include_base
        self.nv_opt = self.nv
        self.x_nom = np.concatenate((np.zeros(6), robot.q0))
        self.f_idx = self.nv_opt
        self.h_sol = []

    def set_weights(self):   # ocp_centroidal_vel.py:25-49
        Q = np.concatenate(([1000.0] * 6, W.q_base_pos(), W.q_joint_pos(bool(self.arm_ee_frame))))
        R = np.concatenate(([1.0] * self.nv_opt, [1e-3] * self.nf))
        self._set("Q_diag", Q)
        self._set("R_diag", R)

    def state_integrate(self, x, dx):
        return np.concatenate([x[..., :6] + dx[..., :6], integrate_configuration(x[..., 6:], dx[..., 6:])], -1)

    def _append_state(self, x_sol):
        self.h_sol.append(x_sol[:, :6])
        self.q_sol.append(x_sol[:, 6:])

    def _append_solution(self, x_sol, u_sol):
        self._append_state(x_sol)
        self.v_sol.append(u_sol[:, :self.nv_opt].copy())
        self.forces_sol.append(u_sol[:, self.f_idx:].copy())
