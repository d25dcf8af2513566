"""whole_body_aba: U_i = (tau_j, f); dv_next == dv + aba(q, v, tau_j, f) dt (reference optimization/ocp_whole_body_aba.py)."""
import numpy as np

from . import _weights as W
from .ocp import OCP


class OCPWholeBodyABA(OCP):
    dynamics = "whole_body_aba"

    def __init__(self, robot, solver, nodes, batch=1, device=None):
        super().__init__(robot, solver, nodes, batch=batch, device=device)
        self.x_nom = np.concatenate((robot.q0, np.zeros(self.nv)))
        self.tau_sol = []
        self.f_idx = self.nj

    def set_weights(self):   # ocp_whole_body_aba.py:22-50
        Q = np.concatenate((W.q_base_pos(), W.q_joint_pos(bool(self.arm_ee_frame)), W.q_vel(self.nj)))
        R = np.concatenate(([1e-3] * self.nj, [1e-3] * self.nf))
        self._set("Q_diag", Q)
        self._set("R_diag", R)

    def _append_solution(self, x_sol, u_sol):
        self._append_state(x_sol)
        self.tau_sol.append(u_sol[:, :self.nj].copy())
        self.forces_sol.append(u_sol[:, self.f_idx:].copy())
