"""whole_body_rnea: U_i = (a, f, tau_j for i < tau_nodes); rows rnea[:6] == 0, rnea[6:] == tau_j, torque bounds
(reference optimization/ocp_whole_body_rnea.py)."""
import numpy as np

from . import _weights as W
from .ocp import OCP


class OCPWholeBodyRNEA(OCP):
    dynamics = "whole_body_rnea"

    def __init__(self, robot, solver, nodes, tau_nodes, include_acc=True, batch=1, device=None):
        super().__init__(robot, solver, nodes, batch=batch, device=device)
        if not 1 <= tau_nodes <= nodes:
            raise ValueError("tau_nodes must be in [1, nodes]")
        self.tau_nodes = tau_nodes
        # include_acc=False: no acceleration inputs, a_i = (v_{i+1} - v_i) / dt_i inside the node kernel and no dv
        # integrator rows (ocp_whole_body_rnea.py:21-25,156,183-191); the RNEA rows then couple stage i with dv_{i+1}
        self.include_acc = bool(include_acc)
        self.na_opt = self.nv if self.include_acc else 0
        self.x_nom = np.concatenate((robot.q0, np.zeros(self.nv)))
        self.tau_sol = []
        self.f_idx = self.na_opt
        self.tau_idx = self.f_idx + self.nf

    def set_weights(self):   # ocp_whole_body_rnea.py:28-63
        Q = np.concatenate((W.q_base_pos(), W.q_joint_pos(bool(self.arm_ee_frame)), W.q_vel(self.nj)))
        R = np.concatenate(([1e-3] * self.na_opt, [1e-3] * self.nf, [1e-4] * self.nj))
        self._set("Q_diag", Q)
        self._set("R_diag", R)
        self._set("W_diag", np.zeros(self.nj))

    def update_previous_torques(self, tau_prev):
        self._set("tau_prev", tau_prev)

    def get_tau_sol(self, i):
        return self.U_prev[i][:, self.tau_idx:]

    def _append_solution(self, x_sol, u_sol):
        super()._append_solution(x_sol, u_sol)
        self.tau_sol.append(u_sol[:, self.tau_idx:].copy())
