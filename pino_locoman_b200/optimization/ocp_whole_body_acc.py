"""whole_body_acc: U_i = (a, f); path constraint rnea(q, v, a, f)[:6] == 0 (reference optimization/ocp_whole_body_acc.py)."""
import numpy as np

from . import _weights as W
from .ocp import OCP


class OCPWholeBodyAcc(OCP):
    dynamics = "whole_body_acc"

    def __init__(self, robot, solver, nodes, include_base=False, batch=1, device=None):
        super().__init__(robot, solver, nodes, batch=batch, device=device)
        # include_base=False: inputs (a_j, f); a_b = base_acc_dynamics(q, v, a_j, f) is substituted inside the node kernel and
        # the six gap rows are dropped (ocp_whole_body_acc.py:20-24,109-141)
        self.include_base = bool(include_base)
        self.na_opt = self.nv if self.include_base else self.nj
        self.x_nom = np.concatenate((robot.q0, np.zeros(self.nv)))
        self.f_idx = self.na_opt

    def set_weights(self):   # ocp_whole_body_acc.py:25-53
        Q = np.concatenate((W.q_base_pos(), W.q_joint_pos(bool(self.arm_ee_frame)), W.q_vel(self.nj)))
        R = np.concatenate(([1e-3] * self.na_opt, [1e-3] * self.nf))
        self._set("Q_diag", Q)
        self._set("R_diag", R)
