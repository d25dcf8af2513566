"""whole_body_acc: U_i = (a, f); path constraint rnea(q, v, a, f)[:6] == 0 (reference optimization/ocp_whole_body_acc.py)."""
import numpy as np

from . import _weights as W
from .ocp import OCP


class OCPWholeBodyAcc(OCP):
    dynamics = "whole_body_acc"

    def __init__(self, robot, solver, nodes, include_base=False, batch=1, device=None):
        super().__init__(robot, solver, nodes, batch=batch, device=device)
        if not include_base:
            raise NotImplementedError("include_base=False (base acceleration from the dynamics) is not available yet")
        self.include_base = include_base
        self.na_opt = self.nv
        self.x_nom = np.concatenate((robot.q0, np.zeros(self.nv)))
        self.f_idx = self.na_opt

    def set_weights(self):   # ocp_whole_body_acc.py:25-53
        Q = np.concatenate((W.q_base_pos(), W.q_joint_pos(bool(self.arm_ee_frame)), W.q_vel(self.nj)))
        R = np.concatenate(([1e-3] * self.na_opt, [1e-3] * self.nf))
        self._set("Q_diag", Q)
        self._set("R_diag", R)
