"""centroidal_vel: state [h_com/m, q], U_i = (v, f); dh_next == dh + hdot/m dt, path constraint A v - m h == 0
(reference optimization/ocp_centroidal_vel.py)."""
import numpy as np

from . import _weights as W
from .ocp import OCP, integrate_configuration


class OCPCentroidalVel(OCP):
    dynamics = "centroidal_vel"

    def __init__(self, robot, solver, nodes, include_base=False, batch=1, device=None):
        super().__init__(robot, solver, nodes, batch=batch, device=device)
        # include_base=False: inputs (v_j, f); v_b = base_vel_dynamics(h, q, v_j) is substituted inside the node kernel and
        # the six gap rows are dropped (ocp_centroidal_vel.py:19-23,104-120)
        self.include_base = bool(include_base)
        self.nv_opt = self.nv if self.include_base else self.nj
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
        if self.include_base:
            self.v_sol.append(u_sol[:, :self.nv_opt].copy())
        else:   # v = (base_vel(h, q, v_j), v_j)   (ocp_centroidal_vel.py:224-231)
            import torch
            if getattr(self, "_dyn_post", None) is None:
                from ..dynamics import DynamicsCentroidalVel
                self._dyn_post = DynamicsCentroidalVel(self.model, self.mass, self.foot_frames, max_batch=self.batch, device=self.handle.device)
            dev = self.handle.device
            v_j = np.ascontiguousarray(u_sol[:, :self.nv_opt])
            v_b = self._dyn_post.base_vel_dynamics()(torch.from_numpy(np.ascontiguousarray(x_sol[:, :6])).to(dev),
                                                     torch.from_numpy(np.ascontiguousarray(x_sol[:, 6:])).to(dev), torch.from_numpy(v_j).to(dev))
            self.v_sol.append(np.concatenate((v_b.cpu().numpy(), v_j), 1))
        self.forces_sol.append(u_sol[:, self.f_idx:].copy())
