"""State / input weight diagonals shared by the formulations (set_weights of each optimization/ocp_*.py)."""
import numpy as np


def q_base_pos():
    return np.array([0, 0, 1000, 10000, 10000, 0], dtype=float)    # base x/y, z, rot x/y, rot z


def q_joint_pos(has_arm):
    q = np.tile([1000.0, 500.0, 500.0], 4)                          # hip, thigh, calf
    return np.concatenate((q, [100.0] * 6)) if has_arm else q


def q_vel(nj):
    return np.concatenate(([2000, 2000, 1000, 1000, 1000, 2000], [1.0] * nj)).astype(float)
