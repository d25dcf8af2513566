"""Problem files for the CasADi-external shim ``libplm_casadi.so``.

    from pino_locoman_b200.casadi_shim import export_problem, SHIM_PATH
    export_problem(B2(), "whole_body_rnea", nodes=14, path="b2_rnea_N14.plm", tau_nodes=3)
    # in the reference (optimization/ocp.py:299-301), with PLM_CASADI_PROBLEM=b2_rnea_N14.plm in the environment:
    #   self.sqp_data = ca.external("sqp_data", SHIM_PATH)
    #   self.f_data   = ca.external("f_data", SHIM_PATH)
    #   self.g_data   = ca.external("g_data", SHIM_PATH)
"""
import os
import struct

import numpy as np

from . import _lib

SHIM_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libplm_casadi.so")
_MAGIC, _VERSION = 0x504D4C50, 1


def export_problem(robot, dynamics, nodes, path, tau_nodes=None, mu=0.7):
    """Write the robot tables and the formulation of one OCP to ``path`` (read by the shim at its first call)."""
    if dynamics not in _lib.DYNAMICS_ID:
        raise ValueError(f"Dynamics {dynamics} not supported")
    tab = robot.tables()
    if tau_nodes is None:
        tau_nodes = 3 if dynamics == "whole_body_rnea" else 0
    q0 = np.ascontiguousarray(tab["q0"], dtype=np.float64)
    with open(path, "wb") as f:
        f.write(struct.pack("<10i", _MAGIC, _VERSION, _lib.DYNAMICS_ID[dynamics], int(nodes), max(int(tau_nodes), 1), int(tab["nbody"]),
                            int(tab["nfeet"]), int(tab["has_ext_force"]), int(tab["arm_body"]), int(q0.size)))
        f.write(struct.pack("<d", float(mu)))
        f.write(struct.pack("<3d", *[float(v) for v in tab["arm_offset"]]))
        for name in ("parent", "contact_body"):
            f.write(np.ascontiguousarray(tab[name], dtype=np.int32).tobytes())
        for name in ("placement", "axis", "inertia", "contact_offset", "joint_pos_min", "joint_pos_max", "joint_vel_max", "joint_torque_max"):
            f.write(np.ascontiguousarray(tab[name], dtype=np.float64).tobytes())
        f.write(q0.tobytes())
    return path
