"""ctypes binding of libpinolocoman_b200.so (C ABI in include/pino_locoman_b200.h).

The library is CUDA-only: importing works without a GPU (so layouts can be inspected and the symbol table
checked), but creating a handle raises unless a CUDA device is present.  There is no CPU fallback.
"""
import ctypes
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libpinolocoman_b200.so")

DYNAMICS_ID = {"centroidal_vel": 0, "centroidal_acc": 1, "whole_body_acc": 2, "whole_body_aba": 3, "whole_body_rnea": 4}

c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_double_p = ctypes.POINTER(ctypes.c_double)
vp = ctypes.c_void_p


class OcpDesc(ctypes.Structure):
    """ctypes image of ``plm_ocp_desc``."""
    _fields_ = [("dynamics", ctypes.c_int32), ("nodes", ctypes.c_int32), ("tau_nodes", ctypes.c_int32),
                ("mu", ctypes.c_double), ("osqp_max_iter", ctypes.c_int32), ("osqp_check_termination", ctypes.c_int32),
                ("osqp_scaling", ctypes.c_int32), ("osqp_rho", ctypes.c_double), ("osqp_sigma", ctypes.c_double),
                ("osqp_alpha", ctypes.c_double), ("osqp_eps_abs", ctypes.c_double), ("osqp_eps_rel", ctypes.c_double),
                ("osqp_eps_prim_inf", ctypes.c_double), ("osqp_eps_dual_inf", ctypes.c_double),
                ("include_base", ctypes.c_int32), ("include_acc", ctypes.c_int32)]


class Dims(ctypes.Structure):
    """ctypes image of ``plm_dims``."""
    _fields_ = [(k, ctypes.c_int32) for k in ("nq", "nv", "nj", "nf", "nx", "ndx", "n", "m", "np", "nnz", "nodes",
                                               "kkt_factor_doubles")]


# name -> (restype, argtypes); every symbol declared in include/pino_locoman_b200.h
SIGNATURES = {
    "plm_fill_default_ocp_desc": (None, [ctypes.POINTER(OcpDesc), ctypes.c_int32, ctypes.c_int32]),
    "plm_abi_struct_sizes": (None, [c_int32_p]),
    "plm_create": (ctypes.c_int, [vp, ctypes.POINTER(OcpDesc), ctypes.c_int32, ctypes.POINTER(vp)]),
    "plm_destroy": (None, [vp]),
    "plm_last_error": (ctypes.c_char_p, [vp]),
    "plm_get_dims": (ctypes.c_int, [vp, ctypes.POINTER(Dims)]),
    "plm_stage_offsets": (ctypes.c_int, [vp, c_int32_p, c_int32_p, c_int32_p]),
    "plm_param_offsets": (ctypes.c_int, [vp, c_int32_p]),
    "plm_jac_pattern": (ctypes.c_int, [vp, c_int32_p, c_int32_p]),
    "plm_sqp_data": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp, vp, vp, vp, vp]),
    "plm_g_data": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp, vp, vp]),
    "plm_f_data": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp, vp]),
    "plm_hess_diag": (ctypes.c_int, [vp, vp, ctypes.c_int32, vp, vp]),
    "plm_state_integrate": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp]),
    "plm_state_difference": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp]),
    "plm_rnea_dyn": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int32, vp, vp, vp]),
    "plm_aba_dyn": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int32, vp, vp, vp]),
    "plm_dyn_gaps": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp, ctypes.c_int32, vp, vp, vp]),
    "plm_centroidal_vel_gaps": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int32, vp, vp]),
    "plm_com_dyn": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp]),
    "plm_base_solve": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp, vp, ctypes.c_int32, vp, vp]),
    "plm_mpc_step": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_double, ctypes.c_int32, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, vp, vp]),
    "plm_frame_vel": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp]),
    "plm_frame_kinematics": (ctypes.c_int, [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_double), ctypes.c_int32, ctypes.POINTER(ctypes.c_double),
                                            ctypes.c_int32, vp, vp, ctypes.c_int32, vp, vp, vp]),
    "plm_qp_setup": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp]),
    "plm_qp_update": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp, vp, vp]),
    "plm_qp_solve": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp]),
    "plm_qp_get_iterates": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp]),
    "plm_qp_set_iterates": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp]),
    "plm_qp_get_scaling": (ctypes.c_int, [vp, ctypes.c_int32, vp, vp, vp, vp]),
    "plm_line_search": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int32, vp, vp, vp]),
    "plm_sqp_step": (ctypes.c_int, [vp, vp, vp, ctypes.c_int32, vp, vp, vp]),
    "plm_last_phase_ms": (ctypes.c_int, [vp, c_double_p]),
    "plm_launch_count": (ctypes.c_int64, [vp]),
    "plm_fp64_peak": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double)]),
}

_lib = None


def load():
    """Load the shared library and declare all prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(pino_locoman_b200 has no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class PlmError(RuntimeError):
    pass
