"""make_ocp: name -> OCP class, kwargs merged over the defaults (reference optimization/ocp_factory.py:8-27)."""
from .ocp_centroidal_acc import OCPCentroidalAcc
from .ocp_centroidal_vel import OCPCentroidalVel
from .ocp_whole_body_aba import OCPWholeBodyABA
from .ocp_whole_body_acc import OCPWholeBodyAcc
from .ocp_whole_body_rnea import OCPWholeBodyRNEA

OCP_CLASSES = {
    "centroidal_vel": OCPCentroidalVel,
    "centroidal_acc": OCPCentroidalAcc,
    "whole_body_acc": OCPWholeBodyAcc,
    "whole_body_aba": OCPWholeBodyABA,
    "whole_body_rnea": OCPWholeBodyRNEA,
}


def make_ocp(dynamics, default_args, **kwargs):
    """``make_ocp(dynamics, OCP_ARGS[dynamics], robot=, nodes=, solver=[, batch=, device=])``."""
    if dynamics not in OCP_CLASSES:
        raise ValueError(f"Unknown dynamics type: {dynamics}")
    args = default_args.copy()
    args.update(kwargs)
    ocp = OCP_CLASSES[dynamics](**args)
    ocp.setup_problem()
    ocp.set_weights()
    return ocp
