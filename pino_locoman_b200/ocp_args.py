"""Default OCP arguments per dynamics formulation (reference ``ocp_args.py:2-19``)."""
OCP_ARGS = {
    "centroidal_vel": {"include_base": True},
    "centroidal_acc": {"include_base": True},
    "whole_body_acc": {"include_base": True},
    "whole_body_aba": {},
    "whole_body_rnea": {"tau_nodes": 3, "include_acc": True},
}
