from .ocp import OCP  # noqa: F401
from .ocp_centroidal_vel import OCPCentroidalVel  # noqa: F401
from .ocp_centroidal_acc import OCPCentroidalAcc  # noqa: F401
from .ocp_whole_body_acc import OCPWholeBodyAcc  # noqa: F401
from .ocp_whole_body_aba import OCPWholeBodyABA  # noqa: F401
from .ocp_whole_body_rnea import OCPWholeBodyRNEA  # noqa: F401
from .ocp_factory import make_ocp  # noqa: F401
