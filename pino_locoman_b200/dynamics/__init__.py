from .dynamics import (Dynamics, DynamicsCentroidalAcc, DynamicsCentroidalVel, DynamicsWholeBodyAcc,  # noqa: F401
                       DynamicsWholeBodyTorque)
