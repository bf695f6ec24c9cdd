"""scico_b200: B200-native (sm_100a) X-ray CT projector pair behind SCICO's LinearOperator API.

Hot path only: ``XRayTransform2D`` / ``XRayTransform3D`` forward projection and exact-adjoint
back projection (reference: ``scico/linop/xray/_xray2d.py``, ``_xray3d.py``).
"""

from .geometry import (angle_to_vector, convert_from_scico_geometry, convert_to_scico_geometry,
                       matrices_from_euler_angles, rotate_vectors, view_table_2d)
from .linop import LinearOperator, Operator, operator_norm, power_iteration, valid_adjoint
from .xray import XRayTransform2D, XRayTransform3D

__version__ = "0.1.0"
__all__ = [
    "XRayTransform2D",
    "XRayTransform3D",
    "LinearOperator",
    "Operator",
    "valid_adjoint",
    "power_iteration",
    "operator_norm",
    "matrices_from_euler_angles",
    "view_table_2d",
    "angle_to_vector",
    "rotate_vectors",
    "convert_to_scico_geometry",
    "convert_from_scico_geometry",
]
