"""dynamic_multiview_3d_b200 -- B200-native appearance-flow training hot path.

Host-side mirror of the reference's operator and model interfaces for the path
(tf_utils.py, appearance_flow_model.py, ...), over the C ABI of libdmv3d.so
(include/dmv3d.h; hand-written sm_100a CUDA in csrc/).  See DESIGN.md.
"""
from . import _lib, functional, tf_utils  # noqa: F401
from .appearance_flow_model import (AppearanceFlowModel, AppearanceFlowTinghui, AppFlowHighDimAngle,  # noqa: F401
                                    AppFlowLowDimAngle)
from .main_model import Base_Prediction_Model  # noqa: F401
from .multiobject_appflow import MultiObjectAppFlow, MultiViewFusionAppFlow  # noqa: F401
from .optimizer import TFAdam  # noqa: F401
from .variables import VariableStore, use_store  # noqa: F401

__version__ = "0.1.0"
