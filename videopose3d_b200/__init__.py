"""videopose3d_b200 — B200-native (sm_100a) execution of VideoPose3D's temporal-convolution models.

Public surface = the reference's ``common/model.py`` classes; everything else stays the reference's.
"""
from .temporal_model import TemporalModel, TemporalModelBase, TemporalModelOptimized1f
from .data_parallel import GradientReducer, broadcast_buffers

__all__ = ["TemporalModelBase", "TemporalModel", "TemporalModelOptimized1f", "GradientReducer",
           "broadcast_buffers"]
__version__ = "0.1.0"
