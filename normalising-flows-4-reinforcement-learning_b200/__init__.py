"""B200-native coupling-flow hot path (import as ``nf_b200``).

Only what the path needs lives here: ``csrc/`` (hand-written sm_100a CUDA + the C ABI of
include/nf_b200.h), ``_lib`` (ctypes binding) and ``module`` (the host-side mirror of the
reference's ``RealNVP`` / ``MetaBlock`` / ``InvertiblePLU`` interface).
"""
from .module import InvertiblePLU, MetaBlock, RealNVP, count_parameters  # noqa: F401
from .optim import CosineLRSchedule, FlatAdamW  # noqa: F401
from .encoder import GEncoder  # noqa: F401
from .tarflow import TransformerMetaBlock, TransformerRealNVP  # noqa: F401
from .prior import StandardNormal  # noqa: F401

__all__ = ["RealNVP", "MetaBlock", "InvertiblePLU", "count_parameters", "FlatAdamW", "CosineLRSchedule", "GEncoder",
           "TransformerRealNVP", "TransformerMetaBlock", "StandardNormal"]
