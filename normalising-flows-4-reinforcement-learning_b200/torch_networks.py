"""Drop-in for the module the reference's goal-conditioned BC script imports its networks from
(``from utils.torch_networks import RealNVP, GEncoder``, il-and-cil/torch_nfgcbc.py:27; the file itself is absent from
the reference tree): ``RealNVP`` here is the causal-transformer flow with that script's constructor (:119)."""
from .encoder import GEncoder  # noqa: F401
from .tarflow import TransformerRealNVP as RealNVP  # noqa: F401
