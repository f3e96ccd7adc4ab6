"""
itsim_b200 -- B200-native batched discrete-event engine behind ITsim's modelling API.

The compute path is ``libitsim_b200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/itsb.h``), reached
through ctypes; there is no PyTorch on it and no CPU fallback.
"""
from . import ir, random, units  # noqa: F401
from .model import Endpoint, Network, Simulator, Workstation  # noqa: F401

__version__ = "0.1.0"
