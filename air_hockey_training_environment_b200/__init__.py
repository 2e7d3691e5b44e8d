"""B200-native batched air-hockey environment step (drop-in for the `AirHockey` + `Rewards` + `computerAI`
frame loop of Positronic-IO/air-hockey-training-environment).  See DESIGN.md and include/ah_b200.h.

Importing this package does not load the CUDA library; constructing an environment does, and fails loudly
(ImportError / AhError) when the library or a CUDA device is missing.  There is no CPU fallback.
"""
from ._capi import AhError, LIB_PATH, STAT_NAMES
from .env import BatchedAirHockey, InterleavedStepper, ReplayRing, StepOutputs
from .exceptions import InvalidAgentError

__all__ = ["BatchedAirHockey", "InterleavedStepper", "ReplayRing", "StepOutputs", "InvalidAgentError", "AhError", "LIB_PATH", "STAT_NAMES"]
