"""caliscope_b200 -- B200-native sparse bundle adjustment behind Caliscope's
``CaptureVolume.optimize()`` / ``calibrate_extrinsics()`` seam.

Hand-written sm_100a CUDA (``csrc/``) behind a C ABI (``include/caliscope_b200.h``);
this package is the host-side mirror of the reference's interface for that path.
"""
from ._lib import EngineError, EngineUnavailable  # noqa: F401
from .problem import BAProblem, SolveResult, blocks_to_arrays  # noqa: F401

__all__ = ["BAProblem", "SolveResult", "EngineError", "EngineUnavailable", "blocks_to_arrays"]
