"""Install / remove the drop-in behind Caliscope's bundle-adjustment call.

Seam S1: the module attribute ``caliscope.core.capture_volume.least_squares``
(bound by ``from scipy.optimize import least_squares``, capture_volume.py:2) is the only thing
``CaptureVolume.optimize`` calls to solve; replacing it leaves every Caliscope API -- PointData,
CameraArray, BundleParameterization, aniposelib export -- untouched.

    import caliscope_b200.seam as seam
    with seam.installed():            # or seam.install() / seam.uninstall()
        volume = volume.optimize()    # runs on the B200
"""
from __future__ import annotations

import contextlib
import importlib

from . import solver

_TARGET = "caliscope.core.capture_volume"
_original = None


def install(fallback=None) -> None:
    """``fallback``: optional callable with scipy's ``least_squares`` signature for calls this engine does
    not implement (distance-constraint rows).  Default ``None`` = raise ``NotImplementedError``."""
    global _original
    from . import _lib

    _lib.load()  # fail loudly now if the CUDA library is missing
    mod = importlib.import_module(_TARGET)
    if _original is None:
        _original = mod.least_squares
    solver._fallback = fallback
    mod.least_squares = solver.least_squares


def uninstall() -> None:
    global _original
    if _original is not None:
        importlib.import_module(_TARGET).least_squares = _original
        _original = None
    solver._fallback = None


@contextlib.contextmanager
def installed(fallback=None):
    install(fallback)
    try:
        yield
    finally:
        uninstall()
