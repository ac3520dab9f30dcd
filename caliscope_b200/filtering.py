"""Reprojection-error filtering at array level (the part of ``CaptureVolume`` filtering that is
arithmetic): /root/reference/src/caliscope/core/capture_volume.py:607-646 (keep mask with the
``min_per_camera`` safety floor) and :709-753 (per-camera / overall percentile thresholds).

Pixel errors and the per-camera order statistics come from the CUDA engine
(``cb_ba_error_order_stats``); the threshold interpolation between the two order statistics repeats
``np.percentile``'s default 'linear' rule bit for bit, so the keep mask is index-exact.
"""
from __future__ import annotations

import numpy as np

from .problem import BAProblem


def _numpy_linear_interp(a: np.ndarray, b: np.ndarray, t: np.ndarray) -> np.ndarray:
    """numpy.lib._function_base_impl._lerp for the 'linear' percentile method."""
    d = b - a
    out = a + d * t
    hi = t >= 0.5
    out[hi] = b[hi] - d[hi] * (1 - t[hi])
    out[d == 0] = a[d == 0]
    return out


def percentile_thresholds(prob: BAProblem, x, percentile: float, scope: str = "per_camera", want_err: bool = True):
    """(euclidean error per observation or None, threshold per camera index).  ``percentile`` is the share of
    worst observations to remove, as in ``filter_by_percentile_error``; cameras without observations get +inf."""
    if not (0 < percentile <= 100):
        raise ValueError(f"percentile must be between 0 and 100, got {percentile}")
    keep_q = 100 - percentile
    err, lo, hi, cnt = prob.error_order_stats(x, keep_q, want_err=want_err or scope == "overall")
    if scope == "per_camera":
        v = (cnt - 1).astype(np.float64) * (keep_q / 100.0)
        t = v - np.floor(v)
        thr = _numpy_linear_interp(lo, hi, t)
        thr[cnt == 0] = np.inf
        return err, thr
    if scope == "overall":
        return err, np.full(prob.n_cams, float(np.percentile(err, keep_q)))
    raise ValueError(f"scope must be 'per_camera' or 'overall', got {scope}")


def keep_mask(err: np.ndarray, obs_cam: np.ndarray, thresholds: np.ndarray, min_per_camera: int = 10) -> np.ndarray:
    """``error <= threshold[camera]``, then per camera restore the lowest-error observations until
    ``min_per_camera`` are kept (capture_volume.py:622-646)."""
    if min_per_camera < 1:
        raise ValueError(f"min_per_camera must be >= 1, got {min_per_camera}")
    obs_cam = np.asarray(obs_cam)
    keep = err <= thresholds[obs_cam]
    kept = np.bincount(obs_cam[keep], minlength=len(thresholds))
    total = np.bincount(obs_cam, minlength=len(thresholds))
    for c in np.nonzero((kept < min_per_camera) & (kept < total))[0]:
        sel = obs_cam == c
        need = min(min_per_camera, int(total[c])) - int(kept[c])
        dropped = np.sort(err[sel & ~keep])
        if len(dropped) >= need:
            keep[sel] = err[sel] <= dropped[need - 1]
    return keep


def filter_by_percentile_error(prob: BAProblem, x, obs_cam, percentile: float, scope: str = "per_camera",
                               min_per_camera: int = 10):  # fmt: skip
    err, thr = percentile_thresholds(prob, x, percentile, scope)
    return keep_mask(err, obs_cam, thr, min_per_camera), err, thr


def filter_by_absolute_error(prob: BAProblem, x, obs_cam, max_pixels: float, min_per_camera: int = 10):
    if max_pixels <= 0:
        raise ValueError(f"max_pixels must be positive, got {max_pixels}")
    e = prob.reproj_errors_px(x)
    err = np.sqrt(np.sum(e * e, axis=1))
    return keep_mask(err, obs_cam, np.full(prob.n_cams, float(max_pixels)), min_per_camera), err
