"""CPU restatement of the reference's reprojection-error filter at array level.  TEST INFRASTRUCTURE ONLY
(imported by tests/, bench.py's parity leg and __graft_entry__.smoke(); never by caliscope_b200/).

Follows /root/reference/src/caliscope/core/capture_volume.py:
  * :709-753 ``filter_by_percentile_error``: threshold = ``np.percentile(err, 100 - percentile)`` per camera
    (scope "per_camera", :732-741) or over all observations (scope "overall", :743-748);
  * :607-646 ``_filter_by_reprojection_thresholds``: keep ``err <= threshold[camera]``; a camera left with fewer
    than ``min_per_camera`` observations gets its lowest-error dropped observations back (:626-646).
Pinned by tests/test_oracle_golden.py against keep masks produced by the unmodified reference
(tests/golden/filter_*.npz).
"""
from __future__ import annotations

import numpy as np


def euclidean_error(err_xy_px: np.ndarray) -> np.ndarray:
    """capture_volume.py:183 -- ``np.sqrt(np.sum(errors_xy ** 2, axis=1))``."""
    e = np.asarray(err_xy_px, dtype=np.float64).reshape(-1, 2)
    return np.sqrt(np.sum(e**2, axis=1))


def percentile_thresholds(err: np.ndarray, obs_cam: np.ndarray, n_cams: int, percentile: float,
                          scope: str = "per_camera") -> np.ndarray:  # fmt: skip
    if not (0 < percentile <= 100):
        raise ValueError(f"percentile must be between 0 and 100, got {percentile}")
    q = 100 - percentile
    thr = np.full(n_cams, np.inf)
    if scope == "overall":
        thr[:] = np.percentile(err, q)
    elif scope == "per_camera":
        for c in range(n_cams):
            sel = obs_cam == c
            if sel.any():
                thr[c] = np.percentile(err[sel], q)
    else:
        raise ValueError(f"scope must be 'per_camera' or 'overall', got {scope}")
    return thr


def keep_mask(err: np.ndarray, obs_cam: np.ndarray, thresholds: np.ndarray, min_per_camera: int = 10) -> np.ndarray:
    obs_cam = np.asarray(obs_cam)
    keep = err <= thresholds[obs_cam]
    for c in range(len(thresholds)):
        idx = np.flatnonzero(obs_cam == c)
        n_keep = int(keep[idx].sum())
        if n_keep < min_per_camera and n_keep < len(idx):
            n_needed = min(min_per_camera, len(idx)) - n_keep
            dropped = idx[~keep[idx]]
            order = np.argsort(err[dropped], kind="stable")  # reference: nsmallest(n_needed, "euclidean_error")
            keep[dropped[order[:n_needed]]] = True
    return keep
