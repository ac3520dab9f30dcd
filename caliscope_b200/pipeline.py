"""Array-level mirror of the solve / cull / re-solve loop of ``calibrate_extrinsics``
(/root/reference/src/caliscope/core/calibrate_extrinsics.py:206-250): optimize(linear) ->
optimize(soft_l1, f_scale = 1 px / median fx, ftol 1e-4, max_nfev 2000) -> per-camera percentile
cull (2.5 %) -> optimize(linear).  BASELINE.json config 5."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .filtering import percentile_thresholds
from .problem import BAProblem, SolveResult


@dataclass
class PipelineResult:
    x: np.ndarray  # full parameter vector (unobserved-after-cull points keep their last value)
    keep: np.ndarray  # bool per input observation
    stages: list  # SolveResult per solve
    rmse_px: list  # overall RMSE after each stage (on that stage's observation set)


def solve_filter_resolve(cam_flags, cam_const, n_pts, obs_cam, obs_pt, obs_xy, x0, *, filter_percentile: float = 2.5,
                         min_per_camera: int = 10, device: int = 0, ftol: float = 1e-8, verbose: int = 0,
                         want_mask: bool = True) -> PipelineResult:  # fmt: skip
    obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
    obs_pt = np.ascontiguousarray(obs_pt, dtype=np.int32)
    obs_xy = np.ascontiguousarray(obs_xy, dtype=np.float64).reshape(-1, 2)
    f_median = float(np.median(np.asarray(cam_const, dtype=np.float64).reshape(-1, 9)[:, 0]))
    stages: list[SolveResult] = []
    rmse: list[float] = []
    with BAProblem(cam_flags, cam_const, n_pts, obs_cam, obs_pt, obs_xy, device=device) as prob:
        s1 = prob.solve(x0, ftol=ftol, verbose=verbose)
        stages.append(s1)
        rmse.append(prob.overall_rmse_px(s1.x))
        s2 = prob.solve(s1.x, loss="soft_l1", f_scale=1.0 / f_median, ftol=1e-4, max_nfev=2000, verbose=verbose)
        stages.append(s2)
        rmse.append(prob.overall_rmse_px(s2.x))
        # thresholds from exact per-camera order statistics, cull + compaction on the device: the observation
        # list does not come back to the host between the cull and the re-solve
        _, thr = percentile_thresholds(prob, s2.x, filter_percentile, "per_camera", want_err=False)
        prob2, keep = prob.cull(s2.x, thr, min_per_camera, want_mask=want_mask)
    with prob2:
        s3 = prob2.solve(s2.x, ftol=ftol, verbose=verbose)
        stages.append(s3)
        rmse.append(prob2.overall_rmse_px(s3.x))
    return PipelineResult(x=s3.x, keep=keep, stages=stages, rmse_px=rmse)
