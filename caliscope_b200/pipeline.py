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


def solve_filter_resolve_sharded(cam_flags, cam_const, n_pts, obs_cam, obs_pt, obs_xy, x0, *, device: int, group=None,
                                 filter_percentile: float = 2.5, min_per_camera: int = 10, ftol: float = 1e-8,
                                 verbose: int = 0) -> PipelineResult:  # fmt: skip
    """The same loop with the observations sharded by point over the ranks of ``group`` (BASELINE.json config 5 on
    several GPUs): every solve is the sharded solve of ``distributed.solve_sharded``; the cull thresholds are the
    GLOBAL per-camera percentiles (``distributed.global_cull_thresholds``: one all-gather of the pixel errors), the
    keep mask and the compaction stay local to each rank.  Every rank returns the full parameter vector and the full
    keep mask (in the caller's observation order)."""
    import torch
    import torch.distributed as dist

    from . import distributed as D

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    import os
    import time

    prof = rank == 0 and bool(os.environ.get("CB_PROFILE_PIPELINE"))
    t_last = [time.perf_counter()]

    def lap(what: str) -> None:
        if prof:
            now = time.perf_counter()
            print(f"[pipeline] {what:<42s}{1e3 * (now - t_last[0]):9.3f} ms", flush=True)
            t_last[0] = now

    obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
    obs_pt = np.ascontiguousarray(obs_pt, dtype=np.int32)
    obs_xy = np.ascontiguousarray(obs_xy, dtype=np.float64).reshape(-1, 2)
    cam_flags = np.asarray(cam_flags)
    f_median = float(np.median(np.asarray(cam_const, dtype=np.float64).reshape(-1, 9)[:, 0]))
    ncp = int(np.where(cam_flags & 1, 9, 6).sum())
    shard = D.shard_points(obs_cam, obs_pt, obs_xy, n_pts, rank, world)
    kw = dict(rank=rank, world_size=world, verbose=verbose,
              **D.transport_kwargs(device, group, n_camera_dims=len(cam_flags) * (9 if np.any(cam_flags & 1) else 6)))  # fmt: skip
    stages: list[SolveResult] = []
    rmse: list[float] = []

    def global_rmse(prob, x) -> float:
        # the shard's sum of squared pixel errors comes from the device reduction (no 16 B / observation download)
        r = prob.overall_rmse_px(x) if prob.n_obs else 0.0
        acc = torch.tensor([r * r * prob.n_obs, float(prob.n_obs)], dtype=torch.float64,
                           device="cuda" if dist.get_backend(group) == "nccl" else "cpu")  # fmt: skip
        dist.all_reduce(acc, group=group)
        acc = acc.cpu()
        return float(np.sqrt(acc[0].item() / max(acc[1].item(), 1.0)))

    lap("shard + transport")
    order = D.camera_order(obs_cam, obs_pt, len(cam_flags), n_pts, 9 if np.any(cam_flags & 1) else 6)
    lap("camera order")
    with BAProblem(cam_flags, cam_const, shard.n_pts, shard.obs_cam, shard.obs_pt, shard.obs_xy, device=device,
                   cam_order=order) as prob:
        lap("problem create")
        dist.barrier(group=group)
        lap("barrier")
        s1 = prob.solve(D.local_x(np.asarray(x0, dtype=np.float64), ncp, shard), ftol=ftol, **kw)
        stages.append(s1)
        lap("solve 1")
        rmse.append(global_rmse(prob, s1.x))
        lap("rmse 1")
        s2 = prob.solve(s1.x, loss="soft_l1", f_scale=1.0 / f_median, ftol=1e-4, max_nfev=2000, **kw)
        stages.append(s2)
        lap("solve 2")
        # per-observation euclidean pixel error from the device (evaluated with explicit round-to-nearest multiplies / adds,
        # bit-identical to np.sqrt(ex*ex + ey*ey))
        err = prob.error_order_stats(s2.x, 100.0 - filter_percentile, want_err=True)[0]
        rmse.append(global_rmse(prob, s2.x))
        lap("errors + rmse 2")
        thr = D.global_cull_thresholds(err, shard.obs_cam, len(cam_flags), filter_percentile, min_per_camera, group)
        lap("global thresholds")
        prob2, keep_local = prob.cull(s2.x, thr, 0, want_mask=True)
        lap("cull")
    with prob2:
        s3 = prob2.solve(s2.x, ftol=ftol, **kw)
        stages.append(s3)
        lap("solve 3")
        rmse.append(global_rmse(prob2, s3.x))
        lap("rmse 3")
    x = D.gather_points(s3.x, ncp, n_pts, shard, group)
    lap("gather points")
    # full keep mask: every rank contributes its observations' flags at their global positions
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    full = torch.zeros(len(obs_cam), dtype=torch.uint8, device=dev)  # one byte per observation: every row has one owner
    full[torch.as_tensor(shard.obs_index, device=dev)] = torch.as_tensor(keep_local.astype(np.uint8), device=dev)
    dist.all_reduce(full, group=group)
    keep = full.cpu().numpy().astype(bool)
    lap("keep mask")
    return PipelineResult(x=x, keep=keep, stages=stages, rmse_px=rmse)
