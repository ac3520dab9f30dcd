"""Observation sharding across GPUs (one process per GPU, ``torch.distributed``).

Every residual row touches one camera block and one point block, so the observation list
partitions by POINT with no data-path exchange except the reduced camera system: each rank
linearises its own points, forms its partial Schur complement S_g = U_g - Z_g Z_g^T, and one
sum-all-reduce of [S | b | g_c | diag U | cost] per LM trial makes the reduced system identical
on every rank.  The camera step is then solved redundantly and each rank
back-substitutes its own points.  ``torch.distributed`` is the plumbing; the buffer that is
reduced lives in the engine (``cb_engine.cu``: ``d_red``) and is handed to the hook below as a
raw device pointer on the solve stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np


@dataclass
class PointShard:
    rank: int
    world_size: int
    pt_lo: int  # this rank owns global points [pt_lo, pt_hi)
    pt_hi: int
    obs_index: np.ndarray  # indices into the global observation list, ascending
    obs_cam: np.ndarray
    obs_pt: np.ndarray  # LOCAL point ids (global - pt_lo)
    obs_xy: np.ndarray

    @property
    def n_pts(self) -> int:
        return self.pt_hi - self.pt_lo


def point_ranges(obs_pt: np.ndarray, n_pts: int, world_size: int) -> np.ndarray:
    """Contiguous point ranges balanced by observation count: bounds[r] .. bounds[r+1]."""
    counts = np.bincount(np.asarray(obs_pt, dtype=np.int64), minlength=n_pts)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, bounds[-1]), n_pts)
        bounds.append(b)
    bounds.append(n_pts)
    return np.asarray(bounds, dtype=np.int64)


def shard_points(obs_cam, obs_pt, obs_xy, n_pts: int, rank: int, world_size: int) -> PointShard:
    obs_pt = np.asarray(obs_pt)
    bounds = point_ranges(obs_pt, n_pts, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    sel = np.nonzero((obs_pt >= lo) & (obs_pt < hi))[0]
    return PointShard(
        rank=rank,
        world_size=world_size,
        pt_lo=lo,
        pt_hi=hi,
        obs_index=sel,
        obs_cam=np.ascontiguousarray(np.asarray(obs_cam)[sel], dtype=np.int32),
        obs_pt=np.ascontiguousarray(obs_pt[sel] - lo, dtype=np.int32),
        obs_xy=np.ascontiguousarray(np.asarray(obs_xy, dtype=np.float64).reshape(-1, 2)[sel]),
    )


def local_x(x_global: np.ndarray, n_camera_params: int, shard: PointShard) -> np.ndarray:
    pts = x_global[n_camera_params:].reshape(-1, 3)[shard.pt_lo : shard.pt_hi]
    return np.concatenate([x_global[:n_camera_params], pts.ravel()])


class _CudaBuffer:
    """Zero-copy view of a raw device pointer for ``torch.as_tensor``."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(n),), "typestr": "<f8", "version": 3}


def make_allreduce_hook(group=None, device_buffers: bool = True):
    """``CbAllReduceSum`` implementation over ``torch.distributed.all_reduce`` (NCCL over NVLink on
    GPUs; ``device_buffers=False`` wraps a host pointer instead, for the gloo CPU tests)."""
    import torch
    import torch.distributed as dist

    cache: dict = {}  # (ptr, n) -> tensor view, stream ptr -> ExternalStream: keep the per-call Python cost low

    def hook(_user, buf_ptr, n, stream_ptr):
        try:
            if device_buffers:
                t = cache.get((buf_ptr, n))
                if t is None:
                    t = cache[(buf_ptr, n)] = torch.as_tensor(_CudaBuffer(buf_ptr, n), device="cuda")
                if stream_ptr:
                    ext = cache.get(stream_ptr)
                    if ext is None:
                        ext = cache[stream_ptr] = torch.cuda.ExternalStream(int(stream_ptr))
                    with torch.cuda.stream(ext):
                        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                else:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                arr = np.ctypeslib.as_array(C.cast(buf_ptr, C.POINTER(C.c_double)), shape=(int(n),))
                t = torch.from_numpy(arr)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:  # never let an exception cross the C ABI
            print(f"[caliscope_b200] all-reduce hook failed: {e!r}", flush=True)
            return 1

    return hook


def gather_points(x_local: np.ndarray, n_camera_params: int, n_pts_global: int, shard: PointShard, group=None) -> np.ndarray:
    """Every rank returns the full parameter vector (cameras are already identical on all ranks)."""
    import torch
    import torch.distributed as dist

    mine = np.ascontiguousarray(x_local[n_camera_params:])
    sizes = [None] * shard.world_size
    dist.all_gather_object(sizes, (shard.pt_lo, shard.pt_hi), group=group)
    backend = dist.get_backend(group)
    dev = "cuda" if backend == "nccl" else "cpu"
    maxlen = max(3 * (hi - lo) for lo, hi in sizes)
    send = torch.zeros(maxlen, dtype=torch.float64, device=dev)
    send[: len(mine)] = torch.from_numpy(mine).to(dev)
    recv = [torch.zeros(maxlen, dtype=torch.float64, device=dev) for _ in sizes]
    dist.all_gather(recv, send, group=group)
    pts = np.zeros(3 * n_pts_global)
    for (lo, hi), t in zip(sizes, recv):
        pts[3 * lo : 3 * hi] = t[: 3 * (hi - lo)].cpu().numpy()
    return np.concatenate([x_local[:n_camera_params], pts])


def solve_sharded(cam_flags, cam_const, n_pts, obs_cam, obs_pt, obs_xy, x0, *, device: int, group=None, **solve_kw):
    """Shard by point, solve with one all-reduce of the reduced camera system per LM trial, gather."""
    import torch.distributed as dist

    from .problem import BAProblem

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shard = shard_points(obs_cam, obs_pt, obs_xy, n_pts, rank, world)
    ncp = int(np.where(np.asarray(cam_flags) & 1, 9, 6).sum())
    with BAProblem(cam_flags, cam_const, shard.n_pts, shard.obs_cam, shard.obs_pt, shard.obs_xy, device=device) as prob:
        res = prob.solve(
            local_x(np.asarray(x0, dtype=np.float64), ncp, shard),
            allreduce=make_allreduce_hook(group),
            rank=rank,
            world_size=world,
            **solve_kw,
        )
    res.x = gather_points(res.x, ncp, n_pts, shard, group)
    return res, shard
