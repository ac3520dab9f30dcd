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
    pt_index: np.ndarray  # global ids of the points this rank owns, ascending
    obs_index: np.ndarray  # indices into the global observation list, ascending
    obs_cam: np.ndarray
    obs_pt: np.ndarray  # LOCAL point ids (position in pt_index)
    obs_xy: np.ndarray
    constraints: tuple | None = None  # (groups_a, groups_b, distances, weights) in LOCAL point ids
    constraint_index: np.ndarray | None = None  # indices into the global constraint list

    @property
    def n_pts(self) -> int:
        return len(self.pt_index)

    # contiguous-range views kept for callers that shard without constraints
    @property
    def pt_lo(self) -> int:
        return int(self.pt_index[0]) if len(self.pt_index) else 0

    @property
    def pt_hi(self) -> int:
        return int(self.pt_index[-1]) + 1 if len(self.pt_index) else 0


def point_ranges(obs_pt: np.ndarray, n_pts: int, world_size: int) -> np.ndarray:
    """Contiguous point ranges balanced by observation count: bounds[r] .. bounds[r+1].  Large lists are balanced on
    every 16th observation (the same sample on every rank; the ranges only steer load balance, any partition is valid)."""
    obs_pt = np.asarray(obs_pt)
    if len(obs_pt) >= (1 << 20):
        obs_pt = obs_pt[::16]
    counts = np.bincount(obs_pt, minlength=n_pts)
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


def _component_labels(n_pts: int, groups_a: np.ndarray, groups_b: np.ndarray) -> np.ndarray:
    """Connected components of the constraint graph (union-find); every point gets the smallest point id of
    its component as label, so unconstrained points label themselves."""
    parent = np.arange(n_pts, dtype=np.int64)

    def find(a: int) -> int:
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    members = np.concatenate([groups_a, groups_b], axis=1).astype(np.int64)
    for row in members:
        r0 = find(int(row[0]))
        for q in row[1:]:
            rq = find(int(q))
            if rq != r0:
                lo, hi = (r0, rq) if r0 < rq else (rq, r0)
                parent[hi] = lo
                r0 = lo
    return np.array([find(int(j)) for j in range(n_pts)], dtype=np.int64)


def _native_shard_select(obs_cam, obs_pt, obs_xy, lo: int, hi: int, n_threads: int = 0):
    """The selection through the library's multi-threaded ``cb_shard_select`` when the inputs already have the ABI's
    dtypes (int32 / int32 / float64, contiguous) and the list is large; None otherwise (the NumPy path below is the same
    selection)."""
    import ctypes as C

    obs_cam, obs_xy = np.asarray(obs_cam), np.asarray(obs_xy)
    if (len(obs_pt) < (1 << 18) or obs_cam.dtype != np.int32 or obs_pt.dtype != np.int32 or obs_xy.dtype != np.float64
            or not (obs_cam.flags.c_contiguous and obs_pt.flags.c_contiguous and obs_xy.flags.c_contiguous)):  # fmt: skip
        return None
    from . import _lib as L

    lib = L.load()
    n = len(obs_pt)
    n_sel = C.c_int64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    L.check(lib.cb_shard_select(n, p(obs_cam), p(obs_pt), p(obs_xy), lo, hi, 0, C.addressof(n_sel), None, None, None, None,
                                int(n_threads)), "shard_select")  # fmt: skip
    m = n_sel.value
    sel = np.empty(m, np.int64)
    cam_l = np.empty(m, np.int32)
    pt_l = np.empty(m, np.int32)
    xy_l = np.empty((m, 2), np.float64)
    L.check(lib.cb_shard_select(n, p(obs_cam), p(obs_pt), p(obs_xy), lo, hi, m, C.addressof(n_sel), p(sel), p(cam_l), p(pt_l),
                                p(xy_l), int(n_threads)), "shard_select")  # fmt: skip
    return sel, cam_l, pt_l, xy_l


def shard_points(obs_cam, obs_pt, obs_xy, n_pts: int, rank: int, world_size: int, constraints=None) -> PointShard:
    """Partition the points (and with them the observations and constraint rows) over the ranks.

    Without constraints: contiguous point ranges balanced by observation count.  With rigid-distance
    constraints the unit of assignment is a connected component of the constraint graph, so no component is
    split across ranks (each rank eliminates its components locally); units are dealt in order of their
    first point to the rank whose cumulative observation count they fall into."""
    obs_pt = np.asarray(obs_pt)
    if constraints is None or constraints[0] is None or len(constraints[0]) == 0:
        # contiguous ranges: membership is two comparisons per observation, the local index a subtraction (this runs inside
        # every sharded call on every rank; the general path below costs ~25 ms on 2 M observations)
        bounds = point_ranges(obs_pt, n_pts, world_size)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        # every rank of the node runs this at the same moment: share the host cores instead of oversubscribing them
        native = _native_shard_select(obs_cam, obs_pt, obs_xy, lo, hi, max(2, min(8, 24 // max(world_size, 1))))
        if native is not None:
            sel, cam_l, pt_l, xy_l = native
            return PointShard(rank=rank, world_size=world_size, pt_index=np.arange(lo, hi, dtype=np.int64), obs_index=sel,
                              obs_cam=cam_l, obs_pt=pt_l, obs_xy=xy_l)  # fmt: skip
        sel = np.flatnonzero((obs_pt >= lo) & (obs_pt < hi))
        return PointShard(
            rank=rank,
            world_size=world_size,
            pt_index=np.arange(lo, hi, dtype=np.int64),
            obs_index=sel,
            obs_cam=np.ascontiguousarray(np.asarray(obs_cam)[sel], dtype=np.int32),
            obs_pt=np.ascontiguousarray(obs_pt[sel] - lo, dtype=np.int32),
            obs_xy=np.ascontiguousarray(np.asarray(obs_xy, dtype=np.float64).reshape(-1, 2)[sel]),
        )
    counts = np.bincount(obs_pt.astype(np.int64), minlength=n_pts)
    if False:
        pass
    else:
        ga, gb = np.asarray(constraints[0]).reshape(-1, 4), np.asarray(constraints[1]).reshape(-1, 4)
        label = _component_labels(n_pts, ga, gb)
        unit_counts = np.bincount(label, weights=counts.astype(np.float64), minlength=n_pts)
        is_root = label == np.arange(n_pts)
        csum = np.cumsum(np.where(is_root, unit_counts, 0.0))  # cumulative weight at each unit's first point
        total = csum[-1] if n_pts else 0.0
        start = csum - np.where(is_root, unit_counts, 0.0)
        unit_owner = np.minimum((start * world_size / max(total, 1.0)).astype(np.int64), world_size - 1)
        owner = unit_owner[label]
    pts = np.nonzero(owner == rank)[0]
    local_of = np.full(n_pts, -1, dtype=np.int64)
    local_of[pts] = np.arange(len(pts))
    sel = np.nonzero(owner[obs_pt] == rank)[0]
    shard = PointShard(
        rank=rank,
        world_size=world_size,
        pt_index=pts,
        obs_index=sel,
        obs_cam=np.ascontiguousarray(np.asarray(obs_cam)[sel], dtype=np.int32),
        obs_pt=np.ascontiguousarray(local_of[obs_pt[sel]], dtype=np.int32),
        obs_xy=np.ascontiguousarray(np.asarray(obs_xy, dtype=np.float64).reshape(-1, 2)[sel]),
    )
    if constraints is not None and constraints[0] is not None and len(constraints[0]) > 0:
        ga, gb = np.asarray(constraints[0]).reshape(-1, 4), np.asarray(constraints[1]).reshape(-1, 4)
        csel = np.nonzero(owner[ga[:, 0]] == rank)[0]
        shard.constraint_index = csel
        shard.constraints = (
            np.ascontiguousarray(local_of[ga[csel]], dtype=np.int32),
            np.ascontiguousarray(local_of[gb[csel]], dtype=np.int32),
            np.ascontiguousarray(np.asarray(constraints[2], dtype=np.float64)[csel]),
            np.ascontiguousarray(np.asarray(constraints[3], dtype=np.float64)[csel]),
        )
        assert shard.constraints[0].min(initial=0) >= 0 and shard.constraints[1].min(initial=0) >= 0
    return shard


def camera_order(obs_cam, obs_pt, n_cams: int, n_pts: int, cam_stride: int = 6, tile: int = 96) -> np.ndarray:
    """Internal camera order for a SHARDED solve (``CbBaProblemDesc.cam_order``: slot -> camera), computed from the
    whole observation list so that every rank lays the reduced camera system out identically.  Same rule as the
    engine's own choice (``choose_camera_order`` in csrc/cb_engine.cu): co-visibility counts over a sample of points,
    greedy chain (next = the unplaced camera sharing most points with the cameras of the last tile), kept only if it
    removes at least 20 % of the co-visibility mass that falls between different 96-column Schur tiles.  Dense
    rigs get the identity."""
    ident = np.arange(n_cams, dtype=np.int32)
    per_tile = max(1, tile // cam_stride)
    if n_cams * cam_stride <= 2 * tile or len(obs_cam) > n_pts * n_cams / 3.0:
        return ident
    stride = max(1, n_pts // 8192)
    obs_pt = np.asarray(obs_pt)
    sel = np.flatnonzero((obs_pt % stride) == 0)
    obs_cam = np.asarray(obs_cam)[sel].astype(np.int64)
    obs_pt = obs_pt[sel].astype(np.int64)
    sel = slice(None)
    M = np.zeros((n_pts // stride + 1, n_cams), dtype=np.float64)
    M[obs_pt[sel] // stride, obs_cam[sel]] = 1.0
    W = M.T @ M
    np.fill_diagonal(W, 0.0)
    order = [int(np.argmin(W.sum(axis=1)))]
    placed = np.zeros(n_cams, bool)
    placed[order[0]] = True
    while len(order) < n_cams:
        w = W[:, order[-per_tile:]].sum(axis=1)
        w[placed] = -1.0
        c = int(np.argmax(w))
        order.append(c)
        placed[c] = True
    order = np.asarray(order, dtype=np.int32)

    def off_mass(o):
        t = np.empty(n_cams, np.int64)
        t[o] = (np.arange(n_cams) * cam_stride) // tile
        return W[t[:, None] != t[None, :]].sum()

    return order if off_mass(order) < 0.8 * off_mass(ident) else ident


def local_x(x_global: np.ndarray, n_camera_params: int, shard: PointShard) -> np.ndarray:
    pts = x_global[n_camera_params:].reshape(-1, 3)[shard.pt_index]
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


class EngineNcclComm:
    """An NCCL communicator owned by the engine: ``cb_ba_solve`` issues ``ncclAllReduce`` on its own
    stream, with no Python in the loop.  torch.distributed only carries the 128-byte unique id
    (``cb_nccl_unique_id`` on rank 0 -> broadcast -> ``cb_nccl_comm_create`` on every rank)."""

    def __init__(self, device: int, group=None):
        import torch
        import torch.distributed as dist

        from . import _lib as L

        self._lib = L.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{device}")
        if rank == 0:
            buf = (C.c_char * 128)()
            L.check(self._lib.cb_nccl_unique_id(buf), "nccl_unique_id")
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(uid.cpu().numpy().tobytes())
        h = C.c_void_p()
        L.check(self._lib.cb_nccl_comm_create(raw, rank, world, int(device), C.byref(h)), "nccl_comm_create")
        self.handle = h.value
        self.rank, self.world_size = rank, world

    def close(self):
        if getattr(self, "handle", None):
            self._lib.cb_nccl_comm_destroy(C.c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EnginePeerGroup:
    """Symmetric CUDA-IPC buffers for the peer-memory all-reduce (``cb_peer_*``, csrc/cb_peer.cuh):
    the Schur finalize kernel sums the reduced camera system straight out of the peers' HBM over
    NVLink.  torch.distributed only carries the 64-byte IPC handles at set-up.  ``n_camera_dims`` is
    the largest ``n_cams * P`` (P = 9 with free intrinsics else 6) this group will be used for."""

    def __init__(self, device: int, n_camera_dims: int, group=None):
        import torch
        import torch.distributed as dist

        from . import _lib as L

        self._lib = L.load()
        self._group = group
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        cap = int(n_camera_dims) ** 2 + 3 * int(n_camera_dims) + 65
        h = C.c_void_p()
        mine = (C.c_char * 64)()
        L.check(self._lib.cb_peer_create(rank, world, int(device), cap, C.byref(h), mine), "peer_create")
        self.handle = h.value
        self.capacity = cap
        self.poisoned = False  # set when a solve on this group failed part-way (see BAProblem.solve)
        self.rank, self.world_size = rank, world
        send = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(f"cuda:{device}")
        recv = [torch.empty(64, dtype=torch.uint8, device=f"cuda:{device}") for _ in range(world)]
        dist.all_gather(recv, send, group=group)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in recv)
        try:
            L.check(self._lib.cb_peer_connect(C.c_void_p(self.handle), blob), "peer_connect")
            ok = 1
        except Exception:
            ok = 0
            self._connect_error = True
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{device}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.usable = bool(flag.item())

    def close(self):
        if getattr(self, "handle", None):
            import torch.distributed as dist

            if dist.is_initialized():
                dist.barrier(group=self._group)  # nobody may still be reading this rank's buffer
            self._lib.cb_peer_destroy(C.c_void_p(self.handle))
            self.handle = None


_COMMS: dict = {}
_PEERS: dict = {}


def engine_peer_group(device: int, n_camera_dims: int, group=None) -> EnginePeerGroup:
    """Cached per (device, group); re-created (collectively) when a larger reduced system comes along."""
    key = (int(device), id(group))
    g = _PEERS.get(key)
    need = int(n_camera_dims) ** 2 + 3 * int(n_camera_dims) + 65
    if g is None or g.handle is None or g.capacity < need or getattr(g, "poisoned", False):
        if g is not None:
            g.close()
        g = _PEERS[key] = EnginePeerGroup(device, n_camera_dims, group)
    return g


def engine_comm(device: int, group=None) -> EngineNcclComm:
    """Per-(device, group) cached communicator (creation is a collective: all ranks call it together)."""
    key = (int(device), id(group))
    c = _COMMS.get(key)
    if c is None or c.handle is None:
        c = _COMMS[key] = EngineNcclComm(device, group)
    return c


def close_comms() -> None:
    """Destroy every cached engine communicator (call before ``dist.destroy_process_group``)."""
    for g in list(_PEERS.values()):
        g.close()
    _PEERS.clear()
    for c in list(_COMMS.values()):
        c.close()
    _COMMS.clear()


def transport_kwargs(device: int, group=None, n_camera_dims: int | None = None) -> dict:
    """``solve(...)`` keyword arguments selecting the all-reduce transport on an NCCL process group
    (``CB_ALLREDUCE`` = ``peer`` | ``nccl`` | ``torch``):

    * ``peer``  — all-reduce over NVLink peer memory fused into the Schur finalize kernel (default when
      ``n_camera_dims`` is given; ``nccl`` is used instead when CUDA IPC cannot map the peers);
    * ``nccl``  — the engine's own NCCL communicator;
    * ``torch`` — the ``CbAllReduceSum`` callback over ``torch.distributed`` (also what the gloo CPU
      tests use, through ``make_allreduce_hook`` directly)."""
    import os

    import torch.distributed as dist

    if dist.get_backend(group) != "nccl":
        return {"allreduce": make_allreduce_hook(group)}
    mode = os.environ.get("CB_ALLREDUCE", "peer")
    if mode == "peer" and n_camera_dims is not None and dist.get_world_size(group) <= 16:
        g = engine_peer_group(device, n_camera_dims, group)
        if g.usable:
            return {"peer_group": g}
        mode = "nccl"
    if mode == "torch":
        return {"allreduce": make_allreduce_hook(group)}
    return {"nccl_comm": engine_comm(device, group)}


def gather_points(x_local: np.ndarray, n_camera_params: int, n_pts_global: int, shard: PointShard, group=None) -> np.ndarray:
    """Every rank returns the full parameter vector (cameras are already identical on all ranks)."""
    import torch
    import torch.distributed as dist

    mine = np.ascontiguousarray(x_local[n_camera_params:])
    sizes = [None] * shard.world_size
    dist.all_gather_object(sizes, int(shard.n_pts), group=group)
    backend = dist.get_backend(group)
    dev = "cuda" if backend == "nccl" else "cpu"
    maxn = max(sizes)
    send = torch.zeros(4 * maxn, dtype=torch.float64, device=dev)  # [ids | xyz]
    send[: len(shard.pt_index)] = torch.from_numpy(shard.pt_index.astype(np.float64)).to(dev)
    send[maxn : maxn + len(mine)] = torch.from_numpy(mine).to(dev)
    recv = [torch.zeros(4 * maxn, dtype=torch.float64, device=dev) for _ in sizes]
    dist.all_gather(recv, send, group=group)
    pts = np.zeros((n_pts_global, 3))
    for n, t in zip(sizes, recv):
        t = t.cpu().numpy()
        pts[t[:n].astype(np.int64)] = t[maxn : maxn + 3 * n].reshape(-1, 3)
    return np.concatenate([x_local[:n_camera_params], pts.ravel()])


def global_cull_thresholds(err_local, cam_local, n_cams: int, percentile: float, min_per_camera: int = 10, group=None,
                           scope: str = "per_camera") -> np.ndarray:
    """Per-camera pixel thresholds of ``filter_by_percentile_error`` + the ``min_per_camera`` floor of
    ``_filter_by_reprojection_thresholds`` (capture_volume.py:607-646, 709-753) when the observations are
    sharded: every rank contributes its local euclidean errors, every rank returns the identical final
    thresholds ``t`` such that the reference's keep mask is exactly ``err <= t[camera]`` (apply locally with
    ``BAProblem.cull(x, t, min_per_camera=0)``).

    One all-gather of (error, camera) — 12 B per observation over NVLink — then a two-key stable sort on the
    group's device; the two order statistics ``np.percentile`` interpolates between are read at their exact
    global positions, so the thresholds are bit-identical to the single-process ones."""
    import torch
    import torch.distributed as dist

    from .filtering import _numpy_linear_interp

    if not (0 < percentile <= 100):
        raise ValueError(f"percentile must be between 0 and 100, got {percentile}")
    if min_per_camera < 1:
        raise ValueError(f"min_per_camera must be >= 1, got {min_per_camera}")
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    e = torch.as_tensor(np.ascontiguousarray(err_local, dtype=np.float64)).to(dev)
    c = torch.as_tensor(np.ascontiguousarray(cam_local, dtype=np.int64)).to(dev)
    n_loc = torch.tensor([e.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_loc, group=group)
    sizes = [int(t.item()) for t in sizes]
    cap = max(max(sizes), 1)
    pad_e = torch.zeros(cap, dtype=torch.float64, device=dev)
    pad_c = torch.zeros(cap, dtype=torch.int64, device=dev)
    pad_e[: e.numel()] = e
    pad_c[: c.numel()] = c
    ge = [torch.empty(cap, dtype=torch.float64, device=dev) for _ in range(world)]
    gc = [torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(ge, pad_e, group=group)
    dist.all_gather(gc, pad_c, group=group)
    E = torch.cat([t[:n] for t, n in zip(ge, sizes)])
    Cm = torch.cat([t[:n] for t, n in zip(gc, sizes)])
    # sort by (camera, error): stable sort on the minor key first
    i1 = torch.sort(E, stable=True).indices
    i2 = torch.sort(Cm[i1], stable=True).indices
    Es = E[i1][i2]
    cnt = torch.bincount(Cm, minlength=n_cams).cpu().numpy().astype(np.int64)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    keep_q = 100 - percentile
    if scope == "per_camera":
        v = (cnt - 1).astype(np.float64) * (keep_q / 100.0)
        lo_i = np.floor(np.maximum(v, 0)).astype(np.int64)
        hi_i = np.minimum(lo_i + 1, np.maximum(cnt - 1, 0))
        has = cnt > 0
        pos = np.concatenate([(start + lo_i)[has], (start + hi_i)[has]])
        vals = Es[torch.as_tensor(pos, device=dev)].cpu().numpy() if has.any() else np.zeros(0)
        lo = np.zeros(n_cams)
        hi = np.zeros(n_cams)
        lo[has], hi[has] = vals[: has.sum()], vals[has.sum() :]
        thr = _numpy_linear_interp(lo, hi, v - np.floor(v))
        thr[~has] = np.inf
    elif scope == "overall":
        thr = np.full(n_cams, float(np.percentile(E.cpu().numpy(), keep_q)))
    else:
        raise ValueError(f"scope must be 'per_camera' or 'overall', got {scope}")
    # safety floor on the GLOBAL counts: a camera's sorted segment is kept up to a prefix, so restoring the
    # lowest-error dropped rows is moving the threshold to the (need)-th dropped value (capture_volume.py:626-646)
    Cs = Cm[i1][i2]
    kept = torch.bincount(Cs[Es <= torch.as_tensor(thr, device=dev)[Cs]], minlength=n_cams).cpu().numpy().astype(np.int64)
    short = np.flatnonzero((kept < min_per_camera) & (kept < cnt))
    if len(short):
        need = np.minimum(min_per_camera, cnt[short]) - kept[short]
        pos = start[short] + kept[short] + need - 1
        thr[short] = Es[torch.as_tensor(pos, device=dev)].cpu().numpy()
    return thr


def solve_sharded(cam_flags, cam_const, n_pts, obs_cam, obs_pt, obs_xy, x0, *, device: int, group=None, constraints=None,
                  **solve_kw):
    """Shard by point (by constraint component when rigid-distance rows are present), solve with one
    all-reduce of the reduced camera system per LM trial, gather."""
    import torch.distributed as dist

    from .problem import BAProblem

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shard = shard_points(obs_cam, obs_pt, obs_xy, n_pts, rank, world, constraints)
    ncp = int(np.where(np.asarray(cam_flags) & 1, 9, 6).sum())
    order = camera_order(obs_cam, obs_pt, len(np.asarray(cam_flags)), n_pts, 9 if np.any(np.asarray(cam_flags) & 1) else 6)
    with BAProblem(cam_flags, cam_const, shard.n_pts, shard.obs_cam, shard.obs_pt, shard.obs_xy,
                   constraints=shard.constraints, device=device, cam_order=order) as prob:
        dist.barrier(group=group)  # ranks enter the first fused reduce together (the peer spin times out after 20 s)
        res = prob.solve(
            local_x(np.asarray(x0, dtype=np.float64), ncp, shard),
            rank=rank,
            world_size=world,
            **transport_kwargs(device, group, n_camera_dims=len(np.asarray(cam_flags)) * (9 if np.any(np.asarray(cam_flags) & 1) else 6)),
            **solve_kw,
        )
    res.x = gather_points(res.x, ncp, n_pts, shard, group)
    return res, shard
