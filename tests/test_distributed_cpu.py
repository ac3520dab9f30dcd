"""N > 1 host logic on CPU: point sharding, the all-reduce hook and the point gather under
torch.distributed (gloo, world_size 2)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest

from caliscope_b200 import distributed as D
from caliscope_b200 import synthetic


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_point_sharding_partitions_every_observation_once(world):
    r = synthetic.make_rig(6, 400, 5000, seed=1)
    seen = np.zeros(r.n_obs, int)
    prev_hi = 0
    counts = []
    for k in range(world):
        s = D.shard_points(r.obs_cam, r.obs_pt, r.obs_xy, r.n_pts, k, world)
        assert s.pt_lo == prev_hi
        prev_hi = s.pt_hi
        seen[s.obs_index] += 1
        assert np.array_equal(s.obs_pt + s.pt_lo, r.obs_pt[s.obs_index])
        assert np.array_equal(s.obs_cam, r.obs_cam[s.obs_index])
        assert np.array_equal(s.obs_xy, r.obs_xy[s.obs_index])
        assert s.obs_pt.min() >= 0 and s.obs_pt.max() < s.n_pts
        counts.append(len(s.obs_index))
        xl = D.local_x(r.x0, 36, s)
        assert len(xl) == 36 + 3 * s.n_pts
        assert np.array_equal(xl[36:], r.x0[36 + 3 * s.pt_lo : 36 + 3 * s.pt_hi])
    assert prev_hi == r.n_pts and np.all(seen == 1)
    assert max(counts) - min(counts) <= 0.1 * r.n_obs / world + 50  # balanced by observation count


def _worker(rank: int, world: int, port: int, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) the C-ABI all-reduce hook on a host buffer
        hook = D.make_allreduce_hook(device_buffers=False)
        buf = np.arange(10, dtype=np.float64) * (rank + 1)
        rc = hook(None, buf.ctypes.data, 10, None)
        assert rc == 0
        assert np.array_equal(buf, np.arange(10) * sum(range(1, world + 1)))
        # through the ctypes function type the engine calls
        from caliscope_b200 import _lib

        cfn = _lib.ALLREDUCE_FN(hook)
        buf2 = np.full(4, float(rank))
        assert cfn(None, buf2.ctypes.data, 4, None) == 0
        assert np.array_equal(buf2, np.full(4, float(sum(range(world)))))
        # (2) gather of per-rank points back into the global vector
        r = synthetic.make_rig(6, 400, 5000, seed=1)
        s = D.shard_points(r.obs_cam, r.obs_pt, r.obs_xy, r.n_pts, rank, world)
        xl = D.local_x(r.x0, 36, s) + 0.0
        full = D.gather_points(xl, 36, r.n_pts, s)
        assert np.array_equal(full, r.x0)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_allreduce_hook_and_gather_world_size_2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


@pytest.mark.parametrize("world", [2, 3])
def test_component_aware_sharding_keeps_rigid_bodies_on_one_rank(world):
    """With rigid-distance rows the unit of assignment is a connected component of the constraint graph."""
    from tests._util import load_golden

    g, rig = load_golden("board_truss_constraints_refine0.npz")
    cons = (g["groups_a"], g["groups_b"], g["distances"], g["weights"])
    seen_obs = np.zeros(rig.n_obs, int)
    seen_con = np.zeros(len(cons[0]), int)
    seen_pts = np.zeros(rig.n_pts, int)
    for k in range(world):
        s = D.shard_points(rig.obs_cam, rig.obs_pt, rig.obs_xy, rig.n_pts, k, world, cons)
        seen_obs[s.obs_index] += 1
        seen_con[s.constraint_index] += 1
        seen_pts[s.pt_index] += 1
        assert np.array_equal(s.pt_index[s.obs_pt], rig.obs_pt[s.obs_index])
        ga, gb, dist, w = s.constraints
        assert ga.min() >= 0 and gb.min() >= 0 and ga.max() < s.n_pts and gb.max() < s.n_pts  # fully local
        assert np.array_equal(s.pt_index[ga], g["groups_a"][s.constraint_index])
        assert np.array_equal(s.pt_index[gb], g["groups_b"][s.constraint_index])
        assert np.array_equal(dist, g["distances"][s.constraint_index])
        assert len(s.constraint_index) > 0
    assert np.all(seen_obs == 1) and np.all(seen_con == 1) and np.all(seen_pts == 1)


def _cull_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from caliscope_b200 import distributed as D

        rng = np.random.default_rng(3)
        n, n_cams = 6000, 7
        err = rng.gamma(2.0, 0.4, n)
        cam = rng.integers(0, n_cams - 1, n)  # the last camera has no observations
        cam[:3] = 5
        err[cam == 4] += 10.0  # one camera far above an absolute threshold
        owner = rng.integers(0, world, n)
        sel = owner == rank
        res = {}
        for pct, floor in ((2.5, 10), (50.0, 10), (100.0, 25), (99.9, 400)):
            res[(pct, floor, "per_camera")] = D.global_cull_thresholds(err[sel], cam[sel], n_cams, pct, floor)
        res[(5.0, 10, "overall")] = D.global_cull_thresholds(err[sel], cam[sel], n_cams, 5.0, 10, scope="overall")
        out.put((rank, "ok", res))
    except Exception as e:  # pragma: no cover
        import traceback

        out.put((rank, "err: " + repr(e) + traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


def test_sharded_cull_thresholds_equal_the_single_process_filter():
    """filter_by_percentile_error + the min_per_camera floor over observations split across 2 ranks (gloo):
    the keep mask ``err <= t[camera]`` from the gathered thresholds equals filtering.keep_mask on the whole list."""
    import torch.multiprocessing as mp

    from caliscope_b200 import filtering

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cull_worker, args=(k, 2, port, out)) for k in range(2)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in results] == ["ok", "ok"], results
    rng = np.random.default_rng(3)
    n, n_cams = 6000, 7
    err = rng.gamma(2.0, 0.4, n)
    cam = rng.integers(0, n_cams - 1, n)
    cam[:3] = 5
    err[cam == 4] += 10.0
    for key, thr0 in results[0][2].items():
        pct, floor, scope = key
        assert np.array_equal(thr0, results[1][2][key])  # identical on both ranks
        keep_q = 100 - pct
        if scope == "per_camera":
            base = np.array([np.percentile(err[cam == c], keep_q) if np.any(cam == c) else np.inf for c in range(n_cams)])
        else:
            base = np.full(n_cams, np.percentile(err, keep_q))
        ref = filtering.keep_mask(err, cam, base, floor)
        assert np.array_equal(err <= thr0[cam], ref), key
        untouched = np.array([np.sum((cam == c) & (err <= base[c])) >= min(floor, np.sum(cam == c)) for c in range(n_cams)])
        assert np.array_equal(thr0[untouched], base[untouched])  # bit-identical to np.percentile where no floor applies


def test_camera_order_groups_cameras_that_share_points():
    """distributed.camera_order: on a rig with local visibility numbered ring by ring, the chosen order must cut the number of
    (point, tile pair) incidences the Schur product has to visit; on dense rigs and small rigs it is the identity; it is a
    pure function of the observation list (every rank computes the same order)."""
    from caliscope_b200 import distributed as D
    from caliscope_b200 import synthetic

    r = synthetic.make_rig(64, 20_000, 160_000, cams_per_point=8, seed=2)
    o = D.camera_order(r.obs_cam, r.obs_pt, 64, r.n_pts)
    assert sorted(o.tolist()) == list(range(64)) and not np.array_equal(o, np.arange(64))
    assert np.array_equal(o, D.camera_order(r.obs_cam.copy(), r.obs_pt.copy(), 64, r.n_pts))

    def incidences(order):
        slot = np.empty(64, np.int64)
        slot[order] = np.arange(64)
        m = np.zeros((r.n_pts, 4), bool)
        m[r.obs_pt, slot[r.obs_cam] // 16] = True
        k = m.sum(1)
        return float((k * (k + 1) / 2).mean())

    assert incidences(o) < 0.4 * incidences(np.arange(64))
    dense = synthetic.make_rig(64, 2000, 80_000, seed=1)
    assert np.array_equal(D.camera_order(dense.obs_cam, dense.obs_pt, 64, dense.n_pts), np.arange(64))
    small = synthetic.cfg2()
    assert np.array_equal(D.camera_order(small.obs_cam, small.obs_pt, 8, small.n_pts), np.arange(8))


def test_native_shard_selection_equals_numpy_path(monkeypatch):
    """distributed.shard_points cuts a rank's point range out of the caller's arrays through the library's multi-threaded
    cb_shard_select when the list is large and already has the ABI's dtypes; the NumPy path is the same selection.  Both must
    give identical shards (rows in the caller's order, local point indices), for every rank, and the shards must partition
    the list."""
    from caliscope_b200 import distributed as D
    from caliscope_b200 import synthetic

    r = synthetic.make_rig(16, 40_000, 400_000, seed=3)
    xy = np.ascontiguousarray(r.obs_xy)
    world = 5
    seen = np.zeros(len(r.obs_cam), np.int64)
    for rank in range(world):
        a = D.shard_points(r.obs_cam, r.obs_pt, xy, r.n_pts, rank, world)
        assert D._native_shard_select(r.obs_cam, r.obs_pt, xy, int(a.pt_index[0]), int(a.pt_index[-1]) + 1) is not None
        with monkeypatch.context() as m:
            m.setattr(D, "_native_shard_select", lambda *args: None)
            b = D.shard_points(r.obs_cam, r.obs_pt, xy, r.n_pts, rank, world)
        for f in ("pt_index", "obs_index", "obs_cam", "obs_pt", "obs_xy"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (rank, f)
        assert a.obs_cam.dtype == np.int32 and a.obs_pt.dtype == np.int32 and a.obs_xy.dtype == np.float64
        assert np.array_equal(r.obs_pt[a.obs_index] - a.pt_index[0], a.obs_pt)
        seen[a.obs_index] += 1
    assert (seen == 1).all()
    # lists that are small, or not in the ABI's dtypes, take the NumPy path
    assert D._native_shard_select(r.obs_cam[:1000], r.obs_pt[:1000], xy[:1000], 0, 10) is None
    assert D._native_shard_select(r.obs_cam.astype(np.int64), r.obs_pt, xy, 0, 10) is None
