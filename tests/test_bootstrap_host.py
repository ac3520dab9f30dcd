"""Host bookkeeping of the bootstrap mirror (caliscope_b200/bootstrap.py: relative poses, IQR outlier rule, quaternion
averaging) against the unmodified reference's stage outputs (tests/golden/bootstrap_*.npz).  CPU only: these stages take
per-group results as input; the device stages are in tests/test_gpu_bootstrap.py."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from caliscope_b200 import bootstrap as B

GOLD = Path(__file__).parent / "golden"


def tables(g) -> B.CameraTables:
    ids = g["cam_ids"].astype(np.int64)
    return B.CameraTables(ids, {int(c): i for i, c in enumerate(ids)}, g["cam_k"], g["cam_dist"], g["cam_fisheye"].astype(np.int32),
                          g["cam_ignore"].astype(bool), np.ones(len(ids), bool))  # fmt: skip


@pytest.mark.parametrize("name", ["session4", "session11"])
def test_relative_filter_aggregate_match_reference(name):
    g = dict(np.load(GOLD / f"bootstrap_{name}.npz"))
    tab = tables(g)
    rel = B.relative_pose_arrays(g["pnp_keys"], g["pnp_R"], g["pnp_t"], tab)
    gk = {(int(a), int(b), int(s), int(o)): i for i, (a, b, s, o) in enumerate(g["rel_keys"])}
    assert len(rel.pair_a) == len(gk)  # incl. the reference's dict-order quirk (pairs never formed)
    idx = np.array([gk[(int(a), int(b), int(s), int(o))] for a, b, s, o in zip(rel.pair_a, rel.pair_b, rel.sync, rel.obj)])
    ok = np.isfinite(g["rel_R"][idx]).all(axis=(1, 2))
    assert np.array_equal(ok, np.isfinite(rel.R).all(axis=(1, 2)))
    assert np.abs(rel.R[ok] - g["rel_R"][idx][ok]).max() < 1e-12
    assert np.abs(rel.t[ok] - g["rel_t"][idx][ok]).max() < 1e-12
    pairs, keep, R, t, counts = B.filter_and_aggregate(rel, 1.5)
    gold_cnt = {(int(a), int(b)): int(n) for (a, b), n in zip(g["filt_pairs"], g["filt_count"])}
    for (a, b), n in zip(pairs, counts):
        assert gold_cnt[(int(a), int(b))] == int(n)
    gold_agg = {(int(a), int(b)): i for i, (a, b) in enumerate(g["agg_pairs"])}
    assert sorted(gold_agg) == sorted((int(a), int(b)) for a, b in pairs)
    for k, (a, b) in enumerate(pairs):
        i = gold_agg[(int(a), int(b))]
        assert np.abs(R[k] - g["agg_R"][i]).max() < 1e-9 and np.abs(t[k] - g["agg_t"][i]).max() < 1e-9


def test_quaternions_round_trip():
    rng = np.random.default_rng(0)
    from oracle.ippe import _rodrigues

    Rs = np.array([_rodrigues(rng.normal(0, 1.5, 3)) for _ in range(200)])
    q = B._quat_wxyz(Rs)
    back = np.array([B._quat_to_matrix(v) for v in q])
    assert np.abs(back - Rs).max() < 1e-12
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-12


def test_board_session_chain_recovers_the_rig():
    """Synthetic board session -> oracle PnP -> this package's relative / IQR / average stages: the aggregated relative poses
    are the generator's cameras (sub-millimetre, sub-milliradian at 0.3 px noise); checks the bench workload's generator and the
    array stages together, without a GPU."""
    from caliscope_b200 import synthetic
    from oracle import bootstrap as OB

    s = synthetic.make_board_session(12, 80, seed=4)
    norm = OB.undistort_all(s.cam_ids, s.cam_k, s.cam_dist, s.cam_fisheye, s.cam_id, s.img_xy)
    poses = OB.pnp_poses(s.cam_ids, norm, s.sync_index, s.cam_id, s.object_id, s.obj_xyz)
    keys = np.array(list(poses), np.int64)
    R = np.array([v[0] for v in poses.values()])
    t = np.array([v[1] for v in poses.values()])
    tab = B.CameraTables(s.cam_ids, {int(c): i for i, c in enumerate(s.cam_ids)}, s.cam_k, s.cam_dist, s.cam_fisheye,
                         np.zeros(12, bool), np.ones(12, bool))  # fmt: skip
    rel = B.relative_pose_arrays(keys, R, t, tab)
    pairs, keep, Ra, ta, cnt = B.filter_and_aggregate(rel, 1.5)
    assert len(pairs) > 20 and keep.sum() > 0.7 * len(keep)
    worst_R = worst_t = 0.0
    for k, (a, b) in enumerate(pairs):
        if cnt[k] < 5:
            continue
        RA, RB = synthetic._rot(s.rvec[a]), synthetic._rot(s.rvec[b])
        worst_R = max(worst_R, np.abs(Ra[k] - RB @ RA.T).max())
        worst_t = max(worst_t, np.abs(ta[k] - (s.tvec[b] - RB @ RA.T @ s.tvec[a])).max())
    assert worst_R < 5e-3 and worst_t < 2e-2
