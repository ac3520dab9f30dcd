"""Device stages of the extrinsic bootstrap (batched planar PnP, stereo RMSE of every camera pair) through the C ABI,
against the unmodified reference's outputs (tests/golden/bootstrap_*.npz) and the oracle."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import bootstrap as OB
from tests.test_bootstrap_host import tables

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module", params=["session4", "session11"])
def case(request):
    from caliscope_b200 import bootstrap as B

    g = dict(np.load(GOLD / f"bootstrap_{request.param}.npz"))
    tab = tables(g)
    res = B.pnp_arrays(tab, g["cam_id"], g["sync_index"], g["object_id"], g["img_xy"], g["obj_xyz"])
    return g, tab, res


def test_pnp_groups_match_cv2_solvepnp(case):
    """IPPE groups == cv2.solvePnP(SOLVEPNP_IPPE) of the reference run (1e-6; the oracle agrees with cv2 to 1e-13 and the
    kernel with the oracle to 1e-9).  Groups where OpenCV's IPPE gives up (~1 %): the reference's ITERATIVE fallback starts
    from a numerically arbitrary homography there; the device fallback must reach a minimum at least as good and equal the
    oracle's fallback.  Collinear groups: NaN pose, kept (cv2 reports success)."""
    from caliscope_b200 import bootstrap as B

    g, tab, res = case
    poses = B.poses_dict(res)
    keys = [tuple(int(v) for v in k) for k in g["pnp_keys"]]
    assert list(poses) == keys  # same groups in the reference's groupby order
    norm = OB.undistort_all(g["cam_ids"], g["cam_k"], g["cam_dist"], g["cam_fisheye"], g["cam_id"], g["img_xy"])
    fb_keys: list = []
    orc = OB.pnp_poses(g["cam_ids"], norm, g["sync_index"], g["cam_id"], g["object_id"], g["obj_xyz"], fallback_keys=fb_keys)
    status = {tuple(int(v) for v in k): int(s) for k, s in zip(res.keys, res.status)}
    assert {k for k, s in status.items() if s == B.PNP_OK_FALLBACK} == set(fb_keys)
    n_fb = 0
    for i, k in enumerate(keys):
        R, t, rm = poses[k]
        Ro, to, rmo = orc[k]
        if not np.isfinite(g["pnp_R"][i]).all():
            assert status[k] == B.PNP_DEGENERATE and not np.isfinite(R).any()
            continue
        assert np.abs(R - Ro).max() < 1e-6 and np.abs(t - to).max() < 1e-6 and abs(rm - rmo) < 1e-8  # same algorithm, conditioning of 4-point groups
        if status[k] == B.PNP_OK_FALLBACK:
            n_fb += 1
            assert rm <= g["pnp_rmse"][i] * (1 + 1e-3) + 1e-9
            continue
        assert np.abs(R - g["pnp_R"][i]).max() < 1e-6 and np.abs(t - g["pnp_t"][i]).max() < 1e-6
        assert abs(rm - g["pnp_rmse"][i]) < 1e-4 * g["pnp_rmse"][i] + 1e-8  # the reference evaluates its RMSE in float32
    assert n_fb <= 0.02 * len(keys)
    assert res.launches > 0


def test_stereo_rmse_of_every_pair(case):
    """== calculate_stereo_rmse_for_pair on the reference's aggregated poses; the reference works in float32 here
    (cv2.triangulatePoints / projectPoints on float32 input), hence 2e-5 relative."""
    from caliscope_b200 import bootstrap as B

    g, tab, _ = case
    rmse, cnt = B.stereo_rmse_arrays(tab, g["agg_pairs"], g["agg_R"], g["agg_t"], g["cam_id"], g["sync_index"], g["object_id"],
                                     g["keypoint_id"], g["img_xy"])  # fmt: skip
    for i in range(len(rmse)):
        if np.isnan(g["rmse_pair"][i]):
            assert np.isnan(rmse[i])  # too few common observations, or the reference's dict-order quirk
            continue
        assert cnt[i] == g["rmse_common"][i]
        assert abs(rmse[i] - g["rmse_pair"][i]) < 2e-5 * g["rmse_pair"][i]


def test_chain_end_to_end(case):
    """PnP (device) -> relative / IQR / aggregate (host) -> stereo RMSE (device): pair set identical; aggregated poses
    equal the reference's where no fallback group is involved, within one sample's weight otherwise."""
    from caliscope_b200 import bootstrap as B

    g, tab, res = case
    live = res.status != B.PNP_TOO_FEW
    rel = B.relative_pose_arrays(res.keys[live], res.R[live], res.t[live], tab)
    pairs, _, R, t, _ = B.filter_and_aggregate(rel, 1.5)
    gold = {(int(a), int(b)): i for i, (a, b) in enumerate(g["agg_pairs"])}
    assert sorted(gold) == sorted((int(a), int(b)) for a, b in pairs)
    touched = {int(k[0]) for k, s in zip(res.keys, res.status) if s == B.PNP_OK_FALLBACK}
    for k, (a, b) in enumerate(pairs):
        i = gold[(int(a), int(b))]
        d = max(np.abs(R[k] - g["agg_R"][i]).max(), np.abs(t[k] - g["agg_t"][i]).max())
        assert d < (5e-3 if (int(a) in touched or int(b) in touched) else 1e-6), ((a, b), d)
    rmse, _ = B.stereo_rmse_arrays(tab, pairs, R, t, g["cam_id"], g["sync_index"], g["object_id"], g["keypoint_id"], g["img_xy"])
    for k, (a, b) in enumerate(pairs):
        i = gold[(int(a), int(b))]
        if np.isnan(g["rmse_pair"][i]):
            assert np.isnan(rmse[k])
        elif not (int(a) in touched or int(b) in touched):
            assert abs(rmse[k] - g["rmse_pair"][i]) < 1e-4 * g["rmse_pair"][i]


def _tiny_tables(n=3):
    from caliscope_b200 import bootstrap as B

    ids = np.arange(n, dtype=np.int64)
    k = np.tile([1000.0, 1000.0, 640.0, 360.0, 0.0], (n, 1))
    return B.CameraTables(ids, {i: i for i in range(n)}, k, np.zeros((n, 12)), np.zeros(n, np.int32), np.zeros(n, bool), np.ones(n, bool))


def _project(obj, rvec, t, k):
    from oracle.ippe import _rodrigues

    Xc = obj @ _rodrigues(np.asarray(rvec, float)).T + t
    return np.stack([k[0] * Xc[:, 0] / Xc[:, 2] + k[2], k[1] * Xc[:, 1] / Xc[:, 2] + k[3]], axis=1)


def test_pnp_edge_cases_too_few_collinear_nonplanar_unknown_camera():
    """The reference's group rules (pose_network_builder.py:272-298): fewer than min_points rows -> no pose; all model points
    on a line -> cv2 reports success with a NaN pose, kept; NaN obj_loc_z counts as 0; a camera without intrinsics is skipped;
    a non-planar group is refused (the reference switches to SQPNP)."""
    from caliscope_b200 import bootstrap as B

    tab = _tiny_tables(3)
    tab.has_intrinsics[2] = False
    k = tab.k[0]
    board = np.array([[0, 0, 0], [0.1, 0, 0], [0.1, 0.1, 0], [0, 0.1, 0], [0.05, 0.05, 0]], float)
    line = np.array([[0, 0, 0], [0.1, 0, 0], [0.2, 0, 0], [0.3, 0, 0]], float)
    rows = []

    def add(cam, sync, obj):
        xy = _project(obj, [0.2, -0.1, 0.05], np.array([0.05, -0.02, 2.0]), k)
        for kp, (o, p) in enumerate(zip(obj, xy)):
            rows.append((cam, sync, 0, kp, p[0], p[1], o[0], o[1], o[2]))

    add(0, 0, board)           # well posed
    add(0, 1, board[:3])       # too few
    add(1, 0, line)            # collinear
    nan_z = board.copy()
    add(1, 1, nan_z)           # z given as NaN below
    add(2, 0, board)           # camera without intrinsics
    a = np.array(rows)
    obj = a[:, 6:9].copy()
    obj[(a[:, 0] == 1) & (a[:, 1] == 1), 2] = np.nan
    res = B.pnp_arrays(tab, a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2].astype(np.int64), a[:, 4:6], obj)
    st = {tuple(int(v) for v in kk[:2]): int(s) for kk, s in zip(res.keys, res.status)}
    assert st == {(0, 0): B.PNP_OK, (0, 1): B.PNP_TOO_FEW, (1, 0): B.PNP_DEGENERATE, (1, 1): B.PNP_OK}
    poses = B.poses_dict(res)
    assert set(poses) == {(0, 0, 0), (1, 0, 0), (1, 1, 0)}
    assert np.isnan(poses[(1, 0, 0)][0]).all()
    from oracle.ippe import _rodrigues

    for key in ((0, 0, 0), (1, 1, 0)):
        R, t, rm = poses[key]
        # a 0.1 m board at 2 m: the pose is conditioned ~1e9 x the homography's rounding (the fp64 oracle itself is 2e-7 off)
        assert np.abs(R - _rodrigues(np.array([0.2, -0.1, 0.05]))).max() < 1e-4
        assert np.abs(t - [0.05, -0.02, 2.0]).max() < 1e-4
        assert rm < 1e-5
    # non-planar target
    cube = np.array([[0, 0, 0], [0.1, 0, 0], [0.1, 0.1, 0.05], [0, 0.1, 0], [0.05, 0.05, 0.1], [0.02, 0.07, 0.03]], float)
    xy = _project(cube, [0.1, 0.2, 0.0], np.array([0.0, 0.0, 2.0]), k)
    res2 = B.pnp_arrays(tab, np.zeros(6, np.int64), np.zeros(6, np.int64), np.zeros(6, np.int64), xy, cube)
    assert int(res2.status[0]) == B.PNP_NON_PLANAR
    with pytest.raises(NotImplementedError):
        B.poses_dict(res2)


def test_stereo_rmse_exact_geometry_and_missing_pairs():
    """Noise-free two-view geometry gives an RMSE at float32 rounding level; a pair without min_common common observations
    comes back NaN (the reference returns None there, pose_network_builder.py:656-659)."""
    from caliscope_b200 import bootstrap as B
    from oracle.ippe import _rodrigues

    tab = _tiny_tables(3)
    k = tab.k[0]
    rng = np.random.default_rng(0)
    X = np.stack([rng.uniform(-0.3, 0.3, 40), rng.uniform(-0.3, 0.3, 40), rng.uniform(1.5, 2.5, 40)], axis=1)
    R = _rodrigues(np.array([0.0, 0.4, 0.05]))
    t = np.array([-0.8, 0.02, 0.2])
    xa = np.stack([k[0] * X[:, 0] / X[:, 2] + k[2], k[1] * X[:, 1] / X[:, 2] + k[3]], axis=1)
    Xb = X @ R.T + t
    xb = np.stack([k[0] * Xb[:, 0] / Xb[:, 2] + k[2], k[1] * Xb[:, 1] / Xb[:, 2] + k[3]], axis=1)
    cam = np.concatenate([np.zeros(40), np.ones(40), np.full(3, 2)]).astype(np.int64)
    sync = np.concatenate([np.arange(40), np.arange(40), np.arange(3)]).astype(np.int64)
    z = np.zeros(len(cam), np.int64)
    xy = np.concatenate([xa, xb, xa[:3]])
    pairs = np.array([[0, 1], [0, 2]])
    Rs = np.stack([R, np.eye(3)])
    ts = np.stack([t, np.array([0.1, 0.0, 0.0])])
    rmse, cnt = B.stereo_rmse_arrays(tab, pairs, Rs, ts, cam, sync, z, z, xy)
    assert cnt.tolist() == [40, 3]
    assert rmse[0] < 5e-7 and np.isnan(rmse[1])


def test_pose_network_on_device_matches_host_arrays(case):
    """cb_relative_pose_network (relative poses, IQR rule, quaternion / translation averages of every pair in one device
    call) against the array implementation pinned on the reference (tests/test_bootstrap_host.py): same pairs, same keep
    mask row for row, same counts, aggregated poses to 1e-12 -- from the reference's own PnP poses, NaN groups included."""
    from caliscope_b200 import bootstrap as B

    g, tab, _ = case
    rel = B.relative_pose_arrays(g["pnp_keys"], g["pnp_R"], g["pnp_t"], tab)
    pairs_h, keep_h, R_h, t_h, cnt_h = B.filter_and_aggregate(rel, 1.5)
    pairs, R, t, cnt, keep, st = B.pose_network_arrays(g["pnp_keys"], g["pnp_R"], g["pnp_t"], tab, 1.5, want_keep=True)
    assert pairs.tolist() == pairs_h.tolist()
    assert cnt.tolist() == cnt_h.tolist()
    # the host rows are frame-major in the same combination order
    assert len(keep) == len(keep_h) and (keep == keep_h).all()
    assert np.abs(R - R_h).max() < 1e-12 and np.abs(t - t_h).max() < 1e-12
    gold = {(int(a), int(b)): i for i, (a, b) in enumerate(g["agg_pairs"])}
    for k, (a, b) in enumerate(pairs):
        i = gold[(int(a), int(b))]
        assert np.abs(R[k] - g["agg_R"][i]).max() < 1e-9 and np.abs(t[k] - g["agg_t"][i]).max() < 1e-9
    assert st.kernel_launches > 0
    # separate multipliers and a pass-everything threshold
    p2, R2, t2, c2, k2, _ = B.pose_network_arrays(g["pnp_keys"], g["pnp_R"], g["pnp_t"], tab, 1.5, rotation_threshold_multiplier=0.5,
                                                   translation_threshold_multiplier=3.0, want_keep=True)  # fmt: skip
    ph, kh, Rh, th, ch = B.filter_and_aggregate(rel, 1.5, 0.5, 3.0)
    assert p2.tolist() == ph.tolist() and c2.tolist() == ch.tolist() and (k2 == kh).all()
    assert np.abs(R2 - Rh).max() < 1e-12 and np.abs(t2 - th).max() < 1e-12


def test_pose_network_edge_cases():
    """No frame with two cameras -> no pair; a pair with fewer than 5 samples keeps all of them; a lone sample is passed
    through bit for bit; poses of an ignored camera form no pair; the dict-order quirk drops (b, a) combinations."""
    from caliscope_b200 import bootstrap as B
    from oracle.ippe import _rodrigues

    tab = _tiny_tables(4)
    tab.ignore[3] = True
    rng = np.random.default_rng(1)

    def pose():
        return _rodrigues(rng.normal(size=3) * 0.2), rng.normal(size=3)

    keys, Rs, ts = [], [], []
    for cam, sync in [(0, 0), (1, 0), (0, 1), (1, 1), (2, 1), (3, 1), (0, 2), (0, 3), (2, 3)]:
        R, t = pose()
        keys.append((cam, sync, 0)); Rs.append(R); ts.append(t)
    keys, Rs, ts = np.array(keys), np.array(Rs), np.array(ts)
    rel = B.relative_pose_arrays(keys, Rs, ts, tab)
    ph, kh, Rh, th, ch = B.filter_and_aggregate(rel, 1.5)
    p, R, t, c, k, _ = B.pose_network_arrays(keys, Rs, ts, tab, 1.5, want_keep=True)
    assert p.tolist() == ph.tolist() == [[0, 1], [0, 2], [1, 2]]
    assert c.tolist() == ch.tolist() == [2, 2, 1] and k.all()
    assert np.abs(R - Rh).max() < 1e-13 and np.abs(t - th).max() < 1e-13
    i12 = 2  # single sample: the relative pose itself, not a quaternion round trip (products contracted to FMAs on the device)
    assert np.abs(R[i12] - rel.R[(rel.pair_a == 1) & (rel.pair_b == 2)][0]).max() < 1e-15
    # nothing to pair
    p0, *_ = B.pose_network_arrays(keys[:1], Rs[:1], ts[:1], tab, 1.5)
    assert len(p0) == 0
    # camera ids whose dict order is not id order: (2, 1) comes first in the dict, so the pair (1, 2) is never formed
    ids = np.array([0, 2, 1], dtype=np.int64)
    tab2 = B.CameraTables(ids, {0: 0, 2: 1, 1: 2}, tab.k[:3], tab.dist[:3], tab.fisheye[:3], np.zeros(3, bool), np.ones(3, bool))
    keys2 = np.array([(0, 0, 0), (1, 0, 0), (2, 0, 0)])
    rel2 = B.relative_pose_arrays(keys2, Rs[:3], ts[:3], tab2)
    p3, *_ = B.pose_network_arrays(keys2, Rs[:3], ts[:3], tab2, 1.5)
    assert p3.tolist() == sorted({(int(a), int(b)) for a, b in zip(rel2.pair_a, rel2.pair_b)}) == [(0, 1), (0, 2)] or \
        p3.tolist() == [[0, 1], [0, 2]]
