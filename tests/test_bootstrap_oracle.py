"""The bootstrap oracle (oracle/ippe.py, oracle/bootstrap.py) against the UNMODIFIED reference's stage outputs
(tests/golden/bootstrap_*.npz, made by tests/golden/make_bootstrap_golden.py).  CPU only."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import bootstrap as OB

GOLD = Path(__file__).parent / "golden"


def load(name):
    return dict(np.load(GOLD / f"bootstrap_{name}.npz"))


@pytest.fixture(scope="module", params=["session4", "session11"])
def case(request):
    g = load(request.param)
    norm = OB.undistort_all(g["cam_ids"], g["cam_k"], g["cam_dist"], g["cam_fisheye"], g["cam_id"], g["img_xy"])
    fb: list = []
    poses = OB.pnp_poses(g["cam_ids"], norm, g["sync_index"], g["cam_id"], g["object_id"], g["obj_xyz"], fallback_keys=fb)
    return g, norm, poses, set(fb)


def golden_poses(g) -> dict:
    return {tuple(int(v) for v in k): (g["pnp_R"][i], g["pnp_t"][i], g["pnp_rmse"][i]) for i, k in enumerate(g["pnp_keys"])}


def test_pnp_poses_match_cv2_solvepnp(case):
    """IPPE groups: the restated Collins-Bartoli / Harker-O'Leary pose equals cv2.solvePnP(SOLVEPNP_IPPE) (1e-6; most
    groups 1e-13).  Collinear groups: NaN pose on both sides.  Groups where OpenCV's IPPE gives up (degenerate
    homography, ~1 % of the groups) go through cv2's ITERATIVE fallback in the reference, whose start is numerically
    arbitrary there: the restated fallback must reach a minimum at least as good as the reference's."""
    from oracle import ippe

    g, norm, poses, fb = case
    keys = [tuple(int(v) for v in k) for k in g["pnp_keys"]]
    assert list(poses) == keys  # same groups, same (groupby) order
    n_fb = 0
    for i, k in enumerate(keys):
        R, t, rm = poses[k]
        if not np.isfinite(g["pnp_R"][i]).all():
            assert not np.isfinite(R).all()
            continue
        if k in fb:
            n_fb += 1
            assert rm <= g["pnp_rmse"][i] * (1 + 1e-3) + 1e-9
            continue
        assert np.abs(R - g["pnp_R"][i]).max() < 1e-6 and np.abs(t - g["pnp_t"][i]).max() < 1e-6
        assert abs(rm - g["pnp_rmse"][i]) < 1e-4 * g["pnp_rmse"][i] + 1e-8  # the reference evaluates its RMSE in float32
    assert n_fb <= 0.02 * len(keys)


def test_relative_outlier_aggregate_chain(case):
    """Stages after PnP, fed with the REFERENCE's poses (so that the arbitrary fallback groups do not enter)."""
    g = case[0]
    poses = golden_poses(g)
    rel = OB.relative_poses(poses, g["cam_ids"], g["cam_ignore"])
    gk = {(int(a), int(b), int(s), int(o)): i for i, (a, b, s, o) in enumerate(g["rel_keys"])}
    assert len(rel) == len(gk)
    for ((a, b), s, o), (R, t) in rel.items():
        i = gk[(a, b, s, o)]
        if np.isnan(g["rel_R"][i]).any():
            assert np.isnan(R).any()
            continue
        assert np.abs(R - g["rel_R"][i]).max() < 1e-12 and np.abs(t - g["rel_t"][i]).max() < 1e-12
    filt = OB.reject_outliers(rel, 1.5)
    for (a, b), n in zip(g["filt_pairs"], g["filt_count"]):
        assert len(filt[(int(a), int(b))]) == int(n)
    agg = OB.aggregate(filt)
    assert sorted(agg) == sorted((int(a), int(b)) for a, b in g["agg_pairs"])
    for i, (a, b) in enumerate(g["agg_pairs"]):
        R, t = agg[(int(a), int(b))]
        assert np.abs(R - g["agg_R"][i]).max() < 1e-9 and np.abs(t - g["agg_t"][i]).max() < 1e-9


def test_whole_chain_from_own_poses(case):
    """End to end from the oracle's own PnP poses: identical pair set; aggregated poses equal where no fallback group
    is involved (session4: everywhere), within the spread of one sample otherwise."""
    g, _, poses, fb = case
    agg = OB.aggregate(OB.reject_outliers(OB.relative_poses(poses, g["cam_ids"], g["cam_ignore"]), 1.5))
    assert sorted(agg) == sorted((int(a), int(b)) for a, b in g["agg_pairs"])
    touched = {c for c, _, _ in fb}
    for i, (a, b) in enumerate(g["agg_pairs"]):
        R, t = agg[(int(a), int(b))]
        d = max(np.abs(R - g["agg_R"][i]).max(), np.abs(t - g["agg_t"][i]).max())
        assert d < (5e-3 if (int(a) in touched or int(b) in touched) else 1e-6), ((a, b), d)


def test_stereo_rmse_and_network(case):
    g, norm = case[0], case[1]
    agg = {(int(a), int(b)): (g["agg_R"][i], g["agg_t"][i]) for i, (a, b) in enumerate(g["agg_pairs"])}
    rm = OB.stereo_rmse(agg, g["cam_ids"], g["cam_ignore"], norm, g["sync_index"], g["cam_id"], g["object_id"], g["keypoint_id"])
    raw = {}
    for i, (a, b) in enumerate(g["agg_pairs"]):
        v = rm[(int(a), int(b))]
        if np.isnan(g["rmse_pair"][i]):
            assert v is None
            continue
        # the reference evaluates this in float32 (projectPoints / triangulatePoints on float32 input)
        assert abs(v - g["rmse_pair"][i]) < 2e-5 * g["rmse_pair"][i]
        raw[(int(a), int(b))] = (g["agg_R"][i], g["agg_t"][i], float(g["rmse_pair"][i]))
    net = OB.fill_network(raw)
    assert sorted(net) == sorted((int(a), int(b)) for a, b in g["net_pairs"])
    for i, (a, b) in enumerate(g["net_pairs"]):
        R, t, e = net[(int(a), int(b))]
        assert np.abs(R - g["net_R"][i]).max() < 1e-9 and np.abs(t - g["net_t"][i]).max() < 1e-9
        assert abs(e - g["net_err"][i]) < 1e-12
