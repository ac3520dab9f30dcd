"""Host logic of the triangulation mirror on the CPU (no GPU, no compute through the C ABI): key packing,
camera-id mapping, the reference's by-camera-set output order rebuilt from per-group signatures, empty
results and errors.  The device call ``triangulate_groups`` is replaced by an oracle-backed stand-in that
returns exactly what the kernel returns per group (xyz, count, representative row, camera-multiset signature)."""
from __future__ import annotations

import numpy as np
import pytest

from caliscope_b200 import triangulation as T
from oracle import triangulation as OT
from tests._util import fake_triangulate_groups


@pytest.fixture()
def fake_device(monkeypatch):
    monkeypatch.setattr(T, "triangulate_groups", fake_triangulate_groups)


@pytest.mark.parametrize("case", ["s4", "syn"])
def test_output_keys_and_order_equal_the_reference(golden_dir, fake_device, case):
    g = np.load(golden_dir / "triangulation.npz")
    pm = {int(c): g[f"{case}_proj"][i] for i, c in enumerate(g[f"{case}_cam_ids"])}
    s, o, k, xyz = T.triangulate_image_points(pm, g[f"{case}_sync"], g[f"{case}_cam"], g[f"{case}_obj"], g[f"{case}_kp"],
                                              g[f"{case}_xy"])  # fmt: skip
    assert np.array_equal(s, g[f"{case}_out_sync"])
    assert np.array_equal(o, g[f"{case}_out_obj"])
    assert np.array_equal(k, g[f"{case}_out_kp"])
    assert np.abs(xyz - g[f"{case}_out_xyz"]).max() < 1e-10


def test_pack_keys_is_order_isomorphic_to_lexsort():
    rng = np.random.default_rng(0)
    s = rng.integers(-5, 40, 5000)
    o = rng.integers(0, 7, 5000)
    k = rng.integers(100, 160, 5000)
    key = T.pack_keys(s, o, k)
    assert key.dtype == np.int64 and key.min() >= 0
    ref = np.lexsort((k, o, s))
    got = np.argsort(key, kind="stable")
    assert np.array_equal(ref, got)
    # same grouping as the reference's break detection (point_data.py:152-156)
    _, starts, sizes = OT.group_observations(s, o, k)
    assert len(np.unique(key)) == len(starts)
    # value ranges too wide for 62 bits fall back to dense ranks, still order-isomorphic
    big = np.array([0, 2**40, 2**41, 5], dtype=np.int64)
    key2 = T.pack_keys(big, big[::-1].copy(), big)
    assert np.array_equal(np.argsort(key2, kind="stable"), np.lexsort((big, big[::-1], big)))


def test_empty_results_and_unknown_camera(fake_device):
    pm = {0: np.hstack([np.eye(3), [[0.0], [0.0], [2.0]]]), 4: np.hstack([np.eye(3), [[0.5], [0.0], [2.0]]])}
    i64 = lambda *v: np.array(v, dtype=np.int64)  # noqa: E731
    out = T.triangulate_image_points(pm, i64(0), i64(0), i64(0), i64(0), np.zeros((1, 2)))
    assert [len(x) for x in out] == [0, 0, 0, 0] and out[3].shape == (0, 3)
    out = T.triangulate_image_points(pm, i64(0, 1), i64(0, 4), i64(0, 0), i64(0, 0), np.zeros((2, 2)))
    assert [len(x) for x in out] == [0, 0, 0, 0]
    with pytest.raises(KeyError):
        T.triangulate_image_points(pm, i64(0, 0), i64(0, 3), i64(0, 0), i64(0, 0), np.zeros((2, 2)))


def test_camera_tables_validate_their_input():
    K = np.array([[600.0, 0.5, 320.0], [0, 590.0, 240.0], [0, 0, 1]])
    fish, k, dist = T._camera_tables(K[None], [np.array([0.1, -0.05, 0.01, 0.002, 0.3])], [False])
    assert fish.tolist() == [0] and k.tolist() == [[600.0, 590.0, 320.0, 240.0, 0.5]]
    assert dist.shape == (1, 12) and dist[0, :5].tolist() == [0.1, -0.05, 0.01, 0.002, 0.3] and not dist[0, 5:].any()
    with pytest.raises(ValueError):  # fisheye cameras carry exactly 4 coefficients (reprojection.py:26-27)
        T._camera_tables(K[None], [np.zeros(5)], [True])
    with pytest.raises(ValueError):
        T._camera_tables(np.stack([K, K]), [np.zeros(5)], [False, False])
    with pytest.raises(ValueError):  # tilted sensor model
        T._camera_tables(K[None], [np.r_[np.zeros(12), 0.01, 0.0]], [False])
