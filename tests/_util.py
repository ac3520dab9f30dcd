"""Shared helpers for the test-suite (test infrastructure)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
from scipy.sparse import csr_matrix

from oracle.ba_oracle import Rig

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(name: str):
    g = dict(np.load(GOLDEN / name, allow_pickle=False))
    rig = Rig(
        cam_flags=g["cam_flags"],
        cam_const=g["cam_const"],
        n_pts=int(g["n_pts"]),
        obs_cam=g["obs_cam"],
        obs_pt=g["obs_pt"],
        obs_xy=g["obs_xy"],
        groups_a=g.get("groups_a"),
        groups_b=g.get("groups_b"),
        distances=g.get("distances"),
        weights=g.get("weights"),
    )
    return g, rig


def golden_csr(g, rig: Rig, prefix: str = "J_") -> csr_matrix:
    n_rows = 2 * rig.n_obs + rig.n_constraints
    return csr_matrix((g[prefix + "data"], g[prefix + "indices"], g[prefix + "indptr"]), shape=(n_rows, rig.n_params))


def rel_col_err(A: np.ndarray, B: np.ndarray) -> float:
    """Per-column error scaled like the reference's Jacobian test
    (tests/synthetic/test_analytic_jacobian.py:37-50)."""
    diff = np.abs(A - B)
    scale = np.maximum(np.abs(B).max(axis=0), 1e-3)
    return float((diff.max(axis=0) / scale).max())


def fake_triangulate_groups(proj, obs_cam, obs_key, obs_xy, *, undistort=None, **_):
    """CPU stand-in for ``caliscope_b200.triangulation.triangulate_groups`` with the kernel's output contract
    (groups in ascending key order: xyz, count, representative row, camera-multiset signature), built on the
    oracle: optional per-camera undistortion (float32 in / out like the reference), one SVD per group."""
    import hashlib

    from oracle import triangulation as OT

    obs_cam = np.asarray(obs_cam)
    obs_key = np.asarray(obs_key)
    obs_xy = np.asarray(obs_xy, dtype=np.float64).reshape(-1, 2)
    if undistort is not None:
        mats, dists, fish = undistort
        xy = np.empty_like(obs_xy)
        for c in np.unique(obs_cam):
            m = obs_cam == c
            xy[m] = OT.undistort_points(obs_xy[m], mats[c], dists[c], bool(fish[c]), "normalized")
        obs_xy = xy
    order = np.argsort(obs_key, kind="stable")
    keys = obs_key[order]
    starts = np.flatnonzero(np.concatenate([[True], keys[1:] != keys[:-1]]))
    ends = np.concatenate([starts[1:], [len(order)]])
    xyz, count, rep, sig = [], [], [], []
    for b, e in zip(starts, ends):
        rows = order[b:e]
        count.append(e - b)
        rep.append(rows[0])
        h = hashlib.sha256(np.sort(obs_cam[rows]).astype(np.int64).tobytes()).digest()
        sig.append(np.frombuffer(h[:16], dtype=np.uint64))
        if e - b < 2:
            xyz.append([np.nan] * 3)
            continue
        A = np.concatenate([np.stack([obs_xy[r, 0] * proj[obs_cam[r], 2] - proj[obs_cam[r], 0],
                                      obs_xy[r, 1] * proj[obs_cam[r], 2] - proj[obs_cam[r], 1]]) for r in rows])  # fmt: skip
        w = np.linalg.svd(A, full_matrices=False)[2][-1]
        xyz.append(w[:3] / w[3])
    return np.array(xyz).reshape(-1, 3), np.array(count, np.int32), np.array(rep, np.int32), np.array(sig).reshape(-1, 2)
