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
