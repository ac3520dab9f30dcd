#!/usr/bin/env python
"""Committed scipy answers for the synthetic bench workloads (seed 0), so that ``bench.py`` can print a
``parity`` block at every GPU count without spending 40-200 s of CPU per line, and so that the full-size GPU
parity tests have something to compare with when a live scipy run would not fit the test budget.

    python tests/golden/make_bench_golden.py [workload ...]      # CPU only, minutes per workload

Each entry is ``scipy.optimize.least_squares(method="trf", x_scale="jac", jac=<sparse analytic>)`` exactly as
/root/reference/src/caliscope/core/capture_volume.py:387-411 calls it, run on ``oracle/ba_oracle.py`` (the
NumPy restatement pinned against the unmodified reference by ``tests/golden/*.npz``).  Output:
``tests/golden/bench_scipy.json`` (merged: workloads not named on the command line keep their entry).
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from caliscope_b200 import synthetic  # noqa: E402
from oracle import ba_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent / "bench_scipy.json"

# name: (n_cams, n_pts, n_obs, refine_intrinsics, outlier_frac) -- must match bench.py WORKLOADS
RIGS = {
    "cfg2": (8, 2_000, 40_000, False, 0.0),
    "cfg3": (16, 10_000, 400_000, True, 0.0),
    "cfg4": (64, 50_000, 2_000_000, False, 0.0),
    "cfg4_intrinsics": (64, 50_000, 2_000_000, True, 0.0),
    "cfg4_shard8": (64, 6_250, 250_000, False, 0.0),
    "sparse64": None,  # built by synthetic.make_sparse_rig
    "cfg5_stage1": (64, 50_000, 2_000_000, False, 0.02),
}


def build(name: str):
    if name == "sparse64":
        return synthetic.sparse64()
    n_cams, n_pts, n_obs, refine, frac = RIGS[name]
    return synthetic.make_rig(n_cams, n_pts, n_obs, refine_intrinsics=refine, seed=0, outlier_frac=frac, name=name)


def main() -> None:
    names = sys.argv[1:] or ["cfg2", "cfg3", "cfg4"]
    data = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name in names:
        rig = build(name)
        orc = O.Rig(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy)
        t0 = time.perf_counter()
        res = O.solve_scipy(orc, rig.x0)
        dt = time.perf_counter() - t0
        data[name] = {
            "n_cams": rig.n_cams, "n_pts": rig.n_pts, "n_obs": rig.n_obs, "n_params": int(len(rig.x0)),
            "x0_sha_head": float(np.sum(rig.x0[:64])),  # cheap guard against a generator change
            "scipy_status": int(res.status), "scipy_nfev": int(res.nfev), "scipy_njev": int(res.njev),
            "scipy_nit": int(res.nit), "scipy_cost": float(res.cost),
            "scipy_rms_px": float(O.overall_rmse_px(res.x, orc)), "wall_s": round(dt, 2),
            "how": "oracle.ba_oracle.solve_scipy(rig, x0) with the reference's defaults (ftol=xtol=gtol=1e-8)",
        }  # fmt: skip
        print(name, json.dumps(data[name]), flush=True)
        OUT.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
