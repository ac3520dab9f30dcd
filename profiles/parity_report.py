"""Parity table: caliscope_b200 (CUDA) vs scipy TRF on the oracle port, same inputs, per BASELINE config.
Writes one JSON object per configuration.  Run on the GPU box:  python profiles/parity_report.py > gpurun_out/parity.jsonl"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import caliscope_b200 as cb  # noqa: E402
from caliscope_b200 import synthetic  # noqa: E402
from oracle import ba_oracle as O  # noqa: E402
from tests._util import load_golden  # noqa: E402


def run(name, rig, x0, golden=None, **kw):
    with cb.BAProblem(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy) as p:
        p.solve(x0, **kw)
        t0 = time.perf_counter()
        res = p.solve(x0, **kw)
        t_gpu = time.perf_counter() - t0
        rm = p.overall_rmse_px(res.x)
    rec = {"config": name, "n_cams": rig.n_cams, "n_pts": rig.n_pts, "n_obs": rig.n_obs, "n_params": rig.n_params,
           "gpu": {"status": res.status, "nfev": res.nfev, "nit": res.nit, "cost": res.cost, "rms_px": rm,
                   "wall_ms": 1e3 * t_gpu, "solve_ms": res.solve_ms}}  # fmt: skip
    if golden is not None:
        rec["scipy_default"] = {"nfev": int(golden["nfev_default"]), "cost": float(golden["cost_default"]),
                                "rms_px": float(golden["rmse_default"]), "source": "reference run (tests/golden)"}  # fmt: skip
        rec["scipy_tight"] = {"cost": float(golden["cost_tight"]), "rms_px": float(golden["rmse_tight"])}
    else:
        t0 = time.perf_counter()
        ref = O.solve_scipy(rig, x0, **{k: v for k, v in kw.items() if k in ("loss", "f_scale")})
        rec["scipy_default"] = {"nfev": int(ref.nfev), "nit": int(ref.nit), "cost": float(ref.cost),
                                "rms_px": O.overall_rmse_px(ref.x, rig), "wall_s": time.perf_counter() - t0,
                                "source": "scipy on the oracle port, this box"}  # fmt: skip
    rec["abs_diff_rms_px_vs_default"] = abs(rm - rec["scipy_default"]["rms_px"])
    rec["cost_ratio_gpu_over_scipy"] = res.cost / rec["scipy_default"]["cost"]
    print(json.dumps(rec), flush=True)


g, rig = load_golden("session4_refine0.npz")
run("cfg1 tests/sessions/post_optimization, extrinsics only", rig, g["x0"], g)
g, rig = load_golden("session4_refine1.npz")
run("cfg1 + refine_intrinsics", rig, g["x0"], g)
for name, fn in (("cfg2", synthetic.cfg2), ("cfg3", synthetic.cfg3)):
    r = fn()
    rig = O.Rig(r.cam_flags, r.cam_const, r.n_pts, r.obs_cam, r.obs_pt, r.obs_xy)
    run(r.name, rig, r.x0)
