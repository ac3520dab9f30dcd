"""Profiling driver: build the rig, create the device problem, run `n` solves.
Used under ncu (see profiles/README.md); numbers printed by a run under ncu are not bench values."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import caliscope_b200 as cb  # noqa: E402
from bench import make_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pcg_tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
rig = make_workload(name)
with cb.BAProblem(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy) as p:
    for i in range(n):
        r = p.solve(rig.x0, verbose=2 if i == n - 1 else 0, pcg_tol=pcg_tol)
    print(f"{name}: status {r.status} nfev {r.nfev} nit {r.nit} cost {r.cost:.12e} solve_ms {r.solve_ms:.3f} "
          f"rj_ms/launch {r.rj_ms / max(r.rj_launches, 1):.4f} launches {r.kernel_launches} pcg {r.pcg_iterations}")
