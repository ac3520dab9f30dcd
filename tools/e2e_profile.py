"""Where the end-to-end time of one cfg4 solve goes (CB_PROFILE_CREATE=1 prints the stages of problem creation)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import caliscope_b200 as cb  # noqa: E402
from bench import make_parameterization, make_workload  # noqa: E402
from caliscope_b200 import reprojection as R  # noqa: E402
from caliscope_b200 import solver  # noqa: E402

rig = make_workload("cfg4")
par = make_parameterization(rig)
cam16 = rig.obs_cam.astype(np.int16)
xy = np.array(rig.obs_xy)
obj = np.array(rig.obs_pt, dtype=np.int32)
for it in range(4):
    t0 = time.perf_counter()
    p = cb.BAProblem(rig.cam_flags, rig.cam_const, rig.n_pts, cam16, obj, xy)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r = p.solve(rig.x0)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    p.close()
    t3 = time.perf_counter()
    print(f"iter {it}: create {1e3*(t1-t0):.3f} ms solve {1e3*(t2-t1):.3f} ms (device {r.solve_ms:.3f}) close "
          f"{1e3*(t3-t2):.3f} ms mode {r.used_graph_mode}", flush=True)
for it in range(3):
    t0 = time.perf_counter()
    res = solver.least_squares(R.joint_residuals, rig.x0, args=(par, cam16, xy, obj, None, None, None, None),
                               jac=R.joint_jacobian, x_scale="jac", method="trf", bounds=par.bounds(), ftol=1e-8)
    print(f"least_squares total {1e3*(time.perf_counter()-t0):.3f} ms", flush=True)
