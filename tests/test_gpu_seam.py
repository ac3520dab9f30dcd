"""GPU tests of the reference-facing Python surface: the reprojection mirror functions and the
least_squares drop-in, called exactly the way CaptureVolume.optimize calls scipy
(capture_volume.py:387-411), checked against the oracle and the reference goldens."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import ba_oracle as O
from tests._util import golden_csr, load_golden, rel_col_err

pytestmark = pytest.mark.gpu


def mirror_parameterization(g, rig):
    """BundleParameterization (mirror class) equivalent to a golden case's blocks."""
    from caliscope_b200.bundle_parameterization import BundleParameterization, CameraBlock

    blocks = []
    for i in range(rig.n_cams):
        c = rig.cam_const[i]
        fish = bool(rig.cam_flags[i] & 2)
        blocks.append(
            CameraBlock(
                cam_id=int(g["cam_ids"][i]), free_intrinsics=bool(rig.cam_flags[i] & 1), fx_initial=c[0], fy_initial=c[1],
                cx=c[2], cy=c[3], fisheye=fish, dist_fixed=tuple(c[4:8]) if fish else tuple(c[6:9]),
                k1_initial=0.0 if fish else c[4], k2_initial=0.0 if fish else c[5],
            )
        )  # fmt: skip
    return BundleParameterization(blocks=tuple(blocks), n_points=rig.n_pts)


@pytest.mark.parametrize("name", ["session4_refine1.npz", "mixed_fisheye.npz", "small_pinhole_refine0.npz"])
def test_joint_residuals_and_jacobian_mirror(name):
    from caliscope_b200 import reprojection as R

    g, rig = load_golden(name)
    par = mirror_parameterization(g, rig)
    cam = rig.obs_cam.astype(np.int16)
    r = R.joint_residuals(g["x0"], par, cam, rig.obs_xy, rig.obs_pt)
    assert np.abs(r - g["r0"][: 2 * rig.n_obs]).max() < 1e-12
    J = R.joint_jacobian(g["x0"], par, cam, rig.obs_xy, rig.obs_pt)
    Jref = golden_csr(g, rig)[: 2 * rig.n_obs]
    J.sort_indices()
    assert J.shape == Jref.shape
    assert np.array_equal(J.indptr, Jref.indptr) and np.array_equal(J.indices, Jref.indices)
    assert rel_col_err(J.toarray(), Jref.toarray()) < 1e-10
    R.clear_cache()


def test_project_points_and_reprojection_errors_mirror(golden_dir):
    from caliscope_b200 import reprojection as R

    p = dict(np.load(golden_dir / "projection.npz"))
    assert np.abs(R.project_points(p["pts"], p["rvec"], p["tvec"], p["K"], p["d5"], False) - p["uv_pinhole"]).max() < 1e-9
    assert np.abs(R.project_points(p["pts"], p["rvec"], p["tvec"], p["K"], p["d4"], True) - p["uv_fisheye"]).max() < 1e-9
    assert np.abs(R.project_points(p["pts"], p["rvec_tiny"], p["tvec"], p["K"], p["d5"], False) - p["uv_pinhole_tiny"]).max() < 1e-9
    with pytest.raises(ValueError):
        R.project_points(p["pts"], p["rvec"], p["tvec"], p["K"], p["d5"], True)


@pytest.mark.parametrize("name", ["small_pinhole_constraints.npz", "aruco_constraints_refine1.npz"])
def test_constraint_rows_in_the_mirror_functions(name):
    """joint_residuals / joint_jacobian with the four constraint arrays (reprojection.py:112-117, 207-226)."""
    from caliscope_b200 import reprojection as R

    g, rig = load_golden(name)
    par = mirror_parameterization(g, rig)
    args = (par, rig.obs_cam, rig.obs_xy, rig.obs_pt, g["groups_a"], g["groups_b"], g["distances"], g["weights"])
    r = R.joint_residuals(g["x0"], *args)
    assert r.shape == g["r0"].shape
    assert np.abs(r - g["r0"]).max() < 1e-12
    J = R.joint_jacobian(g["x0"], *args)
    Jref = golden_csr(g, rig)
    assert J.shape == Jref.shape
    assert rel_col_err(J.toarray(), Jref.toarray()) < 1e-10
    n_rows = 2 * rig.n_obs
    assert np.abs(J.toarray()[n_rows:] - Jref.toarray()[n_rows:]).max() < 1e-9
    R.clear_cache()


def test_bad_constraint_indices_are_rejected():
    import caliscope_b200 as cb

    g, rig = load_golden("small_pinhole_constraints.npz")
    bad = g["groups_a"].copy()
    bad[0, 0] = rig.n_pts
    with pytest.raises(cb.EngineError):
        cb.BAProblem(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy,
                     constraints=(bad, g["groups_b"], g["distances"], g["weights"]))
    with pytest.raises(ValueError):
        cb.BAProblem(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy,
                     constraints=(g["groups_a"], g["groups_b"][:1], g["distances"], g["weights"]))


@pytest.mark.parametrize("name,refine", [("session4_refine0.npz", False), ("session4_refine1.npz", True)])
def test_least_squares_drop_in_called_like_the_reference(name, refine):
    """The exact call of capture_volume.py:387-411 with our module attribute in place of scipy's."""
    from caliscope_b200 import reprojection as R
    from caliscope_b200.solver import least_squares

    g, rig = load_golden(name)
    par = mirror_parameterization(g, rig)
    result = least_squares(
        R.joint_residuals,
        g["x0"],
        args=(par, rig.obs_cam.astype(np.int16), rig.obs_xy, rig.obs_pt, None, None, None, None),
        jac=R.joint_jacobian,
        verbose=0,
        x_scale="jac",
        loss="linear",
        f_scale=1.0,
        ftol=1e-8,
        max_nfev=None,
        method="trf",
        bounds=par.bounds(),
    )
    assert result.status in (1, 2, 3, 4) and result.success
    assert result.nfev >= 2 and np.isfinite(result.cost)
    assert result.cost <= float(g["cost_default"]) * (1 + 1e-9)
    rm = O.overall_rmse_px(result.x, rig)
    noise = abs(float(g["rmse_default"]) - float(g["rmse_tight"]))
    assert abs(rm - float(g["rmse_default"])) < (1e-6 if not refine else 3 * noise + 1e-6)
    assert par.bound_warnings(result.x) == ()
