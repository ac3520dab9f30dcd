"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz,
produced by tests/golden/make_golden.py from /root/reference)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import ba_oracle as O
from tests._util import golden_csr, load_golden, rel_col_err

CASES = [
    "session4_refine0.npz",
    "session4_refine1.npz",
    "small_pinhole_refine0.npz",
    "small_pinhole_refine1.npz",
    "small_pinhole_constraints.npz",
    "aruco_constraints_refine0.npz",
    "aruco_constraints_refine1.npz",
    "board_truss_constraints_refine0.npz",
    "mixed_fisheye.npz",
    "ring_perfect.npz",
    "ring_noisy_refine1.npz",
]


@pytest.mark.parametrize("name", CASES)
def test_residuals_match_reference(name):
    g, rig = load_golden(name)
    r = O.residuals(g["x0"], rig)
    assert r.shape == g["r0"].shape
    # residuals are px / fx0 (~1e3): 1e-12 here is 1e-9 px, the reference's own
    # projection tolerance (tests/test_reprojection_dispatch.py:13-30)
    assert np.abs(r - g["r0"]).max() < 1e-12


@pytest.mark.parametrize("name", CASES)
def test_jacobian_matches_reference(name):
    g, rig = load_golden(name)
    J = O.jacobian(g["x0"], rig)
    Jref = golden_csr(g, rig)
    assert J.shape == Jref.shape
    J.sort_indices()
    # identical sparsity structure
    assert np.array_equal(J.indptr, Jref.indptr)
    assert np.array_equal(J.indices, Jref.indices)
    assert rel_col_err(J.toarray(), Jref.toarray()) < 1e-10


def test_jacobian_perturbed_point_mixed_models():
    g, rig = load_golden("mixed_fisheye.npz")
    assert np.abs(O.residuals(g["x1"], rig) - g["r1"]).max() < 1e-12
    J = O.jacobian(g["x1"], rig).toarray()
    assert rel_col_err(J, golden_csr(g, rig, "J1_").toarray()) < 1e-10


def test_zero_residual_at_exact_projections():
    g, rig = load_golden("ring_perfect.npz")
    assert np.abs(O.residuals(g["x0"], rig)).max() < 1e-10


def test_project_points_matches_cv2_outputs(golden_dir):
    p = dict(np.load(golden_dir / "projection.npz"))
    uv = O.project_points(p["pts"], p["rvec"], p["tvec"], p["K"], p["d5"], False)
    assert np.abs(uv - p["uv_pinhole"]).max() < 1e-9
    uv = O.project_points(p["pts"], p["rvec"], p["tvec"], p["K"], p["d4"], True)
    assert np.abs(uv - p["uv_fisheye"]).max() < 1e-9
    uv = O.project_points(p["pts"], p["rvec_tiny"], p["tvec"], p["K"], p["d5"], False)
    assert np.abs(uv - p["uv_pinhole_tiny"]).max() < 1e-9
    with pytest.raises(ValueError):
        O.project_points(p["pts"], p["rvec"], p["tvec"], p["K"], p["d5"], True)


@pytest.mark.parametrize("name", ["session4_refine0.npz", "session4_refine1.npz", "ring_noisy_refine1.npz"])
def test_pixel_rmse_matches_reference_report(name):
    g, rig = load_golden(name)
    # the reference report goes rvec -> R (unpack_into) -> rvec (cv2.Rodrigues in
    # reprojection_errors); that round trip alone moves the RMSE by ~1e-9 px
    assert abs(O.overall_rmse_px(g["x0"], rig) - float(g["rmse0"])) < 1e-8
    assert abs(O.overall_rmse_px(g["x_default"], rig) - float(g["rmse_default"])) < 1e-8
    assert abs(O.overall_rmse_px(g["x_tight"], rig) - float(g["rmse_tight"])) < 1e-8


def test_filter_error_inputs_match_reference():
    g, rig = load_golden("session4_refine0.npz")
    e = O.reproj_errors_px(g["x_default"], rig)
    assert np.abs(e - g["filt_err_xy"]).max() < 1e-9


def test_oracle_filter_rule_matches_reference_keep_mask():
    """oracle/filtering.py (thresholds + keep mask + min_per_camera floor) against the unmodified reference's
    filter_by_percentile_error on the session fixture: thresholds to the last bit, mask index for index."""
    from oracle import filtering as OF

    g, rig = load_golden("session4_refine0.npz")
    err = OF.euclidean_error(g["filt_err_xy"])
    assert np.array_equal(err, g["filt_err"])
    thr = OF.percentile_thresholds(err, g["filt_cam"], rig.n_cams, float(g["filt_percentile"]), "per_camera")
    assert np.array_equal(thr, g["filt_thresholds"])
    keep = OF.keep_mask(err, g["filt_cam"], thr, int(g["filt_min_per_camera"]))
    assert np.array_equal(keep, g["filt_keep"])
    # the floor: a threshold below every error keeps exactly min_per_camera lowest-error rows per camera
    keep0 = OF.keep_mask(err, g["filt_cam"], np.zeros(rig.n_cams), 7)
    for c in range(rig.n_cams):
        sel = g["filt_cam"] == c
        assert keep0[sel].sum() == min(7, sel.sum())
        assert err[sel][keep0[sel]].max() <= np.sort(err[sel])[min(7, sel.sum()) - 1]


@pytest.mark.parametrize("name,refine", [("session4_refine0.npz", False), ("small_pinhole_refine1.npz", True)])
def test_scipy_solve_reproduces_reference_run(name, refine):
    """Same scipy, restated fun/jac -> same trajectory as the reference's optimize()."""
    g, rig = load_golden(name)
    res = O.solve_scipy(rig, g["x0"])
    assert res.status == int(g["status_default"])
    assert res.nfev == int(g["nfev_default"])
    # rounding-level differences in J are amplified by LSMR's inexact steps; the
    # trajectory (nfev, status) is the same and the end point agrees to ~1e-8
    assert abs(res.cost - float(g["cost_default"])) < 1e-6 * float(g["cost_default"])
    assert abs(O.overall_rmse_px(res.x, rig) - float(g["rmse_default"])) < 1e-6


def test_scipy_softl1_reproduces_reference_run():
    g, rig = load_golden("session4_softl1.npz")
    res = O.solve_scipy(rig, g["x0"], loss="soft_l1", f_scale=float(g["f_scale"]))
    assert res.nfev == int(g["nfev_default"])
    assert abs(res.cost - float(g["cost_default"])) < 1e-9 * float(g["cost_default"])


@pytest.mark.parametrize("loss", ["soft_l1", "huber", "cauchy", "arctan"])
def test_robust_scaling_matches_scipy_internals(loss):
    from scipy.optimize._lsq.common import scale_for_robust_loss_function
    from scipy.optimize._lsq.least_squares import construct_loss_function

    rng = np.random.default_rng(0)
    f = rng.normal(0, 2.0, 500)
    J = rng.normal(0, 1.0, (500, 4))
    f_scale = 0.7
    lf = construct_loss_function(len(f), loss, f_scale)
    assert abs(lf(f, cost_only=True) - O.robust_cost(f, loss, f_scale)) < 1e-12
    Js, fs = scale_for_robust_loss_function(J.copy(), f.copy(), lf(f))
    js, f2 = O.robust_row_scales(f, loss, f_scale)
    # what the solver consumes: J_s^T J_s and J_s^T f_s.  (Row-wise f_s itself is
    # rounding noise / sqrt(EPS) on huber's linear branch, where rho' + 2 z rho'' == 0.)
    Jo = J * js[:, None]
    assert np.abs(Js.T @ Js - Jo.T @ Jo).max() < 1e-9 * np.abs(Js.T @ Js).max()
    assert np.abs(Js.T @ fs - Jo.T @ f2).max() < 1e-9 * np.abs(Js.T @ fs).max()
    ok = js > 1e-6
    assert np.abs(fs - f2)[ok].max() < 1e-9 * np.abs(fs[ok]).max()


def test_rotation_derivative_identity():
    """d(R(r)X)/dr = -R [X]x Jr(r), against central differences."""
    rng = np.random.default_rng(1)
    for scale in (1.0, 1e-3, 1e-7):
        r = rng.normal(0, scale, 3)
        X = rng.normal(0, 1, 3)
        R = O.rodrigues(r)[0]
        Jr = O.so3_right_jacobian(r)[0]
        Xx = np.array([[0, -X[2], X[1]], [X[2], 0, -X[0]], [-X[1], X[0], 0]])
        A = -R @ Xx @ Jr
        fd = np.zeros((3, 3))
        for k in range(3):
            d = np.zeros(3)
            d[k] = 1e-6
            fd[:, k] = (O.rodrigues(r + d)[0] @ X - O.rodrigues(r - d)[0] @ X) / 2e-6
        assert np.abs(A - fd).max() < 1e-8


# ---- the step in front of bundle adjustment: undistortion + DLT triangulation ----
def _tri_golden(golden_dir):
    return np.load(golden_dir / "triangulation.npz")


@pytest.mark.parametrize("case", ["s4", "syn"])
def test_triangulation_oracle_matches_reference(golden_dir, case):
    """oracle.triangulation.triangulate_image_points == the reference function (point_data.py:122-229):
    same keys in the same order, xyz to rounding."""
    from oracle import triangulation as T

    g = _tri_golden(golden_dir)
    pm = {int(c): g[f"{case}_proj"][i] for i, c in enumerate(g[f"{case}_cam_ids"])}
    s, o, k, xyz = T.triangulate_image_points(pm, g[f"{case}_sync"], g[f"{case}_cam"], g[f"{case}_obj"], g[f"{case}_kp"],
                                              g[f"{case}_xy"])  # fmt: skip
    assert np.array_equal(s, g[f"{case}_out_sync"])
    assert np.array_equal(o, g[f"{case}_out_obj"])
    assert np.array_equal(k, g[f"{case}_out_kp"])
    assert np.abs(xyz - g[f"{case}_out_xyz"]).max() < 1e-10


def test_undistort_oracle_is_bit_exact_against_cv2(golden_dir):
    """oracle.triangulation.undistort_points == CameraData.undistort_points (camera_array.py:135-174)."""
    from oracle import triangulation as T

    g = _tri_golden(golden_dir)
    pts = g["und_pts"]
    assert np.array_equal(T.undistort_points(pts, g["und_Kp"], g["und_d5"], False, "normalized"), g["und_pinhole_norm"])
    assert np.array_equal(T.undistort_points(pts, g["und_Kp"], g["und_d5"], False, "pixels"), g["und_pinhole_px"])
    assert np.array_equal(T.undistort_points(pts, g["und_Kf"], g["und_d4"], True, "normalized"), g["und_fisheye_norm"])
    assert np.array_equal(T.undistort_points(pts, g["und_Kf"], g["und_d4"], True, "pixels"), g["und_fisheye_px"])
    # session rows, camera by camera, as _undistort_batch does (point_data.py:236-252)
    for i, c in enumerate(g["s4_cam_ids"]):
        m = g["s4_px_cam"] == c
        got = T.undistort_points(g["s4_px"][m], g["s4_K"][i], g["s4_dist"][i], False, "normalized")
        assert np.array_equal(got, g["s4_px_undist"][m])
