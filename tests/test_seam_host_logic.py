"""Seam S1 / S2 host logic against the unmodified reference (build container only: needs
/root/reference).  The CUDA solve is replaced by a recorder that returns the reference's own scipy
result, so what is checked is everything AROUND the solve: the arrays handed to the engine, the
vectorised observation->point map, and the CaptureVolume that comes back."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "src").exists(), reason="reference checkout only exists in the build container")


@pytest.fixture(scope="module")
def session_volume():
    for p in (str(ROOT / "tests" / "golden" / "_refshim"), str(REF / "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from caliscope.cameras.camera_array import CameraArray
    from caliscope.core.capture_volume import CaptureVolume
    from caliscope.core.point_data import ImagePoints, WorldPoints

    d = REF / "tests/sessions/post_optimization"
    ca = CameraArray.from_toml(d / "camera_array.toml")
    ip = ImagePoints.from_csv(d / "calibration/extrinsic/CHARUCO/xy_CHARUCO.csv")
    wp = WorldPoints.from_csv(d / "calibration/extrinsic/CHARUCO/xyz_CHARUCO.csv")
    return CaptureVolume(ca, ip, wp)


def _reference_run(cv, **kw):
    """Run the unmodified optimize(), recording what it hands to scipy and what scipy returns."""
    import caliscope.core.capture_volume as mod

    rec = {}
    real = mod.least_squares

    def spy(fun, x0, **k):
        res = real(fun, x0, **k)
        rec.update(x0=np.array(x0), args=k["args"], result=res, kwargs=k)
        return res

    mod.least_squares = spy
    try:
        out = cv.optimize(**kw)
    finally:
        mod.least_squares = real
    return out, rec


def test_img_to_obj_map_vectorised_equals_reference(session_volume):
    from caliscope.core.constraints import ConstraintSet
    from caliscope.core.capture_volume import CaptureVolume
    from caliscope_b200.capture_volume import fast_img_to_obj_map

    cv = session_volume
    assert np.array_equal(fast_img_to_obj_map(cv), cv.img_to_obj_map)
    # static object: world rows moved to STATIC_SYNC_INDEX, observations keep their sync_index
    from caliscope.core.point_data import STATIC_SYNC_INDEX, WorldPoints

    wdf = cv.world_points.df.copy()
    first = wdf.drop_duplicates(subset=["object_id", "keypoint_id"]).copy().reset_index(drop=True)
    first["sync_index"] = STATIC_SYNC_INDEX
    oid = int(wdf["object_id"].iloc[0])
    cs = ConstraintSet(distances=(), static_object_ids=frozenset({oid}))
    cv_static = CaptureVolume(cv.camera_array, cv.image_points, WorldPoints(first), constraints=cs)
    assert np.array_equal(fast_img_to_obj_map(cv_static), cv_static.img_to_obj_map)
    assert (cv_static.img_to_obj_map >= 0).sum() > len(first)  # many observations per static point


@pytest.mark.parametrize("refine", [False, True])
def test_s2_optimize_hands_the_engine_what_the_reference_hands_scipy(session_volume, monkeypatch, refine):
    from caliscope_b200 import capture_volume as S2
    from caliscope_b200 import solver
    from caliscope_b200.problem import SolveResult, blocks_to_arrays

    cv = session_volume
    ref_out, rec = _reference_run(cv, refine_intrinsics=refine)
    par, cam, xy, obj = rec["args"][:4]
    seen = {}

    def fake_solve_arrays(flags, const, n_pts, camera_indices, obj_indices, image_coords, x0, **kw):
        seen.update(flags=flags, const=const, n_pts=n_pts, cam=camera_indices, obj=obj_indices, xy=image_coords, x0=x0, kw=kw)
        r = rec["result"]
        return SolveResult(x=r.x.copy(), status=int(r.status), nfev=int(r.nfev), njev=int(r.njev), nit=0, cost=float(r.cost),
                           initial_cost=0.0, optimality=0.0, lambda_final=0.0, pcg_iterations=0, kernel_launches=0,
                           solve_ms=0.0, rj_ms=0.0, rj_launches=0)  # fmt: skip

    monkeypatch.setattr(solver, "solve_arrays", fake_solve_arrays)
    out = S2.optimize(cv, refine_intrinsics=refine)
    f, c = blocks_to_arrays(par.blocks)
    assert np.array_equal(seen["flags"], f) and np.array_equal(seen["const"], c)
    assert seen["n_pts"] == par.n_points
    assert np.array_equal(seen["cam"], cam) and seen["cam"].dtype == cam.dtype
    assert np.array_equal(seen["xy"], xy) and np.array_equal(seen["obj"], obj)
    assert np.array_equal(seen["x0"], rec["x0"])
    assert seen["kw"]["ftol"] == 1e-8 and seen["kw"]["loss"] == "linear" and seen["kw"]["max_nfev"] is None
    # the returned volume is what the reference builds from the same result
    assert out.optimization_status == ref_out.optimization_status
    assert np.array_equal(out.world_points.points, ref_out.world_points.points)
    for cid, camr in ref_out.camera_array.cameras.items():
        camo = out.camera_array.cameras[cid]
        assert np.array_equal(camo.rotation, camr.rotation) and np.array_equal(camo.translation, camr.translation)
        assert np.array_equal(camo.matrix, camr.matrix) and np.array_equal(camo.distortions, camr.distortions)
    assert abs(out.reprojection_report.overall_rmse - ref_out.reprojection_report.overall_rmse) == 0.0
    assert cv.optimization_status is None  # input untouched


def test_s2_strict_raises_calibration_error(session_volume, monkeypatch):
    from caliscope.exceptions import CalibrationError
    from caliscope_b200 import capture_volume as S2
    from caliscope_b200 import solver
    from caliscope_b200.problem import SolveResult

    def fake(flags, const, n_pts, cam, obj, xy, x0, **kw):
        return SolveResult(x=np.array(x0), status=0, nfev=3, njev=1, nit=0, cost=1.0, initial_cost=1.0, optimality=1.0,
                           lambda_final=0.0, pcg_iterations=0, kernel_launches=0, solve_ms=0.0, rj_ms=0.0, rj_launches=0)  # fmt: skip

    monkeypatch.setattr(solver, "solve_arrays", fake)
    with pytest.raises(CalibrationError, match="max_evaluations"):
        S2.optimize(session_volume, max_nfev=3)
    out = S2.optimize(session_volume, max_nfev=3, strict=False)
    assert out.optimization_status.converged is False and out.optimization_status.iterations == 3


def test_seam_install_full_patches_and_restores(session_volume):
    import caliscope.core.capture_volume as mod
    import caliscope_b200.seam as seam
    from caliscope_b200 import capture_volume as S2
    from caliscope_b200 import solver

    import caliscope.core.point_data as pd_mod
    from caliscope_b200 import triangulation

    def state():
        return (mod.least_squares, mod.CaptureVolume.optimize, mod.CaptureVolume._compute_img_to_obj_map,
                mod.CaptureVolume.__dict__["reprojection_report"], mod.CaptureVolume._filter_by_reprojection_thresholds,
                mod.CaptureVolume.filter_by_percentile_error, pd_mod.triangulate_image_points,
                pd_mod.ImagePoints.triangulate)  # fmt: skip

    before = state()
    with seam.installed(full=True):
        assert mod.least_squares is solver.least_squares
        assert mod.CaptureVolume.optimize is S2.optimize
        assert mod.CaptureVolume._compute_img_to_obj_map is S2.fast_img_to_obj_map
        assert mod.CaptureVolume.__dict__["reprojection_report"].func is S2.reprojection_report
        assert pd_mod.triangulate_image_points is triangulation.triangulate_image_points
        assert pd_mod.ImagePoints.triangulate is triangulation.triangulate
        assert mod.CaptureVolume._filter_by_reprojection_thresholds is S2.filter_by_reprojection_thresholds
        assert mod.CaptureVolume.filter_by_percentile_error is S2.filter_by_percentile_error
    assert state() == before


def test_s2_reprojection_report_equals_reference(session_volume, monkeypatch):
    """capture_volume.py:150-235: every field of the report, with the engine's pixel errors replaced by
    the reference's own ``reprojection_errors`` (this container has no GPU; the engine function is
    parity-tested on the GPU in tests/test_gpu_seam.py)."""
    from caliscope.core.capture_volume import CaptureVolume
    from caliscope.core.reprojection import reprojection_errors
    from caliscope_b200 import capture_volume as S2

    monkeypatch.setattr(S2, "_errors_px", reprojection_errors)
    cv = session_volume
    # drop one camera's world matches so the unmatched counters are exercised
    wdf = cv.world_points.df
    from caliscope.core.point_data import WorldPoints

    cv2 = CaptureVolume(cv.camera_array, cv.image_points, WorldPoints(wdf.iloc[: len(wdf) // 2].copy()))
    for vol in (cv, cv2):
        ref = vol.reprojection_report
        got = S2.reprojection_report(vol)
        assert got.overall_rmse == pytest.approx(ref.overall_rmse, rel=1e-13)
        assert list(got.by_camera) == list(ref.by_camera)
        for k in ref.by_camera:
            assert got.by_camera[k] == pytest.approx(ref.by_camera[k], rel=1e-12)
        assert set(got.by_point) == set(ref.by_point)
        for k in ref.by_point:
            assert got.by_point[k] == pytest.approx(ref.by_point[k], rel=1e-12)
        assert got.unmatched_by_camera == ref.unmatched_by_camera
        assert (got.n_unmatched_observations, got.n_observations_matched, got.n_observations_total, got.n_cameras,
                got.n_points) == (ref.n_unmatched_observations, ref.n_observations_matched, ref.n_observations_total,
                                  ref.n_cameras, ref.n_points)  # fmt: skip
        assert got.unmatched_rate == ref.unmatched_rate
        import pandas as pd

        pd.testing.assert_frame_equal(got.raw_errors, ref.raw_errors)


def _static_volume(cv):
    """The session with its first object declared static: world rows at STATIC_SYNC_INDEX, observations keep
    their sync_index (point_data.py:506-527)."""
    from caliscope.core.capture_volume import CaptureVolume
    from caliscope.core.constraints import ConstraintSet
    from caliscope.core.point_data import STATIC_SYNC_INDEX, WorldPoints

    wdf = cv.world_points.df.copy()
    first = wdf.drop_duplicates(subset=["object_id", "keypoint_id"]).copy().reset_index(drop=True)
    first["sync_index"] = STATIC_SYNC_INDEX
    oid = int(wdf["object_id"].iloc[0])
    cs = ConstraintSet(distances=(), static_object_ids=frozenset({oid}))
    return CaptureVolume(cv.camera_array, cv.image_points, WorldPoints(first), constraints=cs)


@pytest.mark.parametrize("kind", ["session", "static", "half_world"])
def test_s2_filters_equal_the_reference(session_volume, kind):
    """capture_volume.py:607-753: same filtered image rows, same pruned world rows (order, index, dtypes), for
    per-camera and overall percentiles, an absolute threshold, and a min_per_camera floor that has to restore rows."""
    import pandas as pd
    from caliscope.core.capture_volume import CaptureVolume
    from caliscope.core.point_data import WorldPoints
    from caliscope_b200 import capture_volume as S2

    cv = session_volume
    if kind == "static":
        cv = _static_volume(cv)
    elif kind == "half_world":
        wdf = cv.world_points.df
        cv = CaptureVolume(cv.camera_array, cv.image_points, WorldPoints(wdf.iloc[: len(wdf) // 2].copy()))

    def same(a, b):
        pd.testing.assert_frame_equal(a.image_points.df, b.image_points.df)
        pd.testing.assert_frame_equal(a.world_points.df, b.world_points.df)
        assert a.constraints == b.constraints
        assert np.array_equal(a.img_to_obj_map, b.img_to_obj_map)

    for pct, scope, floor in ((5.0, "per_camera", 10), (2.5, "overall", 10), (60.0, "per_camera", 400), (100.0, "per_camera", 7)):
        ref = cv.filter_by_percentile_error(pct, scope, floor)
        got = S2.filter_by_percentile_error(_with_s2_filter(cv), pct, scope, floor)
        same(got, ref)
    thr = {cid: 0.8 for cid in cv.camera_array.posed_cameras}
    same(S2.filter_by_reprojection_thresholds(cv, thr, 25), cv._filter_by_reprojection_thresholds(thr, 25))
    # a camera missing from the thresholds dict keeps nothing but its floor
    thr.pop(next(iter(thr)))
    same(S2.filter_by_reprojection_thresholds(cv, thr, 3), cv._filter_by_reprojection_thresholds(thr, 3))
    with pytest.raises(ValueError):
        S2.filter_by_percentile_error(cv, 0.0)
    with pytest.raises(ValueError):
        S2.filter_by_percentile_error(cv, 5.0, "per_frame")
    with pytest.raises(ValueError):
        S2.filter_by_percentile_error(cv, 5.0, "per_camera", 0)


class _S2FilterProxy:
    """``filter_by_percentile_error`` calls ``self._filter_by_reprojection_thresholds``: route that to the S2 version
    without installing the seam (which needs the CUDA library)."""

    def __init__(self, cv):
        self._cv = cv

    def __getattr__(self, name):
        return getattr(self._cv, name)

    def _filter_by_reprojection_thresholds(self, thresholds, min_per_camera):
        from caliscope_b200 import capture_volume as S2

        return S2.filter_by_reprojection_thresholds(self._cv, thresholds, min_per_camera)


def _with_s2_filter(cv):
    return _S2FilterProxy(cv)


@pytest.mark.parametrize("static", [False, True])
def test_s3_triangulate_equals_the_reference_world_points(session_volume, monkeypatch, static):
    """ImagePoints.triangulate (point_data.py:416-559) through the fused pixels -> undistort -> DLT call, the device
    call replaced by its oracle-backed stand-in: same rows, order, columns and frame times as the reference."""
    import pandas as pd
    from caliscope_b200 import triangulation as T
    from tests._util import fake_triangulate_groups

    monkeypatch.setattr(T, "triangulate_groups", fake_triangulate_groups)
    cv = session_volume
    ip, ca = cv.image_points, cv.camera_array
    ids = frozenset({int(ip.df["object_id"].iloc[0])}) if static else frozenset()
    ref = ip.triangulate(ca, static_object_ids=ids)
    got = T.triangulate(ip, ca, static_object_ids=ids)
    assert list(got.df.columns) == list(ref.df.columns) and len(got.df) == len(ref.df) > 0
    pd.testing.assert_frame_equal(got.df, ref.df, rtol=0, atol=1e-9)
    assert (got.min_index, got.max_index) == (ref.min_index, ref.max_index)
    if static:
        from caliscope.core.point_data import STATIC_SYNC_INDEX

        assert (got.df["sync_index"] == STATIC_SYNC_INDEX).sum() > 0


def test_s2_optimize_passes_the_reference_constraint_arrays(monkeypatch):
    """capture_volume.py:373-383: groups, distances and (pixel_sigma / f_median) / sigma weights."""
    import importlib.util

    from caliscope_b200 import capture_volume as S2
    from caliscope_b200 import solver
    from caliscope_b200.problem import SolveResult

    spec = importlib.util.spec_from_file_location("make_golden", ROOT / "tests" / "golden" / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    vol = mg.aruco_constraint_volume()
    ref_out, rec = _reference_run(vol, pixel_sigma=0.7)
    seen = {}

    def fake(flags, const, n_pts, cam, obj, xy, x0, **kw):
        seen.update(kw)
        r = rec["result"]
        return SolveResult(x=r.x.copy(), status=int(r.status), nfev=int(r.nfev), njev=int(r.njev), nit=0, cost=float(r.cost),
                           initial_cost=0.0, optimality=0.0, lambda_final=0.0, pcg_iterations=0, kernel_launches=0,
                           solve_ms=0.0, rj_ms=0.0, rj_launches=0)  # fmt: skip

    monkeypatch.setattr(solver, "solve_arrays", fake)
    out = S2.optimize(vol, pixel_sigma=0.7)
    ga, gb, dist, w = rec["args"][4:8]
    cons = seen["constraints"]
    assert np.array_equal(cons[0], ga) and np.array_equal(cons[1], gb)
    assert np.array_equal(cons[2], dist) and np.array_equal(cons[3], w)
    assert out.optimization_status == ref_out.optimization_status
    assert out.rigidity_report().rmse_mm == ref_out.rigidity_report().rmse_mm
    S2.optimize(vol, use_constraints=False)
    assert seen["constraints"] is None


# ---------------------------------------------------------------------------------------------------------------------
# seam S4: the bootstrap stage functions with the reference's container types (dict of StereoPair), fed with the
# reference's own PnP poses (the device stages are in tests/test_gpu_bootstrap.py)
# ---------------------------------------------------------------------------------------------------------------------
def test_s4_bootstrap_dict_stages_equal_reference(session_volume):
    from caliscope.core.bootstrap_pose import pose_network_builder as PNB
    from caliscope_b200 import bootstrap as B

    cv = session_volume
    poses = PNB.compute_camera_to_object_poses_pnp(cv.image_points, cv.camera_array)
    rel_ref = PNB.compute_relative_poses(poses, cv.camera_array)
    rel = B.compute_relative_poses(poses, cv.camera_array)
    assert set(rel) == set(rel_ref)
    for k, sp in rel.items():
        r = rel_ref[k]
        if np.isnan(r.rotation).any():
            assert np.isnan(sp.rotation).any()
            continue
        assert np.abs(sp.rotation - r.rotation).max() < 1e-12 and np.abs(sp.translation - r.translation).max() < 1e-12
        assert sp.pair == r.pair and type(sp) is type(r)
    filt_ref = PNB.reject_outliers(rel_ref, threshold=1.5)
    filt = B.reject_outliers(rel_ref, threshold=1.5)
    assert {k: len(v) for k, v in filt.items()} == {k: len(v) for k, v in filt_ref.items()}
    for k in filt:
        assert [id(sp) for sp in filt[k]] == [id(sp) for sp in filt_ref[k]]  # the very same samples survive, same order
    agg_ref = PNB.aggregate_poses(filt_ref)
    agg = B.aggregate_poses(filt_ref)
    assert list(agg) == list(agg_ref)
    for k in agg:
        assert np.abs(agg[k].rotation - agg_ref[k].rotation).max() < 1e-9
        assert np.abs(agg[k].translation - agg_ref[k].translation).max() < 1e-9


def test_s4_fused_entry_point_glue_equals_reference(session_volume, monkeypatch):
    """bootstrap.build_paired_pose_network (the fused PnP branch) with its three device calls replaced by array results
    taken from the REFERENCE's own stage functions: what is tested here, on the CPU, is the glue -- group filtering, pair
    order of the dict handed to PairedPoseNetwork.from_raw_estimates, the None-RMSE rule -- so the network must come out
    identical to the unmodified build_paired_pose_network.  (The device stages are pinned in tests/test_gpu_bootstrap.py.)"""
    from caliscope.core.bootstrap_pose import pose_network_builder as PNB
    from caliscope.core.bootstrap_pose.build_paired_pose_network import build_paired_pose_network as ref_build
    from caliscope_b200 import bootstrap as B

    cv = session_volume
    ref_net = ref_build(cv.image_points, cv.camera_array)
    poses = PNB.compute_camera_to_object_poses_pnp(cv.image_points, cv.camera_array)
    agg_ref = PNB.aggregate_poses(PNB.reject_outliers(PNB.compute_relative_poses(poses, cv.camera_array), threshold=1.5))
    common = PNB._precompute_common_observations(cv.image_points, cv.camera_array)

    def fake_pnp(tab, cam_id, sync, oid, xy, obj, min_points=4, device=0):
        keys = np.array(list(poses), dtype=np.int64).reshape(-1, 3)
        return B.PnPResult(keys, np.array([v[0] for v in poses.values()]), np.array([v[1] for v in poses.values()]),
                           np.array([v[2] for v in poses.values()]), np.zeros(len(keys), np.int32), np.zeros(len(keys), np.int32))

    def fake_network(keys, R, t, tab, threshold=1.5, *a, **k):
        rel = B.relative_pose_arrays(keys, R, t, tab)
        pairs, _, Ra, ta, cnt = B.filter_and_aggregate(rel, threshold)
        return pairs, Ra, ta, cnt, None, None

    def fake_rmse(tab, pairs, R, t, *a, **k):
        out = np.full(len(pairs), np.nan)
        for i, (a_, b_) in enumerate(pairs):
            r = PNB.calculate_stereo_rmse_for_pair(agg_ref[(int(a_), int(b_))], cv.camera_array, common)
            out[i] = np.nan if r is None else r
        return out, np.zeros(len(pairs), np.int64)

    monkeypatch.setattr(B, "pnp_arrays", fake_pnp)
    monkeypatch.setattr(B, "pose_network_arrays", fake_network)
    monkeypatch.setattr(B, "stereo_rmse_arrays", fake_rmse)
    net = B.build_paired_pose_network(cv.image_points, cv.camera_array)
    assert list(net._pairs) == list(ref_net._pairs)  # same pairs, same dict order (gap filling included)
    for k in net._pairs:
        a, b = net._pairs[k], ref_net._pairs[k]
        assert np.abs(a.rotation - b.rotation).max() < 1e-9 and np.abs(a.translation - b.translation).max() < 1e-9
        assert abs(a.error_score - b.error_score) < 1e-9


def test_seam_install_full_patches_bootstrap_and_restores():
    import caliscope.core.bootstrap_pose.pose_network_builder as pnb
    import caliscope_b200.seam as seam
    from caliscope_b200 import bootstrap as B

    import caliscope.core.bootstrap_pose.build_paired_pose_network as bpn

    before = [getattr(pnb, n) for n in seam._BOOTSTRAP_FUNCTIONS]
    top = bpn.build_paired_pose_network
    with seam.installed(full=True):
        for n in seam._BOOTSTRAP_FUNCTIONS:
            assert getattr(pnb, n) is getattr(B, n)
        assert bpn.build_paired_pose_network is B.build_paired_pose_network  # capture_volume.py:287 imports it at call time
    assert [getattr(pnb, n) for n in seam._BOOTSTRAP_FUNCTIONS] == before
    assert bpn.build_paired_pose_network is top


def test_s5_point_tables_round_trip_through_the_seam(session_volume, tmp_path):
    """ImagePoints / WorldPoints .to_csv / .from_csv behind seam S5: same bytes as the reference writes, same frames back."""
    import caliscope.core.point_data as pd_mod
    import caliscope_b200.seam as seam

    cv = session_volume
    ref_xy, ref_xyz = tmp_path / "ref_xy.csv", tmp_path / "ref_xyz.csv"
    cv.image_points.to_csv(ref_xy)
    cv.world_points.to_csv(ref_xyz)
    before = (pd_mod.ImagePoints.__dict__["to_csv"], pd_mod.ImagePoints.__dict__["from_csv"])
    with seam.installed(full=True):
        out_xy, out_xyz = tmp_path / "xy.csv", tmp_path / "xyz.csv"
        cv.image_points.to_csv(out_xy)
        cv.world_points.to_csv(out_xyz)
        assert out_xy.read_bytes() == ref_xy.read_bytes() and out_xyz.read_bytes() == ref_xyz.read_bytes()
        ip = pd_mod.ImagePoints.from_csv(out_xy)
        wp = pd_mod.WorldPoints.from_csv(out_xyz)
    assert (pd_mod.ImagePoints.__dict__["to_csv"], pd_mod.ImagePoints.__dict__["from_csv"]) == before
    import pandas as pd

    pd.testing.assert_frame_equal(ip.df, pd_mod.ImagePoints.from_csv(ref_xy).df, check_exact=True)
    pd.testing.assert_frame_equal(wp.df, pd_mod.WorldPoints.from_csv(ref_xyz).df, check_exact=True)


def test_install_from_env(monkeypatch):
    import caliscope.core.capture_volume as mod
    import caliscope_b200.seam as seam
    from caliscope_b200 import solver

    before = mod.least_squares
    monkeypatch.delenv("CALISCOPE_BA_BACKEND", raising=False)
    assert seam.install_from_env() is False and mod.least_squares is before
    monkeypatch.setenv("CALISCOPE_BA_BACKEND", "b200")
    try:
        assert seam.install_from_env() is True and mod.least_squares is solver.least_squares
    finally:
        seam.uninstall()
    assert mod.least_squares is before
