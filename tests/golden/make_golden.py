#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (mprib/caliscope,
read-only at /root/reference) on its own fixtures.

Build-container tooling: needs /root/reference, opencv and scipy; it is never run
on the GPU box (the .npz files travel instead).  Re-run with

    python tests/golden/make_golden.py

What is stored per case: the exact arrays ``CaptureVolume.optimize`` hands to
scipy (capture_volume.py:346-365), the reference's ``joint_residuals`` /
``joint_jacobian`` outputs at x0, the scipy results (default and tightened
tolerances) and the pixel RMSE the reference reports for them.
"""

from __future__ import annotations

import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
sys.path.insert(0, str(HERE / "_refshim"))
sys.path.insert(0, str(REF / "src"))

import cv2  # noqa: E402
import numpy as np  # noqa: E402
from scipy.optimize import least_squares  # noqa: E402

from caliscope.cameras.camera_array import CameraArray, CameraData  # noqa: E402
from caliscope.core.bundle_parameterization import BundleParameterization  # noqa: E402
from caliscope.core.capture_volume import CaptureVolume  # noqa: E402
from caliscope.core.point_data import ImagePoints, WorldPoints  # noqa: E402
from caliscope.core.reprojection import (  # noqa: E402
    joint_jacobian,
    joint_residuals,
    project_points,
    reprojection_errors,
)


def blocks_to_arrays(par: BundleParameterization):
    flags, const = [], []
    for b in par.blocks:
        f = (1 if b.free_intrinsics else 0) | (2 if b.fisheye else 0)
        if b.fisheye:
            d = list(b.dist_fixed) + [0.0]
            c = [b.fx_initial, b.fy_initial, b.cx, b.cy, d[0], d[1], d[2], d[3], 0.0]
        else:
            d = list(b.dist_fixed)
            c = [b.fx_initial, b.fy_initial, b.cx, b.cy, b.k1_initial, b.k2_initial, d[0], d[1], d[2]]
        flags.append(f)
        const.append(c)
    return np.array(flags, np.int32), np.array(const, np.float64), np.array([b.cam_id for b in par.blocks], np.int32)


def ba_arrays(cv: CaptureVolume):
    """capture_volume.py:346-358."""
    matched = cv.img_to_obj_map >= 0
    posed = cv.image_points.df["cam_id"].isin(set(cv.camera_array.posed_cam_id_to_index)).to_numpy()
    m = matched & posed
    df = cv.image_points.df[m]
    idx = cv.camera_array.posed_cam_id_to_index
    cam = np.array([idx[c] for c in df["cam_id"]], dtype=np.int16)
    xy = df[["img_loc_x", "img_loc_y"]].values.astype(np.float64)
    return cam, xy, cv.img_to_obj_map[m].astype(np.int32), m


def px_rmse(cv: CaptureVolume, par: BundleParameterization, x: np.ndarray) -> float:
    """What reprojection_report.overall_rmse gives after unpack_into(x)."""
    from copy import deepcopy

    ca = deepcopy(cv.camera_array)
    pts = par.unpack_into(ca, x)
    cam, xy, obj, _ = ba_arrays(cv)
    e = reprojection_errors(ca, cam, xy, pts[obj])
    return float(np.sqrt(np.mean(np.sum(e * e, axis=1))))


def csr_parts(J):
    J = J.tocsr()
    J.sort_indices()
    return {"J_data": J.data, "J_indices": J.indices.astype(np.int64), "J_indptr": J.indptr.astype(np.int64)}


def solve_case(cv: CaptureVolume, refine: bool, loss="linear", f_scale=1.0, tight=True, groups=None):
    cam, xy, obj, _ = ba_arrays(cv)
    par = BundleParameterization.from_camera_array(
        cv.camera_array, n_points=len(cv.world_points.points), refine_intrinsics=refine
    )
    x0 = par.pack(cv.camera_array, cv.world_points.points)
    ga = gb = gd = gw = None
    if groups is not None:
        ga, gb, gd, gw = groups
    args = (par, cam, xy, obj, ga, gb, gd, gw)
    flags, const, cam_ids = blocks_to_arrays(par)
    out = {
        "cam_flags": flags,
        "cam_const": const,
        "cam_ids": cam_ids,
        "n_pts": np.int64(par.n_points),
        "obs_cam": cam.astype(np.int32),
        "obs_pt": obj,
        "obs_xy": xy,
        "x0": x0,
        "r0": joint_residuals(x0, *args),
        "rmse0": px_rmse(cv, par, x0),
        "loss": np.array(loss),
        "f_scale": np.float64(f_scale),
    }
    if groups is not None:
        out.update(groups_a=ga, groups_b=gb, distances=gd, weights=gw)
    out.update(csr_parts(joint_jacobian(x0, *args)))
    kw = dict(jac=joint_jacobian, x_scale="jac", loss=loss, f_scale=f_scale, method="trf", bounds=par.bounds())
    res = least_squares(joint_residuals, x0, args=args, ftol=1e-8, **kw)
    out.update(
        x_default=res.x,
        cost_default=res.cost,
        nfev_default=res.nfev,
        njev_default=res.njev,
        status_default=res.status,
        rmse_default=px_rmse(cv, par, res.x),
    )
    if tight:
        rt = least_squares(joint_residuals, x0, args=args, ftol=1e-15, xtol=1e-15, gtol=1e-15, max_nfev=300, **kw)
        out.update(
            x_tight=rt.x,
            cost_tight=rt.cost,
            nfev_tight=rt.nfev,
            status_tight=rt.status,
            rmse_tight=px_rmse(cv, par, rt.x),
        )
    return out


def constraint_groups(cv: CaptureVolume, pixel_sigma: float = 1.0):
    """Exactly what optimize() builds (capture_volume.py:373-383)."""
    arrays = cv._build_constraint_arrays()
    assert arrays is not None
    ga, gb, dist, sig = arrays
    f_median = float(np.median([c.matrix[0, 0] for c in cv.camera_array.posed_cameras.values()]))
    return ga, gb, dist, (pixel_sigma / f_median) / sig


def aruco_constraint_volume():
    """tests/synthetic/test_rigid_constraints.py:32-58 plus a second mobile marker linked by a centre distance."""
    from caliscope.core.aruco_marker import ArucoMarker, ArucoMarkerSet, DistanceLink
    from caliscope.synthetic.camera_synthesizer import CameraSynthesizer
    from caliscope.synthetic.scene_factories import aruco_scene
    from caliscope.synthetic.se3_pose import SE3Pose
    from caliscope.synthetic.trajectory import Trajectory

    camera_array = CameraSynthesizer().add_ring(n=4, radius=2.0, height=0.5).build()
    markers = {0: ArucoMarker(0, 0.1), 1: ArucoMarker(1, 0.1, static=True)}
    marker_set = ArucoMarkerSet(dictionary=cv2.aruco.DICT_4X4_50, markers=markers)
    static_pose = SE3Pose.from_axis_angle(axis=np.array([0.0, 0.0, 1.0]), angle_rad=0.0, translation=np.array([0.3, -0.2, 0.0]))
    trajectories = {0: Trajectory.orbital(n_frames=20, radius=0.5), 1: Trajectory.stationary(n_frames=20, pose=static_pose)}
    scene, constraints = aruco_scene(marker_set=marker_set, trajectories=trajectories, camera_array=camera_array,
                                     pixel_noise_sigma=0.5)  # fmt: skip
    return CaptureVolume.bootstrap(scene.image_points_noisy, scene.intrinsics_only_cameras(), constraints=constraints)


def board_truss_volume():
    """A 5x7 planar grid (the synthetic charuco stand-in) with the reference's truss + brace constraints."""
    from caliscope.core.constraints import ConstraintSet
    from caliscope.synthetic.scene_factories import default_ring_scene

    ring = default_ring_scene(pixel_noise_sigma=0.5, random_seed=42)
    obj = ring.objects[0].calibration_object
    truss = ConstraintSet._truss_distance_constraints(np.asarray(obj.points), 0.05, 0.002)
    cs = ConstraintSet(distances=truss, static_object_ids=frozenset())
    return CaptureVolume.bootstrap(ring.image_points_noisy, ring.intrinsics_only_cameras(), constraints=cs)


def filter_case(cv_opt: CaptureVolume, percentile: float, min_per_camera: int = 10):
    """Keep-mask of filter_by_percentile_error (capture_volume.py:607-646,709-753) in
    matched-observation order, with the inputs it was computed from."""
    rep = cv_opt.reprojection_report
    raw = rep.raw_errors
    filt = cv_opt.filter_by_percentile_error(percentile, scope="per_camera", min_per_camera=min_per_camera)
    keys = ["sync_index", "cam_id", "object_id", "keypoint_id"]
    kept = set(map(tuple, filt.image_points.df[keys].itertuples(index=False, name=None)))
    mask = np.array([tuple(k) in kept for k in raw[keys].itertuples(index=False, name=None)])
    idx = cv_opt.camera_array.posed_cam_id_to_index
    thr = {}
    for cid in cv_opt.camera_array.posed_cameras:
        e = raw[raw["cam_id"] == cid]["euclidean_error"]
        thr[idx[cid]] = float(np.percentile(e, 100 - percentile)) if len(e) else np.inf
    return {
        "filt_percentile": np.float64(percentile),
        "filt_min_per_camera": np.int64(min_per_camera),
        "filt_err_xy": raw[["error_x", "error_y"]].values,
        "filt_err": raw["euclidean_error"].values,
        "filt_cam": np.array([idx[c] for c in raw["cam_id"]], np.int32),
        "filt_thresholds": np.array([thr[i] for i in range(len(thr))]),
        "filt_keep": mask,
        "filt_rmse_after": filt.reprojection_report.overall_rmse,
    }


def session4() -> CaptureVolume:
    p = REF / "tests/sessions/post_optimization"
    ca = CameraArray.from_toml(p / "camera_array.toml")
    ip = ImagePoints.from_csv(p / "calibration/extrinsic/CHARUCO/xy_CHARUCO.csv")
    wp = WorldPoints.from_csv(p / "calibration/extrinsic/CHARUCO/xyz_CHARUCO.csv")
    return CaptureVolume(ca, ip, wp)


def small_pinhole_scene():
    """tests/synthetic/test_analytic_jacobian.py:53-69."""
    from caliscope.synthetic.calibration_object import CalibrationObject
    from caliscope.synthetic.camera_synthesizer import CameraSynthesizer
    from caliscope.synthetic.synthetic_scene import SyntheticScene
    from caliscope.synthetic.trajectory import Trajectory

    camera_array = CameraSynthesizer().add_ring(n=3, radius=2.0, height=0.3).build()
    obj = CalibrationObject.planar_grid(rows=3, cols=4, spacing=0.05)
    traj = Trajectory.linear(
        n_frames=8, start=np.array([0.3, -0.3, -0.2]), end=np.array([-0.3, 0.3, 0.5]), tumble_rate=2.0, origin_frame=0
    )
    return SyntheticScene.single(
        camera_array=camera_array, calibration_object=obj, trajectory=traj, pixel_noise_sigma=0.5
    )


def mixed_fisheye_case():
    """tests/synthetic/test_analytic_jacobian.py:114-176 (fisheye 6-block + free pinhole 9-block)."""
    rvec0, tvec0 = np.array([0.1, -0.05, 0.02]), np.array([0.0, 0.1, 3.0])
    K0 = np.array([[600.0, 0, 320], [0, 590.0, 240], [0, 0, 1]])
    dist0 = np.array([0.1, -0.05, 0.01, 0.002])
    rvec1, tvec1 = np.array([-0.08, 0.12, -0.04]), np.array([0.5, -0.1, 3.2])
    K1 = np.array([[610.0, 0, 315], [0, 605.0, 245], [0, 0, 1]])
    dist1 = np.array([0.08, -0.03, 0.001, -0.002, 0.005])
    ca = CameraArray(
        {
            0: CameraData(
                cam_id=0, size=(640, 480), fisheye=True, matrix=K0, distortions=dist0,
                rotation=cv2.Rodrigues(rvec0)[0], translation=tvec0,
            ),
            1: CameraData(
                cam_id=1, size=(640, 480), matrix=K1, distortions=dist1,
                rotation=cv2.Rodrigues(rvec1)[0], translation=tvec1,
            ),
        }
    )  # fmt: skip
    rng = np.random.default_rng(42)
    points = rng.uniform(-0.6, 0.6, (25, 3))
    pts = points.reshape(-1, 1, 3)
    proj0, _ = cv2.fisheye.projectPoints(pts, rvec0.reshape(3, 1), tvec0.reshape(3, 1), K0, dist0.reshape(4, 1))
    proj1, _ = cv2.projectPoints(pts, rvec1, tvec1, K1, dist1)
    exact = np.vstack([proj0.reshape(-1, 2), proj1.reshape(-1, 2)])
    xy = exact + rng.normal(0, 0.5, exact.shape)
    cam = np.repeat(np.array([0, 1], dtype=np.int16), len(points))
    obj = np.tile(np.arange(len(points), dtype=np.int32), 2)
    par = BundleParameterization.from_camera_array(ca, n_points=len(points), refine_intrinsics=True)
    x0 = par.pack(ca, points)
    flags, const, cam_ids = blocks_to_arrays(par)
    out = {
        "cam_flags": flags, "cam_const": const, "cam_ids": cam_ids, "n_pts": np.int64(len(points)),
        "obs_cam": cam.astype(np.int32), "obs_pt": obj, "obs_xy": xy, "x0": x0,
        "r0": joint_residuals(x0, par, cam, xy, obj),
        "exact_uv": exact,
    }  # fmt: skip
    out.update(csr_parts(joint_jacobian(x0, par, cam, xy, obj)))
    # a perturbed evaluation point too (exercise non-trivial intrinsics values)
    x1 = x0 + rng.normal(0, 1e-2, x0.shape)
    out["x1"] = x1
    out["r1"] = joint_residuals(x1, par, cam, xy, obj)
    J1 = csr_parts(joint_jacobian(x1, par, cam, xy, obj))
    out.update({k.replace("J_", "J1_"): v for k, v in J1.items()})
    return out


def projection_case():
    """reprojection.project_points == cv2 for both camera models (test_reprojection_dispatch.py:13-30)."""
    rng = np.random.default_rng(7)
    pts = rng.uniform(-1.0, 1.0, (200, 3)) + np.array([0, 0, 4.0])
    rvec = np.array([0.21, -0.4, 0.13])
    tvec = np.array([0.1, -0.2, 0.5])
    K = np.array([[1394.6, 0, 960.0], [0, 1380.1, 540.0], [0, 0, 1]])
    d5 = np.array([0.115, -0.219, 0.0012, 0.0086, 0.113])
    d4 = np.array([0.05, -0.01, 0.003, -0.001])
    tiny = np.array([1e-14, -2e-14, 5e-15])
    return {
        "pts": pts, "rvec": rvec, "tvec": tvec, "K": K, "d5": d5, "d4": d4, "rvec_tiny": tiny,
        "uv_pinhole": project_points(pts, rvec, tvec, K, d5, False),
        "uv_fisheye": project_points(pts, rvec, tvec, K, d4, True),
        "uv_pinhole_tiny": project_points(pts, tiny, tvec, K, d5, False),
    }  # fmt: skip


def triangulation_case():
    """The step in front of bundle adjustment: ``triangulate_image_points`` (point_data.py:122-229) and
    ``CameraData.undistort_points`` (camera_array.py:135-174), called as the reference calls them.

    (1) session4: ``ImagePoints.triangulate(camera_array)`` end to end with the array-level call
        recorded (arguments and return value);
    (2) a seeded synthetic case with singleton groups, repeated cameras inside a group (static
        objects), unsorted rows, sparse camera ids and negative sync indices;
    (3) undistortion of a pixel grid + random points for a pinhole and a fisheye camera, both outputs.
    """
    import caliscope.core.point_data as pd_mod

    out: dict = {}
    # (1)
    p = REF / "tests/sessions/post_optimization"
    ca = CameraArray.from_toml(p / "camera_array.toml")
    ip = ImagePoints.from_csv(p / "calibration/extrinsic/CHARUCO/xy_CHARUCO.csv")
    rec: dict = {}
    orig = pd_mod.triangulate_image_points

    def recording(pm, sync, cam, obj, kp, xy):
        res = orig(pm, sync, cam, obj, kp, xy)
        rec.update(pm=pm, sync=sync, cam=cam, obj=obj, kp=kp, xy=xy, res=res)
        return res

    pd_mod.triangulate_image_points = recording
    try:
        wp = ip.triangulate(ca)
    finally:
        pd_mod.triangulate_image_points = orig
    ids = np.array(sorted(rec["pm"]), dtype=np.int64)
    out.update(
        s4_cam_ids=ids, s4_proj=np.stack([rec["pm"][int(c)] for c in ids]), s4_sync=rec["sync"], s4_cam=rec["cam"],
        s4_obj=rec["obj"], s4_kp=rec["kp"], s4_xy=rec["xy"], s4_out_sync=rec["res"][0], s4_out_obj=rec["res"][1],
        s4_out_kp=rec["res"][2], s4_out_xyz=rec["res"][3], s4_world_xyz=wp.points,
    )  # fmt: skip
    # raw pixels of the same rows and what the reference's per-camera undistortion made of them
    df = ip.df
    posed = list(ca.posed_cam_id_to_index.keys())
    df = df[df["cam_id"].isin(posed)]
    und = pd_mod._undistort_batch(df, ca)
    out.update(
        s4_px=und[["img_loc_x", "img_loc_y"]].to_numpy(np.float64), s4_px_cam=und["cam_id"].to_numpy(np.int64),
        s4_px_undist=und[["img_loc_undistort_x", "img_loc_undistort_y"]].to_numpy(np.float64),
        s4_K=np.stack([ca.cameras[int(c)].matrix for c in ids]),
        s4_dist=np.stack([np.asarray(ca.cameras[int(c)].distortions, np.float64).ravel() for c in ids]),
    )  # fmt: skip
    # (2)
    rng = np.random.default_rng(11)
    cam_ids = np.array([1, 3, 4, 8, 15, 16, 23], dtype=np.int64)
    pm = {}
    for c in cam_ids:
        R = cv2.Rodrigues(rng.normal(0, 0.4, 3))[0]
        t = np.array([0.0, 0.0, 3.0]) + rng.normal(0, 0.3, 3)
        pm[int(c)] = np.hstack([R, t[:, None]])
    rows = []
    for j in range(400):
        X = rng.uniform(-0.5, 0.5, 3)
        sync = int(j // 20) - 3
        k = int(rng.integers(1, 8))
        cams = rng.choice(cam_ids, size=k, replace=(j % 9 == 0))
        for c in cams:
            h = pm[int(c)] @ np.append(X, 1.0)
            rows.append((sync, int(c), j % 4, j % 20, *(h[:2] / h[2] + rng.normal(0, 2e-3, 2))))
    rows = np.array(rows)
    rng.shuffle(rows)
    a = (rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64), rows[:, 2].astype(np.int64), rows[:, 3].astype(np.int64),
         np.ascontiguousarray(rows[:, 4:6]))  # fmt: skip
    res = orig(pm, *a)
    out.update(
        syn_cam_ids=cam_ids, syn_proj=np.stack([pm[int(c)] for c in cam_ids]), syn_sync=a[0], syn_cam=a[1], syn_obj=a[2],
        syn_kp=a[3], syn_xy=a[4], syn_out_sync=res[0], syn_out_obj=res[1], syn_out_kp=res[2], syn_out_xyz=res[3],
    )  # fmt: skip
    # (3)
    gx, gy = np.meshgrid(np.linspace(0, 1920, 33), np.linspace(0, 1080, 19))
    pts = np.vstack([np.stack([gx.ravel(), gy.ravel()], 1), rng.uniform([0, 0], [1920, 1080], (3000, 2))])
    Kp = np.array([[1394.6, 0, 960.0], [0, 1380.1, 540.0], [0, 0, 1]])
    d5 = np.array([0.115, -0.219, 0.0012, 0.0086, 0.113])
    Kf = np.array([[600.0, 0, 955.0], [0, 605.0, 545.0], [0, 0, 1]])
    d4 = np.array([0.05, -0.01, 0.003, -0.0005])
    cp = CameraData(cam_id=0, size=(1920, 1080), matrix=Kp, distortions=d5)
    cf = CameraData(cam_id=1, size=(1920, 1080), fisheye=True, matrix=Kf, distortions=d4)
    out.update(
        und_pts=pts, und_Kp=Kp, und_d5=d5, und_Kf=Kf, und_d4=d4,
        und_pinhole_norm=cp.undistort_points(pts, output="normalized").astype(np.float64),
        und_pinhole_px=cp.undistort_points(pts, output="pixels").astype(np.float64),
        und_fisheye_norm=cf.undistort_points(pts, output="normalized").astype(np.float64),
        und_fisheye_px=cf.undistort_points(pts, output="pixels").astype(np.float64),
    )  # fmt: skip
    return out


def main() -> None:
    np.set_printoptions(precision=12)
    out_dir = HERE

    tri = triangulation_case()
    np.savez_compressed(out_dir / "triangulation.npz", **tri)
    print("triangulation", len(tri["s4_out_xyz"]), "session points,", len(tri["syn_out_xyz"]), "synthetic points")
    if "--only-triangulation" in sys.argv:
        return

    cv = session4()
    for refine in (False, True):
        g = solve_case(cv, refine)
        opt = cv.optimize(refine_intrinsics=refine)
        assert abs(opt.reprojection_report.overall_rmse - g["rmse_default"]) < 1e-12
        g["optimize_iterations"] = opt.optimization_status.iterations
        g["optimize_final_cost"] = opt.optimization_status.final_cost
        if not refine:
            g.update(filter_case(opt, 50.0))
        name = f"session4_refine{int(refine)}.npz"
        np.savez_compressed(out_dir / name, **g)
        print(name, "rmse0", g["rmse0"], "default", g["rmse_default"], "tight", g["rmse_tight"],
              "nfev", g["nfev_default"], g["nfev_tight"], "cost", g["cost_default"], g["cost_tight"])  # fmt: skip

    g = solve_case(cv, False, loss="soft_l1", f_scale=cv.pixel_f_scale(1.0), tight=True)
    np.savez_compressed(out_dir / "session4_softl1.npz", **g)
    print("session4_softl1", g["cost_default"], g["cost_tight"], g["rmse_default"], g["rmse_tight"])

    scene = small_pinhole_scene()
    cvs = CaptureVolume.bootstrap(scene.image_points_noisy, scene.intrinsics_only_cameras())
    for refine in (False, True):
        g = solve_case(cvs, refine)
        np.savez_compressed(out_dir / f"small_pinhole_refine{int(refine)}.npz", **g)
        print("small_pinhole", refine, g["rmse0"], g["rmse_default"], g["rmse_tight"], g["nfev_default"])
    groups = (
        np.array([[0, 0, 0, 0], [0, 1, 2, 3]], dtype=np.int32),
        np.array([[5, 5, 5, 5], [8, 9, 10, 11]], dtype=np.int32),
        np.array([0.11, 0.07]),
        np.array([2.0, 3.5]),
    )
    g = solve_case(cvs, False, groups=groups, tight=False)
    np.savez_compressed(out_dir / "small_pinhole_constraints.npz", **g)

    for name, vol in (("aruco_constraints", aruco_constraint_volume()), ("board_truss_constraints", board_truss_volume())):
        for refine in (False, True):
            g = solve_case(vol, refine, groups=constraint_groups(vol), tight=True)
            opt = vol.optimize(refine_intrinsics=refine)
            assert abs(opt.optimization_status.final_cost - g["cost_default"]) < 1e-12 * g["cost_default"]
            g["rigidity_rmse_mm_before"] = vol.rigidity_report().rmse_mm
            g["rigidity_rmse_mm_after"] = opt.rigidity_report().rmse_mm
            np.savez_compressed(out_dir / f"{name}_refine{int(refine)}.npz", **g)
            print(name, refine, "n_c", len(g["groups_a"]), "obs", len(g["obs_cam"]), "pts", int(g["n_pts"]), "rmse",
                  g["rmse0"], g["rmse_default"], g["rmse_tight"], "nfev", g["nfev_default"], "cost", g["cost_default"], g["cost_tight"])  # fmt: skip

    np.savez_compressed(out_dir / "mixed_fisheye.npz", **mixed_fisheye_case())
    np.savez_compressed(out_dir / "projection.npz", **projection_case())

    # default_ring_scene: exact projections of ground truth (zero residual), and the
    # reference's pinned noisy CSV (tests/fixtures/synthetic/default_ring_baseline)
    from caliscope.synthetic.scene_factories import default_ring_scene

    ring = default_ring_scene(pixel_noise_sigma=0.5, random_seed=42)
    cvr = CaptureVolume(ring.camera_array, ring.image_points_perfect, ring.world_points)
    g = solve_case(cvr, False, tight=False)
    assert np.abs(g["r0"]).max() < 1e-10
    import pandas as pd

    base = pd.read_csv(REF / "tests/fixtures/synthetic/default_ring_baseline/image_points_noisy.csv")
    dfn = ring.image_points_noisy.df.sort_values(["sync_index", "cam_id", "object_id", "keypoint_id"])
    assert np.allclose(dfn["img_loc_x"].values, base["img_loc_x"].values, atol=1e-10)
    cvn = CaptureVolume(ring.camera_array, ring.image_points_noisy, ring.world_points)
    gn = solve_case(cvn, True)
    np.savez_compressed(out_dir / "ring_perfect.npz", **g)
    np.savez_compressed(out_dir / "ring_noisy_refine1.npz", **gn)
    print("ring_noisy", gn["rmse0"], gn["rmse_default"], gn["rmse_tight"], gn["nfev_default"], gn["nfev_tight"])


if __name__ == "__main__":
    main()
