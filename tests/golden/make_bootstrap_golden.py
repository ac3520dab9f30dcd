#!/usr/bin/env python
"""Generate tests/golden/bootstrap_*.npz by running the UNMODIFIED reference's extrinsic bootstrap
(/root/reference/src/caliscope/core/bootstrap_pose/pose_network_builder.py) stage by stage on its own fixtures.

Build-container tooling (needs /root/reference, opencv, scipy, pandas); the .npz files travel, this script does not run
on the GPU box.      python tests/golden/make_bootstrap_golden.py

Stored per case: the ImagePoints columns and camera tables (inputs), then every stage's output as flat arrays:
  pnp_*    compute_camera_to_object_poses_pnp   (:211-330)  keys (cam_id, sync_index, object_id), R, t, rmse
  rel_*    compute_relative_poses               (:488-534)  keys ((a, b), sync, object), R, t
  filt_*   reject_outliers                      (:333-411)  kept count per pair
  agg_*    aggregate_poses                      (:537-575)  pair, R, t
  rmse_*   calculate_stereo_rmse_for_pair       (:638-685)  per aggregated pair (NaN where the reference returns None)
  net_*    PairedPoseNetwork.from_raw_estimates (paired_pose_network.py:26-99)  all pairs after gap filling
"""
from __future__ import annotations

import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
sys.path.insert(0, str(HERE / "_refshim"))
sys.path.insert(0, str(REF / "src"))

import numpy as np  # noqa: E402

from caliscope.cameras.camera_array import CameraArray  # noqa: E402
from caliscope.core.bootstrap_pose import pose_network_builder as PNB  # noqa: E402
from caliscope.core.bootstrap_pose.paired_pose_network import PairedPoseNetwork  # noqa: E402
from caliscope.core.bootstrap_pose.stereopairs import StereoPair  # noqa: E402
from caliscope.core.point_data import ImagePoints  # noqa: E402


def camera_tables(ca: CameraArray):
    ids = list(ca.cameras)  # dict order: the reference iterates the cameras in this order (it matters, see oracle/bootstrap.py)
    K = np.zeros((len(ids), 5))
    D = np.zeros((len(ids), 12))
    fish = np.zeros(len(ids), np.int32)
    ignore = np.zeros(len(ids), np.int32)
    for i, c in enumerate(ids):
        cam = ca.cameras[c]
        M = np.asarray(cam.matrix, float)
        K[i] = [M[0, 0], M[1, 1], M[0, 2], M[1, 2], M[0, 1]]
        d = np.asarray(cam.distortions, float).ravel()
        D[i, : len(d)] = d
        fish[i] = 1 if getattr(cam, "fisheye", False) else 0
        ignore[i] = 1 if cam.ignore else 0
    return np.array(ids, np.int32), K, D, fish, ignore


def run_case(name: str, ca: CameraArray, ip: ImagePoints) -> None:
    df = ip.df
    ids, K, D, fish, ignore = camera_tables(ca)
    out = {
        "cam_ids": ids, "cam_k": K, "cam_dist": D, "cam_fisheye": fish, "cam_ignore": ignore,
        "sync_index": df["sync_index"].to_numpy(np.int64), "cam_id": df["cam_id"].to_numpy(np.int64),
        "object_id": df["object_id"].to_numpy(np.int64), "keypoint_id": df["keypoint_id"].to_numpy(np.int64),
        "img_xy": df[["img_loc_x", "img_loc_y"]].to_numpy(np.float64),
        "obj_xyz": df[["obj_loc_x", "obj_loc_y", "obj_loc_z"]].to_numpy(np.float64),
    }  # fmt: skip
    poses = PNB.compute_camera_to_object_poses_pnp(ip, ca)
    keys = list(poses)
    out["pnp_keys"] = np.array(keys, np.int64).reshape(-1, 3)
    out["pnp_R"] = np.array([poses[k][0] for k in keys]).reshape(-1, 3, 3)
    out["pnp_t"] = np.array([poses[k][1] for k in keys]).reshape(-1, 3)
    out["pnp_rmse"] = np.array([float(poses[k][2]) for k in keys])
    rel = PNB.compute_relative_poses(poses, ca)
    rk = list(rel)
    out["rel_keys"] = np.array([[k[0][0], k[0][1], k[1], k[2]] for k in rk], np.int64).reshape(-1, 4)
    out["rel_R"] = np.array([rel[k].rotation for k in rk]).reshape(-1, 3, 3)
    out["rel_t"] = np.array([rel[k].translation for k in rk]).reshape(-1, 3)
    filt = PNB.reject_outliers(rel, threshold=1.5)
    fk = list(filt)
    out["filt_pairs"] = np.array(fk, np.int64).reshape(-1, 2)
    out["filt_count"] = np.array([len(filt[k]) for k in fk], np.int64)
    agg = PNB.aggregate_poses(filt)
    ak = list(agg)
    out["agg_pairs"] = np.array(ak, np.int64).reshape(-1, 2)
    out["agg_R"] = np.array([agg[k].rotation for k in ak]).reshape(-1, 3, 3)
    out["agg_t"] = np.array([agg[k].translation for k in ak]).reshape(-1, 3)
    common = PNB._precompute_common_observations(ip, ca)
    rm, cnt = [], []
    for k in ak:
        r = PNB.calculate_stereo_rmse_for_pair(agg[k], ca, common)
        rm.append(np.nan if r is None else r)
        cnt.append(len(common[k]) if k in common else 0)
    out["rmse_pair"] = np.array(rm)
    out["rmse_common"] = np.array(cnt, np.int64)
    net = PNB.estimate_pnp_paired_pose_network(agg, ca, ip)
    nk = list(net._pairs)
    out["net_pairs"] = np.array(nk, np.int64).reshape(-1, 2)
    out["net_R"] = np.array([net._pairs[k].rotation for k in nk]).reshape(-1, 3, 3)
    out["net_t"] = np.array([net._pairs[k].translation for k in nk]).reshape(-1, 3)
    out["net_err"] = np.array([net._pairs[k].error_score for k in nk])
    np.savez_compressed(HERE / f"bootstrap_{name}.npz", **out)
    print(name, "groups", len(keys), "relative", len(rk), "pairs", len(ak), "network", len(nk),
          "rmse", np.round(out["rmse_pair"], 6))


def session(name: str):
    p = REF / "tests/sessions" / name
    ca = CameraArray.from_toml(p / "camera_array.toml")
    ip = ImagePoints.from_csv(p / "calibration/extrinsic/CHARUCO/xy_CHARUCO.csv")
    return ca, ip


def main() -> None:
    run_case("session4", *session("post_optimization"))
    # 11 cameras / 141k observations, subsampled to every 6th sync index to keep the fixture small
    ca, ip = session("larger_calibration_post_monocal")
    df = ip.df[ip.df["sync_index"] % 6 == 0].reset_index(drop=True)
    run_case("session11", ca, ImagePoints(df))


if __name__ == "__main__":
    main()
