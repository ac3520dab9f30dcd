"""GPU-backed mirror of the reference's extrinsic bootstrap -- what produces bundle adjustment's start vector
(/root/reference/src/caliscope/core/bootstrap_pose/pose_network_builder.py).

Same function names, arguments and return types as the reference module, so that ``seam.install(full=True)`` can swap
them in (seam S4) and ``build_paired_pose_network`` / ``PoseNetworkBuilder`` run unchanged on top:

  compute_camera_to_object_poses_pnp   :211-330   one batched device call (undistortion + grouping + planar PnP)
  compute_relative_poses               :488-534   array arithmetic on the PnP output (no per-pair Python set building)
  reject_outliers                      :333-411   per-pair IQR rule, vectorised
  aggregate_poses                      :537-575   quaternion average (4x4 eigenvector) + mean translation
  estimate_pnp_paired_pose_network     :578-614   stereo RMSE of ALL pairs in one device call (:638-685)

The heavy arithmetic (undistortion, PnP per group, two-view triangulation + reprojection per common observation) runs in
CUDA through the C ABI (``cb_pnp_ippe``, ``cb_stereo_rmse``); there is no OpenCV / NumPy computation path for it here.
What stays on the host is bookkeeping on per-group / per-pair results (a few thousand 3x3 matrices).

Planar targets only (ChArUco / ArUco / chessboard: obj_loc_z constant); a non-planar group raises
``NotImplementedError`` (the reference switches to SOLVEPNP_SQPNP there).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from itertools import combinations

import numpy as np

from . import _lib as L

DEFAULT_MIN_PNP_POINTS = 4
DEFAULT_OUTLIER_THRESHOLD = 1.5

PNP_OK, PNP_TOO_FEW, PNP_NON_PLANAR, PNP_DEGENERATE, PNP_OK_FALLBACK = 0, 1, 2, 3, 4


def _ptr(a):
    return a.ctypes.data


# ---------------------------------------------------------------------------------------------------------------------
# camera tables
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class CameraTables:
    cam_ids: np.ndarray  # dict order of camera_array.cameras (the reference iterates in this order)
    index_of: dict
    k: np.ndarray  # (n, 5) fx fy cx cy skew
    dist: np.ndarray  # (n, 12)
    fisheye: np.ndarray  # (n,) int32
    ignore: np.ndarray  # (n,) bool
    has_intrinsics: np.ndarray  # (n,) bool


def camera_tables(camera_array) -> CameraTables:
    ids = list(camera_array.cameras)
    n = len(ids)
    k = np.zeros((n, 5))
    k[:, 0] = k[:, 1] = 1.0
    dist = np.zeros((n, 12))
    fish = np.zeros(n, np.int32)
    ign = np.zeros(n, bool)
    has = np.zeros(n, bool)
    for i, c in enumerate(ids):
        cam = camera_array.cameras[c]
        ign[i] = bool(getattr(cam, "ignore", False))
        if cam.matrix is None or cam.distortions is None:
            continue
        M = np.asarray(cam.matrix, dtype=np.float64)
        k[i] = [M[0, 0], M[1, 1], M[0, 2], M[1, 2], M[0, 1]]
        d = np.asarray(cam.distortions, dtype=np.float64).ravel()
        dist[i, : min(len(d), 12)] = d[:12]
        fish[i] = 1 if getattr(cam, "fisheye", False) else 0
        has[i] = True
    return CameraTables(np.asarray(ids, dtype=np.int64), {int(c): i for i, c in enumerate(ids)}, k, dist, fish, ign, has)


def _offset(values: np.ndarray):
    """(values - min, span): an order-preserving non-negative code without sorting the column (np.unique on a 2 M-row column
    costs more than the PnP kernel)."""
    values = np.asarray(values, dtype=np.int64)
    lo = int(values.min()) if len(values) else 0
    return values - lo, (int(values.max()) - lo + 1 if len(values) else 1)


def _pack(*cols) -> np.ndarray:
    """Lexicographic key of integer columns (most significant first); raises if it does not fit 63 bits.  One output array,
    updated in place (these columns have millions of rows; every temporary is a pass over memory)."""
    cols = [np.asarray(c, dtype=np.int64) for c in cols]
    los, spans = [], []
    bits = 0.0
    for c in cols:
        lo = int(c.min()) if len(c) else 0
        span = int(c.max()) - lo + 1 if len(c) else 1
        los.append(lo)
        spans.append(span)
        bits += np.log2(max(span, 1))
    if bits > 62.0:
        raise ValueError("identifier ranges too wide to pack (sync_index, object_id, keypoint_id) into a 63-bit key")
    key = cols[0] - los[0]
    for c, lo, span in zip(cols[1:], los[1:], spans[1:]):
        key *= span
        key += c
        if lo:
            key -= lo
    return np.ascontiguousarray(key)


def _slots(tab: CameraTables, cam_id: np.ndarray) -> np.ndarray:
    """camera id -> row of the camera tables, through a small lookup array."""
    lo = int(min(tab.cam_ids.min(), cam_id.min()))
    hi = int(max(tab.cam_ids.max(), cam_id.max()))
    lut = np.full(hi - lo + 1, -1, dtype=np.int32)
    lut[tab.cam_ids - lo] = np.arange(len(tab.cam_ids), dtype=np.int32)
    return lut[cam_id - lo]


# ---------------------------------------------------------------------------------------------------------------------
# stage 1: PnP per (camera, sync_index, object)
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class PnPResult:
    keys: np.ndarray  # (g, 3) cam_id, sync_index, object_id in the reference's groupby order
    R: np.ndarray  # (g, 3, 3)
    t: np.ndarray  # (g, 3)
    rmse: np.ndarray  # (g,)
    status: np.ndarray  # (g,) PNP_*
    count: np.ndarray
    group_ms: float = 0.0
    kernel_ms: float = 0.0
    launches: int = 0


def pnp_arrays(tab: CameraTables, cam_id, sync_index, object_id, img_xy, obj_xyz, min_points: int = DEFAULT_MIN_PNP_POINTS,
               device: int = 0) -> PnPResult:  # fmt: skip
    """Array-level ``compute_camera_to_object_poses_pnp``: every group of the frame in one ``cb_pnp_ippe`` call."""
    lib = L.load()
    cam_id = np.asarray(cam_id, dtype=np.int64)
    if len(cam_id) == 0:
        raise ValueError("No valid camera data found for PnP")
    slot_all = _slots(tab, cam_id)
    if tab.has_intrinsics.all() and int(slot_all.min()) >= 0:
        sel = None  # every row belongs to a calibrated camera of the array: nothing to drop (the usual case)
    else:
        sel = (slot_all >= 0) & tab.has_intrinsics[np.maximum(slot_all, 0)]
    if sel is not None and not sel.all():
        cam_id, slot_all = cam_id[sel], slot_all[sel]
        sync = np.asarray(sync_index, dtype=np.int64)[sel]
        obj_id = np.asarray(object_id, dtype=np.int64)[sel]
        px = np.ascontiguousarray(np.asarray(img_xy, dtype=np.float64).reshape(-1, 2)[sel])
        obj = np.ascontiguousarray(np.asarray(obj_xyz, dtype=np.float64).reshape(-1, 3)[sel])
    else:
        sync = np.asarray(sync_index, dtype=np.int64)
        obj_id = np.asarray(object_id, dtype=np.int64)
        px = np.ascontiguousarray(np.asarray(img_xy, dtype=np.float64).reshape(-1, 2))
        obj = np.ascontiguousarray(np.asarray(obj_xyz, dtype=np.float64).reshape(-1, 3))
    n = len(cam_id)
    if n == 0:
        raise ValueError("No valid camera data found for PnP")
    key = _pack(cam_id, sync, obj_id)
    cam_slot = np.ascontiguousarray(slot_all, dtype=np.int32)
    max_groups = n  # upper bound (one group per row); the untouched tail of the outputs costs address space only
    R = np.empty((max_groups, 3, 3))
    t = np.empty((max_groups, 3))
    rmse = np.empty(max_groups)
    status = np.empty(max_groups, np.int32)
    count = np.empty(max_groups, np.int32)
    rep = np.empty(max_groups, np.int32)
    ng = C.c_int32()
    st = L.TriStats()
    L.check(
        lib.cb_pnp_ippe(len(tab.cam_ids), _ptr(tab.fisheye), _ptr(tab.k), _ptr(tab.dist), n, _ptr(cam_slot), _ptr(key), _ptr(px),
                        _ptr(obj), int(min_points), max_groups, C.byref(ng), _ptr(R), _ptr(t), _ptr(rmse), _ptr(status),
                        _ptr(count), _ptr(rep), C.byref(st), int(device), None),
        "pnp_ippe",
    )  # fmt: skip
    g = int(ng.value)
    keys = np.stack([cam_id[rep[:g]], sync[rep[:g]], obj_id[rep[:g]]], axis=1)
    return PnPResult(keys, R[:g], t[:g], rmse[:g], status[:g], count[:g], st.group_ms, st.dlt_ms, st.kernel_launches)


def compute_camera_to_object_poses_pnp(image_points, camera_array, min_points: int = DEFAULT_MIN_PNP_POINTS,
                                       pnp_flags: int | None = None, fallback_flags: int | None = None) -> dict:  # fmt: skip
    """Drop-in for pose_network_builder.compute_camera_to_object_poses_pnp: dict (cam_id, sync_index, object_id) ->
    (R, t, rmse) in the reference's iteration order.  ``pnp_flags`` / ``fallback_flags`` other than the reference's
    defaults (IPPE / ITERATIVE) are not implemented."""
    if pnp_flags not in (None, 6) or fallback_flags not in (None, 0):  # cv2.SOLVEPNP_IPPE = 6, SOLVEPNP_ITERATIVE = 0
        raise NotImplementedError("caliscope_b200 implements the reference's default PnP flags (IPPE with the ITERATIVE fallback)")
    df = image_points.df
    tab = camera_tables(camera_array)
    res = pnp_arrays(tab, df["cam_id"].to_numpy(), df["sync_index"].to_numpy(), df["object_id"].to_numpy(),
                     df[["img_loc_x", "img_loc_y"]].to_numpy(np.float64),
                     df[["obj_loc_x", "obj_loc_y", "obj_loc_z"]].to_numpy(np.float64), min_points)  # fmt: skip
    return poses_dict(res)


def poses_dict(res: PnPResult) -> dict:
    if np.any(res.status == PNP_NON_PLANAR):
        raise NotImplementedError("non-planar PnP group (obj_loc_z spread): the reference uses SOLVEPNP_SQPNP there, "
                                  "which caliscope_b200 does not implement")  # fmt: skip
    out = {}
    for i in np.flatnonzero(res.status != PNP_TOO_FEW):  # degenerate groups stay, with a NaN pose, like cv2's
        c, s, o = (int(v) for v in res.keys[i])
        out[(c, s, o)] = (res.R[i].copy(), res.t[i].copy(), float(res.rmse[i]))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# stage 2: relative poses, outlier rejection, aggregation (arrays)
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class RelativePoses:
    pair_a: np.ndarray  # (m,) cam ids, a < b
    pair_b: np.ndarray
    sync: np.ndarray
    obj: np.ndarray
    R: np.ndarray  # (m, 3, 3)  T_B_A = T_B_obj inv(T_A_obj)
    t: np.ndarray  # (m, 3)


def relative_pose_arrays(keys: np.ndarray, R: np.ndarray, t: np.ndarray, tab: CameraTables) -> RelativePoses:
    """All (pair, sync, object) relative poses at once.  Reference quirk kept (:505-506): pairs are the combinations of the
    non-ignored cameras in DICT order, filtered by a < b -- a pair whose larger id comes first in the dict is never formed."""
    order_pos = {int(c): i for i, c in enumerate(tab.cam_ids) if not tab.ignore[i]}
    keys = np.asarray(keys, dtype=np.int64)
    live = np.array([int(c) in order_pos for c in keys[:, 0]], dtype=bool) if len(keys) else np.zeros(0, bool)
    idx = np.flatnonzero(live)
    if len(idx) == 0:
        z = np.zeros(0, np.int64)
        return RelativePoses(z, z, z, z, np.zeros((0, 3, 3)), np.zeros((0, 3)))
    k = keys[idx]
    # group rows by (sync, object)
    o = np.lexsort((k[:, 0], k[:, 2], k[:, 1]))
    k, idx = k[o], idx[o]
    brk = np.flatnonzero((np.diff(k[:, 1]) != 0) | (np.diff(k[:, 2]) != 0)) + 1
    starts = np.concatenate([[0], brk, [len(k)]])
    ia, ib = [], []
    for s, e in zip(starts[:-1], starts[1:]):
        if e - s < 2:
            continue
        loc = np.arange(s, e)
        a, b = np.triu_indices(e - s, 1)
        ia.append(loc[a])
        ib.append(loc[b])
    if not ia:
        z = np.zeros(0, np.int64)
        return RelativePoses(z, z, z, z, np.zeros((0, 3, 3)), np.zeros((0, 3)))
    ia, ib = np.concatenate(ia), np.concatenate(ib)
    ca, cb = k[ia, 0], k[ib, 0]  # ca < cb (rows sorted by camera id within a group)
    pos = np.array([order_pos[int(c)] for c in k[:, 0]])
    keep = pos[ia] < pos[ib]  # the dict-order combination (first, second) must also satisfy first < second
    ia, ib, ca, cb = ia[keep], ib[keep], ca[keep], cb[keep]
    Ra, Rb, ta, tb = R[idx[ia]], R[idx[ib]], t[idx[ia]], t[idx[ib]]
    Rrel = Rb @ np.transpose(Ra, (0, 2, 1))
    ta_inv = -np.einsum("nji,nj->ni", Ra, ta)
    trel = np.einsum("nij,nj->ni", Rb, ta_inv) + tb
    return RelativePoses(ca, cb, k[ia, 1], k[ia, 2], Rrel, trel)


def _quat_wxyz(R: np.ndarray) -> np.ndarray:
    """Unit quaternions (w, x, y, z), w >= 0 ... sign is immaterial below (outer products / averaged eigenvector)."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    q = np.empty((len(R), 4))
    # Shepperd's method: pick the largest of (trace, m00, m11, m22) for stability
    choice = np.argmax(np.stack([m00 + m11 + m22, m00, m11, m22], axis=1), axis=1)
    for c in range(4):
        s = choice == c
        if not s.any():
            continue
        r = R[s]
        if c == 0:
            w = np.sqrt(np.maximum(1 + r[:, 0, 0] + r[:, 1, 1] + r[:, 2, 2], 0)) / 2
            q[s] = np.stack([w, (r[:, 2, 1] - r[:, 1, 2]) / (4 * w), (r[:, 0, 2] - r[:, 2, 0]) / (4 * w), (r[:, 1, 0] - r[:, 0, 1]) / (4 * w)], axis=1)
        elif c == 1:
            x = np.sqrt(np.maximum(1 + r[:, 0, 0] - r[:, 1, 1] - r[:, 2, 2], 0)) / 2
            q[s] = np.stack([(r[:, 2, 1] - r[:, 1, 2]) / (4 * x), x, (r[:, 0, 1] + r[:, 1, 0]) / (4 * x), (r[:, 0, 2] + r[:, 2, 0]) / (4 * x)], axis=1)
        elif c == 2:
            y = np.sqrt(np.maximum(1 - r[:, 0, 0] + r[:, 1, 1] - r[:, 2, 2], 0)) / 2
            q[s] = np.stack([(r[:, 0, 2] - r[:, 2, 0]) / (4 * y), (r[:, 0, 1] + r[:, 1, 0]) / (4 * y), y, (r[:, 1, 2] + r[:, 2, 1]) / (4 * y)], axis=1)
        else:
            z = np.sqrt(np.maximum(1 - r[:, 0, 0] - r[:, 1, 1] + r[:, 2, 2], 0)) / 2
            q[s] = np.stack([(r[:, 1, 0] - r[:, 0, 1]) / (4 * z), (r[:, 0, 2] + r[:, 2, 0]) / (4 * z), (r[:, 1, 2] + r[:, 2, 1]) / (4 * z), z], axis=1)
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _quat_to_matrix(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])  # fmt: skip


def quaternion_average(quaternions: np.ndarray) -> np.ndarray:
    """pose_network_builder.py:414-437: eigenvector of sum q q^T for the largest eigenvalue, w >= 0."""
    quaternions = np.asarray(quaternions, dtype=np.float64)
    if len(quaternions) == 0:
        raise ValueError("Cannot average empty quaternion array")
    if len(quaternions) == 1:
        return quaternions[0]
    _, V = np.linalg.eigh(quaternions.T @ quaternions)
    q = V[:, -1]
    if q[0] < 0:
        q = -q
    nrm = np.linalg.norm(q)
    return quaternions[0] if nrm < 1e-10 else q / nrm


def _segment_percentiles(values: np.ndarray, seg: np.ndarray, n_seg: int, qs=(25.0, 75.0)):
    """np.percentile(values[seg == p], qs) for every segment p at once (default 'linear' rule, reproduced with NumPy's own
    interpolation formula so the numbers are the ones the reference's per-pair np.percentile calls give)."""
    from .filtering import _numpy_linear_interp

    order = np.lexsort((values, seg))
    v = values[order]
    cnt = np.bincount(seg, minlength=n_seg).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    out = []
    for q in qs:
        vi = (cnt - 1).astype(np.float64) * (q / 100.0)
        lo = np.floor(np.maximum(vi, 0)).astype(np.int64)
        hi = np.minimum(lo + 1, np.maximum(cnt - 1, 0))
        has = cnt > 0
        a = np.zeros(n_seg)
        b = np.zeros(n_seg)
        a[has] = v[(start + lo)[has]]
        b[has] = v[(start + hi)[has]]
        out.append(_numpy_linear_interp(a, b, vi - np.floor(vi)))
    return out


def _segment_quaternion_average(quats: np.ndarray, seg: np.ndarray, n_seg: int) -> np.ndarray:
    """quaternion_average of every segment at once: eigenvector of sum q q^T for the largest eigenvalue, w >= 0.
    `seg` must be sorted ascending (rows of a segment contiguous) and every segment non-empty."""
    outer = quats[:, :, None] * quats[:, None, :]
    starts = np.flatnonzero(np.concatenate([[True], np.diff(seg) != 0]))
    M = np.add.reduceat(outer, starts, axis=0)
    assert len(M) == n_seg
    _, V = np.linalg.eigh(M)
    q = V[:, :, -1]
    q = np.where(q[:, :1] < 0, -q, q)
    nrm = np.linalg.norm(q, axis=1, keepdims=True)
    single = np.bincount(seg, minlength=n_seg) == 1
    q = q / np.where(nrm < 1e-10, 1.0, nrm)
    if single.any():  # the reference returns the lone quaternion itself
        q[single] = quats[starts[single]]
    return q


def _quats_to_matrices(q: np.ndarray) -> np.ndarray:
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def filter_and_aggregate(rel: RelativePoses, threshold: float = DEFAULT_OUTLIER_THRESHOLD,
                         rotation_threshold_multiplier: float | None = None,
                         translation_threshold_multiplier: float | None = None):  # fmt: skip
    """``reject_outliers`` + ``aggregate_poses`` on arrays, every camera pair at once (no per-pair Python loop: percentiles
    by one two-key sort, quaternion averages by one batched 4x4 eigen-decomposition).  Returns (pairs (p, 2), kept mask over
    the rows of ``rel``, aggregated R (p, 3, 3), t (p, 3), kept count (p,)), pairs in ascending (a, b) order."""
    rot_m = rotation_threshold_multiplier if rotation_threshold_multiplier is not None else threshold
    tr_m = translation_threshold_multiplier if translation_threshold_multiplier is not None else threshold
    m = len(rel.pair_a)
    keep = np.zeros(m, bool)
    empty = (np.zeros((0, 2), np.int64), keep, np.zeros((0, 3, 3)), np.zeros((0, 3)), np.zeros(0, np.int64))
    if m == 0:
        return empty
    finite = np.isfinite(rel.R).all(axis=(1, 2)) & np.isfinite(rel.t).all(axis=1)  # the reference's NaN filter (:364-367)
    rows = np.flatnonzero(finite)
    if len(rows) == 0:
        return empty
    span = int(rel.pair_b.max()) + 1
    pk = rel.pair_a[rows] * span + rel.pair_b[rows]
    o = np.argsort(pk, kind="stable")
    rows, pk = rows[o], pk[o]
    uniq, seg = np.unique(pk, return_inverse=True)  # seg ascending: rows of a pair contiguous
    n_seg = len(uniq)
    cnt = np.bincount(seg, minlength=n_seg)
    R, t = rel.R[rows], rel.t[rows]
    quats = _quat_wxyz(R)
    tmag = np.linalg.norm(t, axis=1)
    # IQR rule on pairs with at least 5 valid samples (:369-372)
    big = cnt >= 5
    ok = np.ones(len(rows), bool)
    if big.any():
        t_q1, t_q3 = _segment_percentiles(tmag, seg, n_seg)
        t_lo, t_hi = t_q1 - tr_m * (t_q3 - t_q1), t_q3 + tr_m * (t_q3 - t_q1)
        Rm = _quats_to_matrices(_segment_quaternion_average(quats, seg, n_seg))
        tr = np.clip(np.einsum("nij,nij->n", R, Rm[seg]), -1.0, 3.0)  # trace(R Rm^T)
        ang = np.degrees(np.arccos((tr - 1) / 2))
        r_q1, r_q3 = _segment_percentiles(ang, seg, n_seg)
        r_hi = r_q3 + rot_m * (r_q3 - r_q1)
        bad = (tmag < t_lo[seg]) | (tmag > t_hi[seg]) | (ang > r_hi[seg])
        ok = ~(bad & big[seg])
    keep[rows[ok]] = True
    rows, seg2, R, t, quats = rows[ok], seg[ok], R[ok], t[ok], quats[ok]
    if len(rows) == 0:
        return empty
    uniq2, seg3 = np.unique(seg2, return_inverse=True)
    n2 = len(uniq2)
    cnt2 = np.bincount(seg3, minlength=n2)
    starts = np.flatnonzero(np.concatenate([[True], np.diff(seg3) != 0]))
    R_agg = _quats_to_matrices(_segment_quaternion_average(quats, seg3, n2))
    t_agg = np.add.reduceat(t, starts, axis=0) / cnt2[:, None]
    one = cnt2 == 1
    if one.any():  # a single survivor is passed through untouched (:551-553)
        R_agg[one] = R[starts[one]]
        t_agg[one] = t[starts[one]]
    pk2 = uniq[uniq2]
    pairs = np.stack([pk2 // span, pk2 % span], axis=1).astype(np.int64)
    return pairs, keep, R_agg, t_agg, cnt2.astype(np.int64)


def pose_network_arrays(keys: np.ndarray, R: np.ndarray, t: np.ndarray, tab: CameraTables,
                        threshold: float = DEFAULT_OUTLIER_THRESHOLD, rotation_threshold_multiplier: float | None = None,
                        translation_threshold_multiplier: float | None = None, want_keep: bool = False, device: int = 0):  # fmt: skip
    """``relative_pose_arrays`` + ``filter_and_aggregate`` in ONE device call (``cb_relative_pose_network``): every
    (pair, sync, object) relative pose, the IQR rule and the quaternion / translation averages of every camera pair.
    Returns (pairs (p, 2) ascending, R (p, 3, 3), t (p, 3), kept count (p,), keep mask over the rows ``relative_pose_arrays``
    would return -- or None -- , stats).  The relative poses themselves never come back to the host."""
    lib = L.load()
    rot_m = rotation_threshold_multiplier if rotation_threshold_multiplier is not None else threshold
    tr_m = translation_threshold_multiplier if translation_threshold_multiplier is not None else threshold
    keys = np.asarray(keys, dtype=np.int64).reshape(-1, 3)
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    t = np.asarray(t, dtype=np.float64).reshape(-1, 3)
    empty = (np.zeros((0, 2), np.int64), np.zeros((0, 3, 3)), np.zeros((0, 3)), np.zeros(0, np.int64),
             np.zeros(0, bool) if want_keep else None, L.TriStats())
    if len(keys) == 0:
        return empty
    pos_of = np.full(int(max(tab.cam_ids.max(), keys[:, 0].max())) + 1, -1, np.int64)
    live_ids = tab.cam_ids[~tab.ignore]
    pos_of[live_ids] = np.flatnonzero(~tab.ignore)
    if keys[:, 0].min() < 0:
        raise ValueError("negative camera id")
    idx = np.flatnonzero(pos_of[keys[:, 0]] >= 0)  # poses of cameras the array ignores (or does not know) form no pair
    if len(idx) == 0:
        return empty
    k = keys[idx]
    o = np.lexsort((k[:, 0], k[:, 2], k[:, 1]))  # (sync, object) groups, camera id ascending inside
    k, idx = k[o], idx[o]
    brk = np.flatnonzero((np.diff(k[:, 1]) != 0) | (np.diff(k[:, 2]) != 0)) + 1
    frame_start = np.concatenate([[0], brk, [len(k)]]).astype(np.int32)
    sizes = np.diff(frame_start).astype(np.int64)
    n_rel = int((sizes * (sizes - 1) // 2).sum())
    if n_rel == 0:
        return empty
    cam_id = np.ascontiguousarray(k[:, 0], dtype=np.int32)
    cam_pos = np.ascontiguousarray(pos_of[k[:, 0]], dtype=np.int32)
    Rg = np.ascontiguousarray(R[idx])
    tg = np.ascontiguousarray(t[idx])
    n_ids = len(np.unique(cam_id))
    max_pairs = n_ids * (n_ids - 1) // 2
    pa = np.empty(max(max_pairs, 1), np.int32)
    pb = np.empty(max(max_pairs, 1), np.int32)
    Ro = np.empty((max(max_pairs, 1), 3, 3))
    to = np.empty((max(max_pairs, 1), 3))
    cnt = np.empty(max(max_pairs, 1), np.int64)
    n_out = C.c_int32()
    valid = np.empty(n_rel, np.uint8) if want_keep else None
    keep = np.empty(n_rel, np.uint8) if want_keep else None
    st = L.TriStats()
    L.check(
        lib.cb_relative_pose_network(len(k), len(frame_start) - 1, _ptr(frame_start), _ptr(cam_id), _ptr(cam_pos), _ptr(Rg), _ptr(tg),
                                     float(rot_m), float(tr_m), int(max_pairs), C.byref(n_out), _ptr(pa), _ptr(pb), _ptr(Ro), _ptr(to),
                                     _ptr(cnt), n_rel if want_keep else 0, _ptr(valid) if want_keep else None,
                                     _ptr(keep) if want_keep else None, C.byref(st), int(device), None),
        "relative_pose_network",
    )  # fmt: skip
    n = n_out.value
    pairs = np.stack([pa[:n], pb[:n]], axis=1).astype(np.int64)
    keep_rows = keep[valid != 0].astype(bool) if want_keep else None
    return pairs, Ro[:n].copy(), to[:n].copy(), cnt[:n].copy(), keep_rows, st


# ---------------------------------------------------------------------------------------------------------------------
# stage 3: stereo RMSE of every pair
# ---------------------------------------------------------------------------------------------------------------------
def stereo_rmse_arrays(tab: CameraTables, pairs: np.ndarray, R: np.ndarray, t: np.ndarray, cam_id, sync_index, object_id,
                       keypoint_id, img_xy, min_common: int = DEFAULT_MIN_PNP_POINTS, device: int = 0):  # fmt: skip
    """(rmse (p,), n_common (p,)) for pairs (a < b) with pose [R | t]; NaN where the reference returns None -- fewer than
    ``min_common`` common observations, or (reference quirk, :589-598 with :655) the pair's cameras appear in the other
    order in the camera dict, so that its common observations are stored under (b, a) and never found."""
    lib = L.load()
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    p = len(pairs)
    rmse = np.full(p, np.nan)
    cnt = np.zeros(p, np.int64)
    if p == 0:
        return rmse, cnt
    cam_id = np.asarray(cam_id, dtype=np.int64)
    slot = _slots(tab, cam_id)
    m = slot >= 0
    sync_index, object_id, keypoint_id = (np.asarray(a, dtype=np.int64) for a in (sync_index, object_id, keypoint_id))
    px = np.asarray(img_xy, dtype=np.float64).reshape(-1, 2)
    if not m.all():
        slot, sync_index, object_id, keypoint_id, px = slot[m], sync_index[m], object_id[m], keypoint_id[m], px[m]
    key = _pack(sync_index, object_id, keypoint_id)
    slot = np.ascontiguousarray(slot, dtype=np.int32)
    px = np.ascontiguousarray(px)
    # device pairs are (slot_lo, slot_hi); a pose given for (a, b) must be inverted when slot[a] > slot[b]
    sa = np.array([tab.index_of[int(a)] for a in pairs[:, 0]], dtype=np.int32)
    sb = np.array([tab.index_of[int(b)] for b in pairs[:, 1]], dtype=np.int32)
    usable = (sa < sb) & ~tab.ignore[sa] & ~tab.ignore[sb]  # dict order == id order for this pair, neither camera ignored
    if not usable.any():
        return rmse, cnt
    Rt = np.ascontiguousarray(np.concatenate([np.asarray(R, np.float64).reshape(p, 9), np.asarray(t, np.float64).reshape(p, 3)], axis=1)[usable])
    pa, pb = np.ascontiguousarray(sa[usable]), np.ascontiguousarray(sb[usable])
    out_r = np.empty(int(usable.sum()))
    out_c = np.empty(int(usable.sum()), np.int64)
    st = L.TriStats()
    L.check(
        lib.cb_stereo_rmse(len(tab.cam_ids), _ptr(tab.fisheye), _ptr(tab.k), _ptr(tab.dist), len(pa), _ptr(pa), _ptr(pb), _ptr(Rt),
                           len(key), _ptr(slot), _ptr(key), _ptr(px), int(min_common), _ptr(out_r), _ptr(out_c), C.byref(st),
                           int(device), None),
        "stereo_rmse",
    )  # fmt: skip
    rmse[usable] = out_r
    cnt[usable] = out_c
    return rmse, cnt


# ---------------------------------------------------------------------------------------------------------------------
# drop-ins with the reference's container types
# ---------------------------------------------------------------------------------------------------------------------
def _stereo_pair_cls():
    try:
        from caliscope.core.bootstrap_pose.stereopairs import StereoPair  # the reference's own class when importable

        return StereoPair
    except Exception:
        return StereoPairLite


@dataclass(frozen=True)
class StereoPairLite:
    """Field-for-field stand-in for the reference's StereoPair (stereopairs.py:14-50) when it is not importable."""

    primary_cam_id: int
    secondary_cam_id: int
    error_score: float
    translation: np.ndarray
    rotation: np.ndarray

    @property
    def pair(self):
        return (self.primary_cam_id, self.secondary_cam_id)


def compute_relative_poses(camera_to_object_poses: dict, camera_array) -> dict:
    """Drop-in for pose_network_builder.compute_relative_poses: dict ((a, b), sync_index, object_id) -> StereoPair."""
    SP = _stereo_pair_cls()
    tab = camera_tables(camera_array)
    keys = np.array(list(camera_to_object_poses), dtype=np.int64).reshape(-1, 3)
    R = np.array([v[0] for v in camera_to_object_poses.values()]).reshape(-1, 3, 3)
    t = np.array([v[1] for v in camera_to_object_poses.values()]).reshape(-1, 3)
    rel = relative_pose_arrays(keys, R, t, tab)
    # the reference's dict is ordered by pair (combination order), then set-iteration order inside a pair; consumers only
    # group by pair, so pair-major order is what is reproduced
    pos = {int(c): i for i, c in enumerate(tab.cam_ids)}
    order = np.lexsort((rel.obj, rel.sync, [pos[int(b)] for b in rel.pair_b], [pos[int(a)] for a in rel.pair_a]))
    out = {}
    for i in order:
        a, b = int(rel.pair_a[i]), int(rel.pair_b[i])
        out[((a, b), int(rel.sync[i]), int(rel.obj[i]))] = SP(primary_cam_id=a, secondary_cam_id=b, error_score=float("nan"),
                                                               translation=rel.t[i].copy(), rotation=rel.R[i].copy())  # fmt: skip
    return out


def _rel_from_dict(relative_poses: dict) -> RelativePoses:
    ks = list(relative_poses)
    vals = list(relative_poses.values())
    return RelativePoses(np.array([k[0][0] for k in ks], np.int64), np.array([k[0][1] for k in ks], np.int64),
                         np.array([k[1] for k in ks], np.int64), np.array([k[2] for k in ks], np.int64),
                         np.array([v.rotation for v in vals]).reshape(-1, 3, 3), np.array([v.translation for v in vals]).reshape(-1, 3))  # fmt: skip


def reject_outliers(relative_poses: dict, threshold: float = DEFAULT_OUTLIER_THRESHOLD,
                    rotation_threshold_multiplier: float | None = None,
                    translation_threshold_multiplier: float | None = None) -> dict:  # fmt: skip
    """Drop-in for pose_network_builder.reject_outliers: dict pair -> list of StereoPair that pass."""
    rel = _rel_from_dict(relative_poses)
    _, keep, _, _, _ = filter_and_aggregate(rel, threshold, rotation_threshold_multiplier, translation_threshold_multiplier)
    out: dict = {}
    for (pair, _s, _o), sp in relative_poses.items():
        out.setdefault(pair, [])
    for i, ((pair, _s, _o), sp) in enumerate(relative_poses.items()):
        if keep[i]:
            out[pair].append(sp)
    return out


def aggregate_poses(filtered_poses: dict) -> dict:
    """Drop-in for pose_network_builder.aggregate_poses."""
    SP = _stereo_pair_cls()
    out = {}
    for pair, lst in filtered_poses.items():
        if not lst:
            continue
        if len(lst) == 1:
            out[pair] = lst[0]
            continue
        q = quaternion_average(_quat_wxyz(np.array([sp.rotation for sp in lst])))
        out[pair] = SP(primary_cam_id=pair[0], secondary_cam_id=pair[1], error_score=float("nan"), rotation=_quat_to_matrix(q),
                       translation=np.mean([sp.translation for sp in lst], axis=0))  # fmt: skip
    return out


def stereo_rmse_for_pairs(aggregated_pairs: dict, camera_array, image_points) -> dict:
    """pair -> RMSE or None, every pair in one device call (== calculate_stereo_rmse_for_pair per pair)."""
    tab = camera_tables(camera_array)
    pairs = np.array(list(aggregated_pairs), dtype=np.int64).reshape(-1, 2)
    R = np.array([sp.rotation for sp in aggregated_pairs.values()]).reshape(-1, 3, 3)
    t = np.array([sp.translation for sp in aggregated_pairs.values()]).reshape(-1, 3)
    df = image_points.df
    rmse, _ = stereo_rmse_arrays(tab, pairs, R, t, df["cam_id"].to_numpy(), df["sync_index"].to_numpy(), df["object_id"].to_numpy(),
                                 df["keypoint_id"].to_numpy(), df[["img_loc_x", "img_loc_y"]].to_numpy(np.float64))  # fmt: skip
    return {tuple(int(v) for v in p): (None if np.isnan(r) else float(r)) for p, r in zip(pairs, rmse)}


def build_paired_pose_network(image_points, camera_array, min_points: int = DEFAULT_MIN_PNP_POINTS,
                              threshold: float = DEFAULT_OUTLIER_THRESHOLD):
    """Drop-in for ``caliscope.core.bootstrap_pose.build_paired_pose_network.build_paired_pose_network`` (:14-31): the PnP
    branch -- ``PoseNetworkBuilder(...).estimate_camera_to_object_poses().estimate_relative_poses().filter_outliers(1.5)
    .build()`` -- as three device calls on arrays (PnP of every group, relative-pose network, stereo RMSE of every pair) with
    no per-pose Python objects in between; only the <= n_cams^2 aggregated pairs become ``StereoPair`` objects for the
    reference's own ``PairedPoseNetwork.from_raw_estimates`` (gap filling: graph bookkeeping, stays the reference's).  With
    2D-only observations (obj_loc all NaN) the reference's epipolar bootstrap, which this package does not replace, is called."""
    from caliscope.core.bootstrap_pose.paired_pose_network import PairedPoseNetwork

    df = image_points.df
    obj = df[["obj_loc_x", "obj_loc_y", "obj_loc_z"]].to_numpy(np.float64)
    if np.isnan(obj).all():
        from caliscope.core.bootstrap_pose.epipolar_pose_builder import build_epipolar_pose_network

        return build_epipolar_pose_network(image_points, camera_array)
    SP = _stereo_pair_cls()
    tab = camera_tables(camera_array)
    cam_id, sync, oid, kp = (df[c].to_numpy() for c in ("cam_id", "sync_index", "object_id", "keypoint_id"))
    xy = df[["img_loc_x", "img_loc_y"]].to_numpy(np.float64)
    res = pnp_arrays(tab, cam_id, sync, oid, xy, obj, min_points)
    if np.any(res.status == PNP_NON_PLANAR):
        raise NotImplementedError("non-planar PnP group (obj_loc_z spread): the reference uses SOLVEPNP_SQPNP there, "
                                  "which caliscope_b200 does not implement")  # fmt: skip
    live = res.status != PNP_TOO_FEW
    pairs, R, t, _cnt, _, _ = pose_network_arrays(res.keys[live], res.R[live], res.t[live], tab, threshold)
    rmse, _ = stereo_rmse_arrays(tab, pairs, R, t, cam_id, sync, oid, kp, xy)
    # the reference's dict order: combinations of the cameras in dict order
    order = np.lexsort(([tab.index_of[int(b)] for b in pairs[:, 1]], [tab.index_of[int(a)] for a in pairs[:, 0]])) if len(pairs) else []
    with_rmse = {}
    for i in order:
        if np.isnan(rmse[i]):
            continue
        a, b = int(pairs[i, 0]), int(pairs[i, 1])
        with_rmse[(a, b)] = SP(primary_cam_id=a, secondary_cam_id=b, error_score=float(rmse[i]), rotation=R[i].copy(),
                               translation=t[i].copy())  # fmt: skip
    return PairedPoseNetwork.from_raw_estimates(with_rmse)


def estimate_pnp_paired_pose_network(aggregated_pairs_wo_rmse: dict, camera_array, image_points):
    """Drop-in for pose_network_builder.estimate_pnp_paired_pose_network (needs the reference's PairedPoseNetwork for the
    gap filling, paired_pose_network.py:26-99, which is graph bookkeeping on <= n_cams^2 pairs and stays the reference's)."""
    from caliscope.core.bootstrap_pose.paired_pose_network import PairedPoseNetwork

    SP = _stereo_pair_cls()
    rm = stereo_rmse_for_pairs(aggregated_pairs_wo_rmse, camera_array, image_points)
    with_rmse = {}
    for pair, sp in aggregated_pairs_wo_rmse.items():
        if rm.get(pair) is None:
            continue
        with_rmse[pair] = SP(primary_cam_id=sp.primary_cam_id, secondary_cam_id=sp.secondary_cam_id, error_score=rm[pair],
                             rotation=sp.rotation, translation=sp.translation)  # fmt: skip
    return PairedPoseNetwork.from_raw_estimates(with_rmse)
