"""CPU restatement of the planar PnP the reference's bootstrap uses.  TEST INFRASTRUCTURE ONLY.

Reference call site: /root/reference/src/caliscope/core/bootstrap_pose/pose_network_builder.py:301-306 --
``cv2.solvePnP(obj_points[f32], img_points[f32, undistorted normalised], K = I, D = 0, flags = SOLVEPNP_IPPE)`` per
(camera, sync_index, object) group, then ``cv2.Rodrigues`` and a reprojection RMSE in the normalised plane (:314-321).

The arithmetic lives in an un-vendored third-party dependency: OpenCV (spec ``opencv-python>=4.8.0.74``,
/root/reference/pyproject.toml:14; installed here: opencv-python-headless 4.13.0.92).  Published algorithms restated:

  * IPPE -- T. Collins, A. Bartoli, "Infinitesimal Plane-Based Pose Estimation", IJCV 109(3), 2014: the two poses
    consistent with the first-order behaviour of the model-to-image homography at the model centroid;
  * the homography by M. Harker, P. O'Leary, "Computation of Homographies", BMVC 2005 (isotropic normalisation,
    elimination of the affine part, smallest eigenvector of the 3x3 reduced system) -- the estimator OpenCV's IPPE uses,
    which matters: with noisy points a different estimator gives a different homography, hence a different pose.

Pinned by tests/golden/bootstrap.npz: poses and errors produced by cv2.solvePnP itself on the session fixture and on
synthetic boards (tests/golden/make_bootstrap_golden.py), compared in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np


def _normalize_isotropic(pts: np.ndarray):
    """2 x n points -> zero mean, mean distance sqrt(2); returns (normalised, T, Tinv)."""
    n = pts.shape[1]
    m = pts.mean(axis=1)
    d = pts - m[:, None]
    kappa = np.sum(d[0] ** 2 + d[1] ** 2)
    beta = np.sqrt(2.0 * n / kappa)
    T = np.array([[beta, 0.0, -beta * m[0]], [0.0, beta, -beta * m[1]], [0.0, 0.0, 1.0]])
    Ti = np.array([[1.0 / beta, 0.0, m[0]], [0.0, 1.0 / beta, m[1]], [0.0, 0.0, 1.0]])
    return d * beta, T, Ti


def homography_ho(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """Harker-O'Leary homography dst ~ H src; src, dst: 2 x n."""
    A, TA, TAi = _normalize_isotropic(src)
    B, TB, TBi = _normalize_isotropic(dst)
    n = A.shape[1]
    C1, C2, C3, C4 = -B[0] * A[0], -B[0] * A[1], -B[1] * A[0], -B[1] * A[1]
    mC1, mC2, mC3, mC4 = C1.mean(), C2.mean(), C3.mean(), C4.mean()
    Mx = np.stack([C1 - mC1, C2 - mC2, -B[0]], axis=1)
    My = np.stack([C3 - mC3, C4 - mC4, -B[1]], axis=1)
    AAt = A @ A.T
    Pp = np.linalg.inv(AAt) @ A  # 2 x n
    Bx, By = Pp @ Mx, Pp @ My  # 2 x 3
    Ex, Ey = A.T @ Bx, A.T @ By
    D = np.concatenate([Mx - Ex, My - Ey], axis=0)  # 2n x 3
    w, V = np.linalg.eigh(D.T @ D)
    h789 = V[:, 0]  # smallest eigenvalue
    h12 = -Bx @ h789
    h45 = -By @ h789
    h3 = -(mC1 * h789[0] + mC2 * h789[1])
    h6 = -(mC3 * h789[0] + mC4 * h789[1])
    H = np.array([[h12[0], h12[1], h3], [h45[0], h45[1], h6], [h789[0], h789[1], h789[2]]])
    H = TBi @ H @ TA
    return H / H[2, 2]


def _rotate_vec_to_z(a: np.ndarray) -> np.ndarray:
    """Rotation taking the direction of `a` onto +z (minimal rotation)."""
    ax, ay, az = a / np.linalg.norm(a)
    if az < 0 and abs(1.0 + az) < 1e-12:
        return np.diag([1.0, -1.0, -1.0])
    d = 1.0 / (1.0 + az)
    return np.array([[1.0 - ax * ax * d, -ax * ay * d, -ax], [-ax * ay * d, 1.0 - ay * ay * d, -ay], [ax, ay, 1.0 - (ax * ax + ay * ay) * d]])


def _rotations_from_homography(H: np.ndarray, want_gamma: bool = False):
    p, q = H[0, 2], H[1, 2]
    J = np.array([[H[0, 0] - H[2, 0] * p, H[0, 1] - H[2, 1] * p], [H[1, 0] - H[2, 0] * q, H[1, 1] - H[2, 1] * q]])
    Rv = _rotate_vec_to_z(np.array([p, q, 1.0])).T  # takes z onto the viewing ray
    Bm = np.array([[Rv[0, 0] - p * Rv[2, 0], Rv[0, 1] - p * Rv[2, 1]], [Rv[1, 0] - q * Rv[2, 0], Rv[1, 1] - q * Rv[2, 1]]])
    A = np.linalg.inv(Bm) @ J
    AtA = A @ A.T
    gamma = np.sqrt(0.5 * (AtA[0, 0] + AtA[1, 1] + np.sqrt((AtA[0, 0] - AtA[1, 1]) ** 2 + 4.0 * AtA[0, 1] ** 2)))
    Rt = A / gamma
    b0 = np.sqrt(max(0.0, 1.0 - Rt[0, 0] ** 2 - Rt[1, 0] ** 2))
    b1 = np.sqrt(max(0.0, 1.0 - Rt[0, 1] ** 2 - Rt[1, 1] ** 2))
    if -Rt[0, 0] * Rt[0, 1] - Rt[1, 0] * Rt[1, 1] < 0:
        b1 = -b1
    out = []
    for s in (1.0, -1.0):
        c0 = np.array([Rt[0, 0], Rt[1, 0], s * b0])
        c1 = np.array([Rt[0, 1], Rt[1, 1], s * b1])
        out.append(Rv @ np.stack([c0, c1, np.cross(c0, c1)], axis=1))
    return (out, gamma) if want_gamma else out


def _translation(obj2: np.ndarray, img: np.ndarray, R: np.ndarray) -> np.ndarray:
    """Least-squares t for points (x, y, 0) rotated by R and seen at normalised image points img (2 x n)."""
    n = obj2.shape[1]
    r = R[:, :2] @ obj2  # 3 x n
    u, v = img[0], img[1]
    ATA = np.array([[n, 0.0, -u.sum()], [0.0, n, -v.sum()], [-u.sum(), -v.sum(), np.sum(u * u + v * v)]])
    bx, by = u * r[2] - r[0], v * r[2] - r[1]
    ATb = np.array([bx.sum(), by.sum(), -np.sum(u * bx + v * by)])
    return np.linalg.solve(ATA, ATb)


def _reproj_sq(obj2, img, R, t) -> float:
    Xc = R[:, :2] @ obj2 + t[:, None]
    e = img - Xc[:2] / Xc[2]
    return float(np.sum(e * e))


def solve_pnp_ippe(obj_points: np.ndarray, img_points: np.ndarray):
    """(R (3,3), t (3,), second R, second t): the better and the worse IPPE pose of a PLANAR target (z constant).
    obj_points (n,3), img_points (n,2) normalised; both are used at float32 precision like the reference's call."""
    obj = np.asarray(obj_points, dtype=np.float32).astype(np.float64)
    obj[:, 2] = np.nan_to_num(obj[:, 2], nan=0.0)
    img = np.asarray(img_points, dtype=np.float32).astype(np.float64).T
    mean = obj.mean(axis=0)
    obj2 = (obj - mean).T[:2]
    nanR, nant = np.full((3, 3), np.nan), np.full(3, np.nan)
    try:
        with np.errstate(all="ignore"):
            H = homography_ho(obj2, img)
            sols = []
            for R in _rotations_from_homography(H):
                t = _translation(obj2, img, R)
                sols.append((_reproj_sq(obj2, img, R, t), R, t))
    except np.linalg.LinAlgError:
        sols = []
    if len(sols) != 2 or not all(np.all(np.isfinite(s[1])) and np.all(np.isfinite(s[2])) for s in sols):
        # collinear model points: no homography.  cv2.solvePnP reports success with a NaN pose there (the reference keeps
        # the group and drops it later with its NaN filter, pose_network_builder.py:364-367)
        return nanR, nant, nanR, nant
    sols.sort(key=lambda s: s[0])
    out = []
    for _, R, t in sols:
        out.append(R)
        out.append(t - R @ mean)  # canonical (centred) frame -> the caller's object frame
    return tuple(out)


def pnp_reprojection_rmse(obj_points, img_points, R, t) -> float:
    """pose_network_builder.py:317-318 -- sqrt(mean over points of the squared normalised-plane distance)."""
    obj = np.asarray(obj_points, dtype=np.float32).astype(np.float64)
    img = np.asarray(img_points, dtype=np.float32).astype(np.float64)
    obj[:, 2] = np.nan_to_num(obj[:, 2], nan=0.0)
    Xc = obj @ R.T + t
    e = img - Xc[:, :2] / Xc[:, 2:3]
    return float(np.sqrt(np.mean(np.sum(e * e, axis=1))))


# ----------------------------------------------------------------------------------------------------------------
# Fallback of the reference for groups where IPPE fails (pose_network_builder.py:308-311: SOLVEPNP_ITERATIVE).  OpenCV's
# IPPE gives up when the homography is degenerate (three of four points collinear, four collinear + one, ...): the
# first-order scale `gamma` of the homography at the centroid collapses to ~1e-10.  cv2's ITERATIVE is a Levenberg-
# Marquardt minimisation of the reprojection error from a homography-based start.  Restated as: the two IPPE poses of
# the AFFINE fit (which stays well defined as long as the points are not all collinear), each refined by Gauss-Newton
# on the reprojection error, best one returned.  Agreement with cv2 is at the optimiser's tolerance (1e-6), and only
# if both land in the same local minimum -- such groups are ill-conditioned by construction.
# ----------------------------------------------------------------------------------------------------------------
IPPE_GAMMA_MIN = 1e-7


def _rodrigues(r):
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3) + np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def refine_pose(obj: np.ndarray, img: np.ndarray, R: np.ndarray, t: np.ndarray, iters: int = 30):
    """Gauss-Newton on sum |img - proj(R X + t)|^2 with a left-multiplied rotation increment."""
    for _ in range(iters):
        Xc = obj @ R.T + t
        z = Xc[:, 2]
        u, v = Xc[:, 0] / z, Xc[:, 1] / z
        r = np.concatenate([u - img[:, 0], v - img[:, 1]])
        n = len(obj)
        J = np.zeros((2 * n, 6))
        Xr = Xc - t  # R X
        for i in range(n):
            du = np.array([1 / z[i], 0, -u[i] / z[i]])
            dv = np.array([0, 1 / z[i], -v[i] / z[i]])
            S = -np.array([[0, -Xr[i, 2], Xr[i, 1]], [Xr[i, 2], 0, -Xr[i, 0]], [-Xr[i, 1], Xr[i, 0], 0]])  # d(exp(w) R X)/dw
            J[i, :3], J[i, 3:] = du @ S, du
            J[n + i, :3], J[n + i, 3:] = dv @ S, dv
        H = J.T @ J
        g = J.T @ r
        try:
            d = -np.linalg.solve(H + 1e-12 * np.trace(H) * np.eye(6), g)
        except np.linalg.LinAlgError:
            break
        R = _rodrigues(d[:3]) @ R
        t = t + d[3:]
        if np.linalg.norm(d) < 1e-14:
            break
    return R, t


def solve_pnp_planar(obj_points, img_points):
    """The reference's planar-group behaviour: IPPE, or the fallback when IPPE's homography is degenerate.
    Returns (R, t, used_fallback)."""
    obj = np.asarray(obj_points, dtype=np.float32).astype(np.float64)
    obj[:, 2] = np.nan_to_num(obj[:, 2], nan=0.0)
    img = np.asarray(img_points, dtype=np.float32).astype(np.float64)
    mean = obj.mean(axis=0)
    obj2 = (obj - mean).T[:2]
    imT = img.T
    with np.errstate(all="ignore"):
        try:
            H = homography_ho(obj2, imT)
            _, gamma = _rotations_from_homography(H, want_gamma=True)  # first-order scale of H at the centroid
        except np.linalg.LinAlgError:
            gamma = np.nan
    if np.isfinite(gamma) and gamma >= IPPE_GAMMA_MIN:
        R, t, _, _ = solve_pnp_ippe(obj_points, img_points)
        return R, t, False
    # affine fit img ~ M obj2 + c
    AAt = obj2 @ obj2.T
    if not np.isfinite(gamma) or abs(np.linalg.det(AAt)) < 1e-12 * max(np.trace(AAt), 1e-300) ** 2:
        return np.full((3, 3), np.nan), np.full(3, np.nan), False  # all points collinear: cv2 reports a NaN pose
    c = imT.mean(axis=1)
    M = (imT - c[:, None]) @ obj2.T @ np.linalg.inv(AAt)
    Haff = np.array([[M[0, 0], M[0, 1], c[0]], [M[1, 0], M[1, 1], c[1]], [0.0, 0.0, 1.0]])
    best = None
    for R in _rotations_from_homography(Haff):
        t = _translation(obj2, imT, R)
        R2, t2 = refine_pose(np.stack([obj2[0], obj2[1], np.zeros(obj2.shape[1])], axis=1), img, R, t)
        e = _reproj_sq(obj2, imT, R2, t2)
        if best is None or e < best[0]:
            best = (e, R2, t2)
    _, R, t = best
    return R, t - R @ mean, True
