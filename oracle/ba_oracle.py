"""NumPy restatement of Caliscope's bundle-adjustment residual / Jacobian path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Parity status: PINNED.
``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` (in the build container) and stores its outputs for the
reference's own fixtures as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them.

What is restated (reference file:line, relative to /root/reference):

* parameter layout         src/caliscope/core/bundle_parameterization.py:114-186
* ``project_points``       src/caliscope/core/reprojection.py:18-32 -> cv2.projectPoints /
                           cv2.fisheye.projectPoints (OpenCV 4.x, un-vendored third-party;
                           closed forms below follow OpenCV's documented camera model)
* ``joint_residuals``      src/caliscope/core/reprojection.py:75-119
* ``joint_jacobian``       src/caliscope/core/reprojection.py:128-234
* ``reprojection_errors``  src/caliscope/core/reprojection.py:35-72
* robust loss rescaling    scipy/optimize/_lsq/least_squares.py (loss functions) and
                           scipy/optimize/_lsq/common.py:720-731 (un-vendored third-party,
                           scipy 1.18.1 installed)
* the solver call          src/caliscope/core/capture_volume.py:387-411

The restatement is written array-at-a-time over ALL observations (the reference
loops over cameras and calls OpenCV per camera); the arithmetic per observation
is the same, so outputs agree to rounding (checked at 1e-9 px / 1e-10 relative).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.sparse import csr_matrix

FLAG_FREE_INTRINSICS = 1
FLAG_FISHEYE = 2


@dataclass
class Rig:
    """Array-level description of one bundle-adjustment problem.

    Mirrors what ``CaptureVolume.optimize`` hands to scipy
    (capture_volume.py:353-365) plus the per-camera constants that
    ``BundleParameterization.blocks`` carries (bundle_parameterization.py:36-51).

    cam_const columns: fx0, fy0, cx, cy, c4, c5, c6, c7, c8 where for a
    Brown-Conrady camera (c4..c8) = (k1_initial, k2_initial, p1, p2, k3) and for
    a fisheye camera (c4..c7) = (k1, k2, k3, k4), c8 unused.
    """

    cam_flags: np.ndarray  # (n_cams,) int32  bit0 free intrinsics, bit1 fisheye
    cam_const: np.ndarray  # (n_cams, 9) float64
    n_pts: int
    obs_cam: np.ndarray  # (n_obs,) int32
    obs_pt: np.ndarray  # (n_obs,) int32
    obs_xy: np.ndarray  # (n_obs, 2) float64
    # optional rigid-distance rows (reprojection.py:112-117)
    groups_a: np.ndarray | None = None  # (n_c, 4) int32
    groups_b: np.ndarray | None = None
    distances: np.ndarray | None = None
    weights: np.ndarray | None = None
    cam_offsets: np.ndarray = field(init=False)

    def __post_init__(self) -> None:
        self.cam_flags = np.ascontiguousarray(self.cam_flags, dtype=np.int32)
        self.cam_const = np.ascontiguousarray(self.cam_const, dtype=np.float64).reshape(-1, 9)
        self.obs_cam = np.ascontiguousarray(self.obs_cam, dtype=np.int32)
        self.obs_pt = np.ascontiguousarray(self.obs_pt, dtype=np.int32)
        self.obs_xy = np.ascontiguousarray(self.obs_xy, dtype=np.float64).reshape(-1, 2)
        widths = np.where(self.cam_flags & FLAG_FREE_INTRINSICS, 9, 6)
        self.cam_offsets = np.concatenate([[0], np.cumsum(widths)]).astype(np.int32)

    @property
    def n_cams(self) -> int:
        return len(self.cam_flags)

    @property
    def n_obs(self) -> int:
        return len(self.obs_cam)

    @property
    def n_camera_params(self) -> int:
        return int(self.cam_offsets[-1])

    @property
    def n_params(self) -> int:
        return self.n_camera_params + 3 * self.n_pts

    @property
    def n_constraints(self) -> int:
        return 0 if self.groups_a is None else len(self.groups_a)

    def bounds(self) -> tuple[np.ndarray, np.ndarray]:
        """bundle_parameterization.py:151-164."""
        lo = np.full(self.n_params, -np.inf)
        hi = np.full(self.n_params, np.inf)
        for i in np.nonzero(self.cam_flags & FLAG_FREE_INTRINSICS)[0]:
            o = self.cam_offsets[i] + 6
            lo[o : o + 3] = (0.5, -1.0, -2.0)
            hi[o : o + 3] = (2.0, 1.0, 2.0)
        return lo, hi


# ----------------------------------------------------------------------------
# rotation helpers
# ----------------------------------------------------------------------------


def rodrigues(r: np.ndarray) -> np.ndarray:
    """(n,3) rotation vectors -> (n,3,3) matrices (cv2.Rodrigues convention)."""
    r = np.atleast_2d(np.asarray(r, dtype=np.float64))
    th = np.linalg.norm(r, axis=1)
    small = th < 1e-12
    k = r / np.where(small, 1.0, th)[:, None]
    c, s = np.cos(th), np.sin(th)
    K = _skew(k)
    kk = k[:, :, None] * k[:, None, :]
    eye = np.eye(3)[None]
    R = c[:, None, None] * eye + (1 - c)[:, None, None] * kk + s[:, None, None] * K
    if small.any():
        R[small] = eye + _skew(r[small])
    return R


def _skew(v: np.ndarray) -> np.ndarray:
    z = np.zeros(len(v))
    return np.stack(
        [
            np.stack([z, -v[:, 2], v[:, 1]], axis=1),
            np.stack([v[:, 2], z, -v[:, 0]], axis=1),
            np.stack([-v[:, 1], v[:, 0], z], axis=1),
        ],
        axis=1,
    )


def so3_right_jacobian(r: np.ndarray) -> np.ndarray:
    """Jr(r) with d(R(r) X)/dr = -R [X]x Jr(r).  (n,3) -> (n,3,3)."""
    r = np.atleast_2d(np.asarray(r, dtype=np.float64))
    th2 = np.sum(r * r, axis=1)
    th = np.sqrt(th2)
    small = th < 1e-4
    ths = np.where(small, 1.0, th)
    B = np.where(small, 0.5 - th2 / 24.0 + th2 * th2 / 720.0, (1 - np.cos(ths)) / (ths * ths))
    C = np.where(small, 1.0 / 6.0 - th2 / 120.0 + th2 * th2 / 5040.0, (ths - np.sin(ths)) / (ths**3))
    K = _skew(r)
    return np.eye(3)[None] - B[:, None, None] * K + C[:, None, None] * (K @ K)


# ----------------------------------------------------------------------------
# camera parameter expansion (bundle_parameterization.py:166-186)
# ----------------------------------------------------------------------------


def _expand_cameras(x: np.ndarray, rig: Rig, *, stored_intrinsics: bool = False):
    """Per-camera rvec, tvec, fx, fy, cx, cy, dist coefficients (5 slots)."""
    offs = rig.cam_offsets[:-1]
    cc = rig.cam_const
    free = (rig.cam_flags & FLAG_FREE_INTRINSICS) != 0
    rvec = np.stack([x[offs + i] for i in range(3)], axis=1)
    tvec = np.stack([x[offs + 3 + i] for i in range(3)], axis=1)
    s = np.ones(rig.n_cams)
    k1 = cc[:, 4].copy()
    k2 = cc[:, 5].copy()
    if free.any():
        fo = offs[free]
        s[free] = x[fo + 6]
        k1[free] = x[fo + 7]
        k2[free] = x[fo + 8]
    fx = s * cc[:, 0]
    fy = s * cc[:, 1]
    return rvec, tvec, fx, fy, cc[:, 2], cc[:, 3], k1, k2, cc[:, 6], cc[:, 7], cc[:, 8]


def _project(x: np.ndarray, rig: Rig, want_jac: bool):
    """Projection (+ analytic blocks) for every observation.

    Returns uv (n_obs,2) and, if want_jac, (J_rvec, J_tvec, J_s, J_k1, J_k2, J_X)
    in PIXEL units (not yet divided by fx0), each (n_obs, 2, k).
    """
    rvec, tvec, fx, fy, cx, cy, k1, k2, d2, d3, d4 = _expand_cameras(x, rig)
    R = rodrigues(rvec)
    pts = x[rig.n_camera_params :].reshape(-1, 3)
    ci, pj = rig.obs_cam, rig.obs_pt
    X = pts[pj]
    Ro = R[ci]
    Xc = np.einsum("nij,nj->ni", Ro, X) + tvec[ci]
    zc = Xc[:, 2]
    iz = np.where(zc != 0.0, 1.0 / np.where(zc != 0.0, zc, 1.0), 1.0)  # OpenCV: z==0 -> 1
    a = Xc[:, 0] * iz
    b = Xc[:, 1] * iz
    fxo, fyo = fx[ci], fy[ci]
    fish = ((rig.cam_flags & FLAG_FISHEYE) != 0)[ci]

    # ---- Brown-Conrady [k1 k2 p1 p2 k3]
    K1, K2, P1, P2, K3 = k1[ci], k2[ci], d2[ci], d3[ci], d4[ci]
    r2 = a * a + b * b
    cd = 1 + r2 * (K1 + r2 * (K2 + r2 * K3))
    xd = a * cd + 2 * P1 * a * b + P2 * (r2 + 2 * a * a)
    yd = b * cd + P1 * (r2 + 2 * b * b) + 2 * P2 * a * b

    # ---- fisheye equidistant [k1 k2 k3 k4]  (coefficients live in slots 4..7)
    if fish.any():
        F1, F2, F3, F4 = k1[ci], k2[ci], d2[ci], d3[ci]
        rr = np.sqrt(r2)
        th = np.arctan(rr)
        th2 = th * th
        thd = th * (1 + th2 * (F1 + th2 * (F2 + th2 * (F3 + th2 * F4))))
        big = rr > 1e-8
        inv_r = np.where(big, 1.0 / np.where(big, rr, 1.0), 1.0)
        cdist = np.where(big, thd * inv_r, 1.0)
        xd = np.where(fish, a * cdist, xd)
        yd = np.where(fish, b * cdist, yd)

    uv = np.stack([fxo * xd + cx[ci], fyo * yd + cy[ci]], axis=1)
    if not want_jac:
        return uv, None

    # d(xd,yd)/d(a,b)
    dcd = K1 + r2 * (2 * K2 + 3 * K3 * r2)
    xa = cd + 2 * a * a * dcd + 2 * P1 * b + 6 * P2 * a
    xb = 2 * a * b * dcd + 2 * P1 * a + 2 * P2 * b
    ya = xb
    yb = cd + 2 * b * b * dcd + 6 * P1 * b + 2 * P2 * a
    if fish.any():
        dthd = 1 + th2 * (3 * F1 + th2 * (5 * F2 + th2 * (7 * F3 + th2 * 9 * F4)))
        # cdist = thd(atan r)/r ; d cdist / d r
        dcdr = np.where(big, (dthd / (1 + r2) - cdist) * inv_r, 0.0)
        # d r / d a = a / r
        fa = np.where(big, a * inv_r, 0.0)
        fb = np.where(big, b * inv_r, 0.0)
        xa = np.where(fish, cdist + a * dcdr * fa, xa)
        xb = np.where(fish, a * dcdr * fb, xb)
        ya = np.where(fish, b * dcdr * fa, ya)
        yb = np.where(fish, cdist + b * dcdr * fb, yb)

    # d(a,b)/dXc = [[iz,0,-a iz],[0,iz,-b iz]]
    Jt = np.empty((rig.n_obs, 2, 3))
    Jt[:, 0, 0] = fxo * xa * iz
    Jt[:, 0, 1] = fxo * xb * iz
    Jt[:, 0, 2] = -fxo * (xa * a + xb * b) * iz
    Jt[:, 1, 0] = fyo * ya * iz
    Jt[:, 1, 1] = fyo * yb * iz
    Jt[:, 1, 2] = -fyo * (ya * a + yb * b) * iz

    JX = Jt @ Ro
    # d(R X)/dr = -R [X]x Jr  ->  J_r = -(J_X [X]x) Jr ; row_i(J_X [X]x) = J_X,i x X
    Jr_so3 = so3_right_jacobian(rvec)[ci]
    JXx = np.cross(JX, X[:, None, :])
    Jrv = -(JXx @ Jr_so3)

    cc = rig.cam_const
    Js = np.stack([cc[ci, 0] * xd, cc[ci, 1] * yd], axis=1)[:, :, None]
    Jk1 = np.stack([fxo * a * r2, fyo * b * r2], axis=1)[:, :, None]
    Jk2 = Jk1 * r2[:, None, None]
    return uv, (Jrv, Jt, Js, Jk1, Jk2, JX)


# ----------------------------------------------------------------------------
# public restatements
# ----------------------------------------------------------------------------


def project_points(world, rvec, tvec, K, dist, fisheye: bool) -> np.ndarray:
    """reprojection.py:18-32 for one camera (used only to pin against cv2)."""
    world = np.asarray(world, dtype=np.float64).reshape(-1, 3)
    d = np.asarray(dist, dtype=np.float64).ravel()
    if fisheye:
        if d.shape[0] != 4:
            raise ValueError(f"Fisheye projection requires 4 distortion coefficients, got {d.shape[0]}")
        const = [K[0][0], K[1][1], K[0][2], K[1][2], d[0], d[1], d[2], d[3], 0.0]
        flags = FLAG_FISHEYE
    else:
        d5 = np.zeros(5)
        d5[: min(5, len(d))] = d[:5]
        const = [K[0][0], K[1][1], K[0][2], K[1][2], *d5]
        flags = 0
    n = len(world)
    rig = Rig(
        cam_flags=np.array([flags]),
        cam_const=np.array([const]),
        n_pts=n,
        obs_cam=np.zeros(n, np.int32),
        obs_pt=np.arange(n, dtype=np.int32),
        obs_xy=np.zeros((n, 2)),
    )
    x = np.concatenate([np.ravel(rvec), np.ravel(tvec), world.ravel()]).astype(np.float64)
    return _project(x, rig, False)[0]


def constraint_residuals(x: np.ndarray, rig: Rig) -> np.ndarray:
    pts = x[rig.n_camera_params :].reshape(-1, 3)
    ea = pts[rig.groups_a].mean(axis=1)
    eb = pts[rig.groups_b].mean(axis=1)
    return (np.linalg.norm(ea - eb, axis=1) - rig.distances) * rig.weights


def residuals(x: np.ndarray, rig: Rig) -> np.ndarray:
    """joint_residuals (reprojection.py:75-119): interleaved (x,y)/fx0, caller's row order."""
    x = np.asarray(x, dtype=np.float64)
    uv, _ = _project(x, rig, False)
    r = ((uv - rig.obs_xy) / rig.cam_const[rig.obs_cam, 0:1]).ravel()
    if rig.n_constraints:
        r = np.concatenate([r, constraint_residuals(x, rig)])
    return r


def jacobian_blocks(x: np.ndarray, rig: Rig) -> tuple[np.ndarray, np.ndarray]:
    """Per-observation dense blocks of joint_jacobian (reprojection.py:171-205).

    Returns Jc (n_obs, 2, 9) -- columns [rvec3, tvec3, s, k1, k2], the last three
    zero for locked cameras -- and Jp (n_obs, 2, 3); both already divided by fx0.
    """
    x = np.asarray(x, dtype=np.float64)
    _, (Jrv, Jt, Js, Jk1, Jk2, JX) = _project(x, rig, True)
    inv = 1.0 / rig.cam_const[rig.obs_cam, 0]
    free = ((rig.cam_flags & FLAG_FREE_INTRINSICS) != 0)[rig.obs_cam]
    Jc = np.concatenate([Jrv, Jt, Js, Jk1, Jk2], axis=2) * inv[:, None, None]
    Jc[~free, :, 6:] = 0.0
    return Jc, JX * inv[:, None, None]


def jacobian(x: np.ndarray, rig: Rig) -> csr_matrix:
    """joint_jacobian (reprojection.py:128-234) as CSR, built directly (rows are
    already in order, so no COO sort is needed)."""
    Jc, Jp = jacobian_blocks(x, rig)
    n_obs = rig.n_obs
    widths = (rig.cam_offsets[1:] - rig.cam_offsets[:-1])[rig.obs_cam]  # 6 or 9
    per_row = widths + 3
    ncp = rig.n_camera_params
    # two rows per observation
    row_nnz = np.repeat(per_row, 2)
    n_c = rig.n_constraints
    data_parts = []
    col_parts = []
    for w in (6, 9):
        sel = np.nonzero(widths == w)[0]
        if len(sel) == 0:
            continue
        cols_c = rig.cam_offsets[rig.obs_cam[sel]][:, None] + np.arange(w)[None]
        cols_p = ncp + 3 * rig.obs_pt[sel].astype(np.int64)[:, None] + np.arange(3)[None]
        cols = np.concatenate([cols_c, cols_p], axis=1)  # (k, w+3)
        vals = np.concatenate([Jc[sel][:, :, :w], Jp[sel]], axis=2)  # (k, 2, w+3)
        data_parts.append((sel, vals, np.repeat(cols[:, None, :], 2, axis=1)))
    indptr = np.concatenate([[0], np.cumsum(row_nnz)]).astype(np.int64)
    nnz_obs = int(indptr[-1])
    data = np.empty(nnz_obs)
    indices = np.empty(nnz_obs, dtype=np.int64)
    for sel, vals, cols in data_parts:
        w3 = vals.shape[2]
        for half in (0, 1):
            starts = indptr[2 * sel + half]
            pos = starts[:, None] + np.arange(w3)[None]
            data[pos] = vals[:, half, :]
            indices[pos] = cols[:, half, :]
    J = csr_matrix((data, indices, indptr), shape=(2 * n_obs, rig.n_params))
    if n_c:
        J = _append_constraint_rows(J, x, rig)
    return J


def _append_constraint_rows(J: csr_matrix, x: np.ndarray, rig: Rig) -> csr_matrix:
    """reprojection.py:207-226: +/- 1/4 w unit per group column, duplicates summed."""
    from scipy.sparse import coo_matrix, vstack

    pts = x[rig.n_camera_params :].reshape(-1, 3)
    diffs = pts[rig.groups_a].mean(axis=1) - pts[rig.groups_b].mean(axis=1)
    norms = np.linalg.norm(diffs, axis=1)
    unit = diffs / np.where(norms > 0, norms, 1.0)[:, None]
    n_c = rig.n_constraints
    rows, cols, vals = [], [], []
    for groups, sign in ((rig.groups_a, 1.0), (rig.groups_b, -1.0)):
        contrib = (sign * 0.25) * rig.weights[:, None] * unit
        for q in range(4):
            base = rig.n_camera_params + 3 * groups[:, q].astype(np.int64)
            for k in range(3):
                rows.append(np.arange(n_c))
                cols.append(base + k)
                vals.append(contrib[:, k])
    C = coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_c, rig.n_params)
    ).tocsr()
    return vstack([J, C]).tocsr()


def reproj_errors_px(x: np.ndarray, rig: Rig) -> np.ndarray:
    """reprojection_errors (reprojection.py:35-72): pixel errors, stored intrinsics.

    ``x`` is the optimised vector, so "stored intrinsics" are the ones
    ``unpack_into`` (bundle_parameterization.py:138-149) writes back.
    """
    uv, _ = _project(np.asarray(x, dtype=np.float64), rig, False)
    return uv - rig.obs_xy


def overall_rmse_px(x: np.ndarray, rig: Rig) -> float:
    """capture_volume.py:183,197."""
    e = reproj_errors_px(x, rig)
    return float(np.sqrt(np.mean(np.sum(e * e, axis=1))))


# ----------------------------------------------------------------------------
# robust losses (scipy least_squares.py:183-252, common.py:720-731)
# ----------------------------------------------------------------------------

LOSSES = ("linear", "soft_l1", "huber", "cauchy", "arctan")


def loss_rho(z: np.ndarray, loss: str) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """rho(z), rho'(z), rho''(z) for z = (f/f_scale)^2."""
    if loss == "linear":
        return z.copy(), np.ones_like(z), np.zeros_like(z)
    if loss == "soft_l1":
        t = 1 + z
        return 2 * (t**0.5 - 1), t**-0.5, -0.5 * t**-1.5
    if loss == "huber":
        m = z <= 1
        zs = np.where(m, 1.0, z)
        return (
            np.where(m, z, 2 * zs**0.5 - 1),
            np.where(m, 1.0, zs**-0.5),
            np.where(m, 0.0, -0.5 * zs**-1.5),
        )
    if loss == "cauchy":
        t = 1 + z
        return np.log1p(z), 1 / t, -1 / t**2
    if loss == "arctan":
        t = 1 + z * z
        return np.arctan(z), 1 / t, -2 * z / t**2
    raise ValueError(loss)


def robust_cost(f: np.ndarray, loss: str, f_scale: float) -> float:
    if loss == "linear":
        return 0.5 * float(f @ f)
    z = (f / f_scale) ** 2
    return 0.5 * f_scale**2 * float(np.sum(loss_rho(z, loss)[0]))


def robust_row_scales(f: np.ndarray, loss: str, f_scale: float) -> tuple[np.ndarray, np.ndarray]:
    """Per-row (J_scale, f_scaled) exactly as scale_for_robust_loss_function."""
    if loss == "linear":
        return np.ones_like(f), f.copy()
    z = (f / f_scale) ** 2
    _, r1, r2 = loss_rho(z, loss)
    js = r1 + 2 * r2 * z  # rho2/f_scale^2 * f^2 = rho2 * z
    js = np.sqrt(np.maximum(js, np.finfo(float).eps))
    return js, f * r1 / js


# ----------------------------------------------------------------------------
# the reference solver call, verbatim in meaning (capture_volume.py:387-411)
# ----------------------------------------------------------------------------


def solve_scipy(
    rig: Rig,
    x0: np.ndarray,
    *,
    ftol: float = 1e-8,
    xtol: float = 1e-8,
    gtol: float = 1e-8,
    max_nfev: int | None = None,
    loss: str = "linear",
    f_scale: float = 1.0,
    verbose: int = 0,
):
    """scipy.optimize.least_squares(method='trf', x_scale='jac', jac=<sparse analytic>)."""
    from scipy.optimize import least_squares

    nit = [0]

    def _cb(intermediate_result):
        nit[0] = int(intermediate_result.nit)

    res = least_squares(
        residuals,
        np.asarray(x0, dtype=np.float64),
        args=(rig,),
        jac=jacobian,
        verbose=verbose,
        x_scale="jac",
        loss=loss,
        f_scale=f_scale,
        ftol=ftol,
        xtol=xtol,
        gtol=gtol,
        max_nfev=max_nfev,
        method="trf",
        bounds=rig.bounds(),
        callback=_cb,
    )
    res.nit = nit[0]
    return res
