"""Host-side mirror of the reference's parameter-vector contract for bundle adjustment.

Same public names, argument meaning and error behaviour as
/root/reference/src/caliscope/core/bundle_parameterization.py (CameraBlock :36-51,
BundleParameterization :54-279), so array-level callers and tests read the same against either
class.  Layout: one block per posed camera in ``posed_index_to_cam_id`` order,
``[rvec(3), tvec(3)]`` plus ``[s, k1, k2]`` when the intrinsics are free (fx = s*fx_initial,
fy = s*fy_initial), then xyz per world point.  Fisheye blocks are always 6 wide.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import cached_property
from typing import Any

import numpy as np

N_EXTRINSIC_PARAMS = 6
N_FREE_INTRINSIC_PARAMS = 3

# (lower, upper) for the free intrinsic slots s, k1, k2
_FREE_BOUNDS = ((0.5, 2.0), (-1.0, 1.0), (-2.0, 2.0))


class CalibrationError(Exception):
    """Same role as caliscope.exceptions.CalibrationError (raised when a camera cannot be optimised)."""


@dataclass(frozen=True)
class BoundWarning:
    cam_id: int
    parameter: str  # "f" | "k1" | "k2"
    bound: str  # "lower" | "upper"
    value: float


@dataclass(frozen=True)
class IntrinsicEstimate:
    cam_id: int
    f_recovered: float
    k1_recovered: float
    k2_recovered: float
    f_initial: float
    k1_initial: float
    k2_initial: float


@dataclass(frozen=True)
class CameraBlock:
    cam_id: int
    free_intrinsics: bool
    fx_initial: float
    fy_initial: float
    cx: float
    cy: float
    fisheye: bool
    dist_fixed: tuple[float, ...]
    k1_initial: float = 0.0
    k2_initial: float = 0.0

    @property
    def n_params(self) -> int:
        return N_EXTRINSIC_PARAMS + (N_FREE_INTRINSIC_PARAMS if self.free_intrinsics else 0)


def _calibration_error(msg: str) -> Exception:
    try:  # raise the reference's own exception type when it is importable
        from caliscope.exceptions import CalibrationError as RefError  # type: ignore

        return RefError(msg)
    except Exception:
        return CalibrationError(msg)


def _block_for(cam_id: int, cam: Any, refine_intrinsics: bool) -> CameraBlock:
    if cam.matrix is None or cam.distortions is None:
        raise _calibration_error(
            f"Camera {cam_id} has no intrinsics. Run intrinsic calibration or synthesize defaults before optimizing."
        )
    K = np.asarray(cam.matrix, dtype=np.float64)
    d = np.asarray(cam.distortions, dtype=np.float64).ravel()
    common = dict(cam_id=cam_id, fx_initial=float(K[0, 0]), fy_initial=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
    if getattr(cam, "fisheye", False):
        if d.size != 4:
            raise _calibration_error(
                f"Fisheye camera {cam_id} requires exactly 4 distortion coefficients (equidistant model), got {d.size}."
            )
        return CameraBlock(free_intrinsics=False, fisheye=True, dist_fixed=tuple(d), **common)
    return CameraBlock(
        free_intrinsics=bool(refine_intrinsics),
        fisheye=False,
        dist_fixed=tuple(d[2:5]),
        k1_initial=float(d[0]),
        k2_initial=float(d[1]),
        **common,
    )


@dataclass(frozen=True)
class BundleParameterization:
    blocks: tuple[CameraBlock, ...]
    n_points: int

    @classmethod
    def from_camera_array(cls, camera_array: Any, n_points: int, *, refine_intrinsics: bool) -> "BundleParameterization":
        index_to_id = camera_array.posed_index_to_cam_id
        blocks = tuple(
            _block_for(index_to_id[i], camera_array.cameras[index_to_id[i]], refine_intrinsics) for i in sorted(index_to_id)
        )
        return cls(blocks=blocks, n_points=n_points)

    # ---- layout -------------------------------------------------------------------------------
    @cached_property
    def camera_param_offsets(self) -> tuple[int, ...]:
        widths = [b.n_params for b in self.blocks]
        return tuple(int(v) for v in np.concatenate([[0], np.cumsum(widths)[:-1]])) if widths else ()

    @cached_property
    def n_camera_params(self) -> int:
        return int(sum(b.n_params for b in self.blocks))

    def _free_slots(self):
        for off, b in zip(self.camera_param_offsets, self.blocks):
            if b.free_intrinsics:
                yield off + N_EXTRINSIC_PARAMS, b

    def pack(self, camera_array: Any, world_points_xyz) -> np.ndarray:
        x = np.empty(self.n_camera_params + 3 * self.n_points)
        for off, b in zip(self.camera_param_offsets, self.blocks):
            cam = camera_array.cameras[b.cam_id]
            x[off : off + 6] = cam.extrinsics_to_vector()
            if b.free_intrinsics:
                d = np.asarray(cam.distortions, dtype=np.float64).ravel()
                x[off + 6 : off + 9] = (1.0, d[0], d[1])
        x[self.n_camera_params :] = np.asarray(world_points_xyz, dtype=np.float64).ravel()
        return x

    def unpack_into(self, camera_array: Any, x) -> np.ndarray:
        x = np.asarray(x)
        for off, b in zip(self.camera_param_offsets, self.blocks):
            cam = camera_array.cameras[b.cam_id]
            cam.extrinsics_from_vector(x[off : off + 6])
            if b.free_intrinsics:
                s, k1, k2 = x[off + 6 : off + 9]
                cam.matrix = np.array([[s * b.fx_initial, 0.0, b.cx], [0.0, s * b.fy_initial, b.cy], [0.0, 0.0, 1.0]])
                cam.distortions = np.array([k1, k2, *b.dist_fixed])
        return x[self.n_camera_params :].reshape(-1, 3)

    def bounds(self) -> tuple[np.ndarray, np.ndarray]:
        n = self.n_camera_params + 3 * self.n_points
        lower, upper = np.full(n, -np.inf), np.full(n, np.inf)
        for slot, _ in self._free_slots():
            for k, (lo, hi) in enumerate(_FREE_BOUNDS):
                lower[slot + k], upper[slot + k] = lo, hi
        return lower, upper

    def trial_projection_inputs(self, x, block_index: int):
        b = self.blocks[block_index]
        off = self.camera_param_offsets[block_index]
        s, k1, k2 = (x[off + 6], x[off + 7], x[off + 8]) if b.free_intrinsics else (1.0, b.k1_initial, b.k2_initial)
        K = np.array([[s * b.fx_initial, 0.0, b.cx], [0.0, s * b.fy_initial, b.cy], [0.0, 0.0, 1.0]])
        dist = np.array(b.dist_fixed) if b.fisheye else np.array([k1, k2, *b.dist_fixed])
        return x[off : off + 3], x[off + 3 : off + 6], K, dist

    def sparsity(self, camera_indices, obj_indices, n_constraints, constraint_groups_a, constraint_groups_b):
        """Structural non-zeros of the Jacobian (same shape and meaning as the reference's lil_matrix)."""
        from scipy.sparse import coo_matrix

        cam = np.asarray(camera_indices, dtype=np.int64)
        pt = np.asarray(obj_indices, dtype=np.int64)
        n_obs = len(cam)
        offs = np.asarray(self.camera_param_offsets, dtype=np.int64)
        widths = np.asarray([b.n_params for b in self.blocks], dtype=np.int64)
        rows, cols = [], []
        for w in np.unique(widths[cam]) if n_obs else []:
            sel = np.nonzero(widths[cam] == w)[0]
            c = (offs[cam[sel]][:, None] + np.arange(w)[None]).ravel()
            for half in (0, 1):
                rows.append(np.repeat(2 * sel + half, w))
                cols.append(c)
        pc = (self.n_camera_params + 3 * pt[:, None] + np.arange(3)[None]).ravel()
        for half in (0, 1):
            rows.append(np.repeat(2 * np.arange(n_obs) + half, 3))
            cols.append(pc)
        if constraint_groups_a is not None and constraint_groups_b is not None and n_constraints > 0:
            for groups in (np.asarray(constraint_groups_a), np.asarray(constraint_groups_b)):
                gc = (self.n_camera_params + 3 * groups.astype(np.int64)[:, :, None] + np.arange(3)[None, None]).reshape(
                    n_constraints, -1
                )
                rows.append(np.repeat(2 * n_obs + np.arange(n_constraints), gc.shape[1]))
                cols.append(gc.ravel())
        r = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        c = np.concatenate(cols) if cols else np.zeros(0, np.int64)
        m = coo_matrix((np.ones(len(r), dtype=int), (r, c)), shape=(2 * n_obs + n_constraints, self.n_camera_params + 3 * self.n_points))
        out = m.tolil()
        out.data = [[1] * len(d) for d in out.data]
        return out

    def bound_warnings(self, x) -> tuple[BoundWarning, ...]:
        """Free intrinsics within 1 % (s) / 0.01 (k1, k2) of a bound (bundle_parameterization.py:232-260)."""
        out: list[BoundWarning] = []
        for slot, b in self._free_slots():
            vals = (float(x[slot]), float(x[slot + 1]), float(x[slot + 2]))
            for name, v, (lo, hi), rel in zip(("f", "k1", "k2"), vals, _FREE_BOUNDS, (True, False, False)):
                for which, bnd in (("lower", lo), ("upper", hi)):
                    tol = 0.01 * bnd if rel else 0.01
                    if abs(v - bnd) <= tol:
                        out.append(BoundWarning(b.cam_id, name, which, v * b.fx_initial if name == "f" else v))
        return tuple(out)

    def intrinsic_estimates(self, camera_array: Any) -> tuple[IntrinsicEstimate, ...]:
        est = []
        for b in self.blocks:
            if b.free_intrinsics:
                cam = camera_array.cameras[b.cam_id]
                est.append(
                    IntrinsicEstimate(
                        cam_id=b.cam_id,
                        f_recovered=float(cam.matrix[0, 0]),
                        k1_recovered=float(cam.distortions[0]),
                        k2_recovered=float(cam.distortions[1]),
                        f_initial=b.fx_initial,
                        k1_initial=b.k1_initial,
                        k2_initial=b.k2_initial,
                    )
                )
        return tuple(est)
