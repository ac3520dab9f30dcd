"""GPU-backed mirror of /root/reference/src/caliscope/core/reprojection.py.

Same function names and signatures (``project_points`` :18-32, ``reprojection_errors`` :35-72,
``joint_residuals`` :75-119, ``joint_jacobian`` :128-234); every number comes from the CUDA
engine through the C ABI -- there is no NumPy/OpenCV computation path here.
Distance-constraint rows (reprojection.py:112-117, 207-226) are evaluated on the GPU too
(``cb_ba_constraint_rows``); only the +-1/4 sparse pattern is assembled on the host.
"""
from __future__ import annotations

import numpy as np

from .problem import BAProblem, blocks_to_arrays

_CACHE: dict = {"key": None, "problem": None}


def problem_for(parameterization, camera_indices, image_coords, obj_indices, constraints=None, *, device: int = 0) -> BAProblem:
    """Device problem for one ``args`` tuple; cached on array identity so scipy-style repeated
    ``fun(x)`` / ``jac(x)`` calls (e.g. finite-difference tests) re-use the uploaded observation list."""
    cid = id(constraints[0]) if constraints is not None and constraints[0] is not None else 0
    key = (id(parameterization), id(camera_indices), id(image_coords), id(obj_indices), len(camera_indices), cid, device)
    if _CACHE["key"] == key and _CACHE["problem"] is not None:
        return _CACHE["problem"]
    if _CACHE["problem"] is not None:
        _CACHE["problem"].close()
    flags, const = blocks_to_arrays(parameterization.blocks)
    prob = BAProblem(flags, const, parameterization.n_points, np.asarray(camera_indices), np.asarray(obj_indices),
                     np.asarray(image_coords, dtype=np.float64), constraints=constraints, device=device)  # fmt: skip
    # keep the key objects alive so their ids cannot be recycled while the entry is cached
    _CACHE.update(key=key, problem=prob, refs=(parameterization, camera_indices, image_coords, obj_indices, constraints))
    return prob


def clear_cache() -> None:
    if _CACHE["problem"] is not None:
        _CACHE["problem"].close()
    _CACHE.update(key=None, problem=None, refs=None)


def _pack_constraints(ga, gb, dist, w):
    if ga is None or len(ga) == 0:
        return None
    return (ga, gb, dist, w)


def project_points(world, rvec, tvec, K, dist, fisheye: bool) -> np.ndarray:
    world = np.asarray(world, dtype=np.float64).reshape(-1, 3)
    d = np.asarray(dist, dtype=np.float64).ravel()
    K = np.asarray(K, dtype=np.float64)
    if fisheye:
        if d.shape[0] != 4:
            raise ValueError(f"Fisheye projection requires 4 distortion coefficients, got {d.shape[0]}")
        flags, c = 2, [K[0, 0], K[1, 1], K[0, 2], K[1, 2], d[0], d[1], d[2], d[3], 0.0]
    else:
        d5 = np.zeros(5)
        d5[: min(5, d.size)] = d[:5]
        flags, c = 0, [K[0, 0], K[1, 1], K[0, 2], K[1, 2], *d5]
    n = len(world)
    x = np.concatenate([np.ravel(rvec), np.ravel(tvec), world.ravel()]).astype(np.float64)
    with BAProblem(np.array([flags], np.int32), np.array([c]), n, np.zeros(n, np.int32), np.arange(n, dtype=np.int32),
                   np.zeros((n, 2))) as prob:  # fmt: skip
        return prob.reproj_errors_px(x)  # observed == 0  ->  error == projection


def reprojection_errors(camera_array, camera_indices, image_coords, world_coords) -> np.ndarray:
    """Pixel errors with each camera's stored intrinsics and extrinsics (reprojection.py:35-72)."""
    from .bundle_parameterization import BundleParameterization

    world_coords = np.asarray(world_coords, dtype=np.float64).reshape(-1, 3)
    par = BundleParameterization.from_camera_array(camera_array, n_points=len(world_coords), refine_intrinsics=False)
    x = par.pack(camera_array, world_coords)
    flags, const = blocks_to_arrays(par.blocks)
    n = len(world_coords)
    with BAProblem(flags, const, n, np.asarray(camera_indices), np.arange(n, dtype=np.int32),
                   np.asarray(image_coords, dtype=np.float64)) as prob:  # fmt: skip
        return prob.reproj_errors_px(x)


def joint_residuals(params, parameterization, camera_indices, image_coords, obj_indices, constraint_groups_a=None,
                    constraint_groups_b=None, constraint_distances=None, constraint_weights=None) -> np.ndarray:  # fmt: skip
    cons = _pack_constraints(constraint_groups_a, constraint_groups_b, constraint_distances, constraint_weights)
    return problem_for(parameterization, camera_indices, image_coords, obj_indices, cons).residuals(params)


def joint_jacobian(params, parameterization, camera_indices, image_coords, obj_indices, constraint_groups_a=None,
                   constraint_groups_b=None, constraint_distances=None, constraint_weights=None):  # fmt: skip
    """CSR matrix with the reference's row/column layout, assembled from the engine's dense blocks."""
    from scipy.sparse import csr_matrix

    cons = _pack_constraints(constraint_groups_a, constraint_groups_b, constraint_distances, constraint_weights)
    prob = problem_for(parameterization, camera_indices, image_coords, obj_indices, cons)
    Jc, Jp = prob.jacobian_blocks(params)
    cam = np.asarray(camera_indices, dtype=np.int64)
    pt = np.asarray(obj_indices, dtype=np.int64)
    widths = (prob.cam_offsets[1:] - prob.cam_offsets[:-1])[cam]
    nnz_row = np.repeat(widths + 3, 2)
    indptr = np.concatenate([[0], np.cumsum(nnz_row)]).astype(np.int64)
    data = np.empty(int(indptr[-1]))
    indices = np.empty(int(indptr[-1]), dtype=np.int64)
    ncp = prob.n_camera_params
    for w in np.unique(widths) if len(cam) else []:
        sel = np.nonzero(widths == w)[0]
        cols = np.concatenate([prob.cam_offsets[cam[sel]][:, None] + np.arange(w)[None],
                               ncp + 3 * pt[sel][:, None] + np.arange(3)[None]], axis=1)  # fmt: skip
        vals = np.concatenate([Jc[sel][:, :, :w], Jp[sel]], axis=2)
        for half in (0, 1):
            pos = indptr[2 * sel + half][:, None] + np.arange(w + 3)[None]
            data[pos] = vals[:, half, :]
            indices[pos] = cols
    J = csr_matrix((data, indices, indptr), shape=(2 * len(cam), prob.n_params))
    if cons is None:
        return J
    # constraint rows: +-1/4 of (weight * unit direction) per group column, repeats summed (reprojection.py:207-226)
    from scipy.sparse import coo_matrix, vstack

    _, d = prob.constraint_rows(params)
    n_c = len(d)
    rows, cols, vals = [], [], []
    for groups, sign in ((np.asarray(cons[0]), 0.25), (np.asarray(cons[1]), -0.25)):
        for q in range(4):
            base = ncp + 3 * groups[:, q].astype(np.int64)
            for k in range(3):
                rows.append(np.arange(n_c))
                cols.append(base + k)
                vals.append(sign * d[:, k])
    Cm = coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_c, prob.n_params))
    return vstack([J, Cm.tocsr()]).tocsr()
