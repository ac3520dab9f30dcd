"""Drop-in for the solver call inside ``CaptureVolume.optimize``
(/root/reference/src/caliscope/core/capture_volume.py:387-411):

    result = least_squares(joint_residuals, x0, args=(parameterization, camera_indices, image_coords,
                           image_to_world_indices, cga, cgb, cdist, cw), jac=joint_jacobian, verbose=...,
                           x_scale="jac", loss=..., f_scale=..., ftol=..., max_nfev=..., method="trf", bounds=...)

``least_squares`` below has that signature.  When ``fun`` is a ``joint_residuals`` (the reference's or
this package's) it never calls ``fun``/``jac``: it hands ``args`` to the CUDA engine and returns an
object with the fields the reference reads (``x``, ``status``, ``nfev``, ``cost``), rigid-distance
constraint rows included.  Anything else is not this path and raises ``NotImplementedError``: the
package has no CPU route of any kind.
"""
from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .problem import BAProblem, SolveResult, blocks_to_arrays

logger = logging.getLogger(__name__)

def is_bundle_adjustment_call(fun: Any, args: tuple) -> bool:
    return (
        callable(fun)
        and getattr(fun, "__name__", "") == "joint_residuals"
        and len(args) >= 4
        and hasattr(args[0], "blocks")
        and hasattr(args[0], "n_points")
    )


def _expected_bounds(par) -> tuple[np.ndarray, np.ndarray, int]:
    """The only bounds the engine implements: ``BundleParameterization.bounds()``
    (/root/reference/src/caliscope/core/bundle_parameterization.py:151-164): s in [0.5, 2], k1 in [-1, 1],
    k2 in [-2, 2] on cameras with free intrinsics, everything else unbounded.  Returns the camera part (the point part is
    all -inf / +inf) and the total parameter count."""
    widths = [9 if (b.free_intrinsics and not b.fisheye) else 6 for b in par.blocks]
    ncp = int(sum(widths))
    lo, hi = np.full(ncp, -np.inf), np.full(ncp, np.inf)
    o = 0
    for w in widths:
        if w == 9:
            lo[o + 6 : o + 9] = (0.5, -1.0, -2.0)
            hi[o + 6 : o + 9] = (2.0, 1.0, 2.0)
        o += w
    return lo, hi, ncp + 3 * int(par.n_points)


def _check_supported(par, lo, hi, use_bounds: bool, x_scale, tr_solver, n: int) -> None:
    """Fail loudly on a call the engine would otherwise answer with different semantics."""
    if use_bounds:
        elo, ehi, n_exp = _expected_bounds(par)
        ncp = len(elo)
        lo_b, hi_b = np.broadcast_to(lo, (n,)), np.broadcast_to(hi, (n,))
        same = n_exp == n and np.array_equal(lo_b[:ncp], elo) and np.array_equal(hi_b[:ncp], ehi)
        # the point part must be unbounded: two reductions instead of materialising +-inf vectors to compare with
        same = same and (n == ncp or (lo_b[ncp:].max() == -np.inf and hi_b[ncp:].min() == np.inf))
        if not same:
            raise NotImplementedError(
                "caliscope_b200.least_squares implements exactly BundleParameterization.bounds() "
                "(s in [0.5, 2], k1 in [-1, 1], k2 in [-2, 2] on free-intrinsics cameras); other bounds are not supported")
    if x_scale is not None and not (isinstance(x_scale, str) and x_scale == "jac"):
        raise NotImplementedError("caliscope_b200.least_squares implements x_scale='jac' only (Marquardt scaling)")
    if tr_solver not in (None, "lsmr"):
        raise NotImplementedError("caliscope_b200.least_squares replaces tr_solver='lsmr' (the sparse path) only")


def solve_arrays(cam_flags, cam_const, n_pts, camera_indices, obj_indices, image_coords, x0, *, use_bounds=True,
                 constraints=None, device: int = 0, **kw) -> SolveResult:  # fmt: skip
    """Array-level entry: build the device problem (optionally with rigid-distance rows), solve, free it."""
    with BAProblem(cam_flags, cam_const, n_pts, camera_indices, obj_indices, image_coords, constraints=constraints,
                   device=device) as prob:  # fmt: skip
        return prob.solve(x0, use_bounds=use_bounds, **kw)


def least_squares(fun, x0, jac="2-point", bounds=(-np.inf, np.inf), method="trf", ftol=1e-8, xtol=1e-8, gtol=1e-8,
                  x_scale=None, loss="linear", f_scale=1.0, diff_step=None, tr_solver=None, tr_options=None,
                  jac_sparsity=None, max_nfev=None, verbose=0, args=(), kwargs=None, callback=None, **extra):  # fmt: skip
    if not is_bundle_adjustment_call(fun, tuple(args)):
        raise NotImplementedError("caliscope_b200.least_squares only replaces the bundle-adjustment call "
                                  "(fun=joint_residuals, args=(parameterization, camera_indices, ...))")  # fmt: skip
    par, camera_indices, image_coords, obj_indices = args[:4]
    constraints = None
    if len(args) >= 8 and args[4] is not None and len(args[4]) > 0:
        constraints = (args[4], args[5], args[6], args[7])
    if method != "trf":
        raise ValueError("caliscope_b200.least_squares replaces method='trf' only")
    flags, const = blocks_to_arrays(par.blocks)
    lo, hi = (np.asarray(b, dtype=np.float64) for b in bounds) if isinstance(bounds, (tuple, list)) else (bounds.lb, bounds.ub)
    use_bounds = bool(np.any(np.isfinite(np.atleast_1d(lo))) or np.any(np.isfinite(np.atleast_1d(hi))))
    _check_supported(par, lo, hi, use_bounds, x_scale, tr_solver, np.asarray(x0).shape[0])
    res = solve_arrays(flags, const, par.n_points, np.asarray(camera_indices), np.asarray(obj_indices),  # int16 indices pass as is
                       np.asarray(image_coords, dtype=np.float64), np.asarray(x0, dtype=np.float64),
                       use_bounds=use_bounds, constraints=constraints, ftol=ftol if ftol is not None else 0.0, xtol=xtol if xtol is not None else 0.0,
                       gtol=gtol if gtol is not None else 0.0, max_nfev=max_nfev, loss=loss, f_scale=f_scale,
                       verbose=2 if verbose >= 2 else int(verbose))  # fmt: skip
    return res
