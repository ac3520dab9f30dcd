"""Install / remove the drop-in behind Caliscope's bundle-adjustment call.

Seam S1: the module attribute ``caliscope.core.capture_volume.least_squares``
(bound by ``from scipy.optimize import least_squares``, capture_volume.py:2) is the only thing
``CaptureVolume.optimize`` calls to solve; replacing it leaves every Caliscope API -- PointData,
CameraArray, BundleParameterization, aniposelib export -- untouched.

    import caliscope_b200.seam as seam
    with seam.installed():            # or seam.install() / seam.uninstall()
        volume = volume.optimize()    # runs on the B200
"""
from __future__ import annotations

import contextlib
import functools
import importlib

from . import solver

_TARGET = "caliscope.core.capture_volume"
_original = None
_original_methods: dict = {}
_original_functions: dict = {}


def install(full: bool = False) -> None:
    """Seam S1: replace the ``least_squares`` attribute ``CaptureVolume.optimize`` calls.

    A call that is not the bundle-adjustment call raises ``NotImplementedError``; there is no CPU route.
    ``full=True`` additionally installs seam S2: ``CaptureVolume.optimize`` and
    ``CaptureVolume._compute_img_to_obj_map`` are replaced by the vectorised versions in
    ``caliscope_b200.capture_volume`` (same signatures and results, no per-row Python loops),
    ``CaptureVolume.reprojection_report`` by the engine-backed, bincount-aggregated version, the percentile /
    threshold filters by merge-free versions (``filter_by_absolute_error`` reaches them unchanged), and seam
    S3: ``caliscope.core.point_data.triangulate_image_points`` (point_data.py:122-229) and
    ``ImagePoints.triangulate`` (:416-559; pixels -> undistortion -> DLT in one device call) become the GPU
    versions; and seam S4: the stage functions of the extrinsic bootstrap
    (``caliscope.core.bootstrap_pose.pose_network_builder``: PnP per group, relative poses, outlier rejection, aggregation,
    stereo RMSE) become ``caliscope_b200.bootstrap``'s; and seam S5: ``ImagePoints`` / ``WorldPoints`` ``.to_csv`` /
    ``.from_csv`` go through the native numeric-CSV writer / parser (``caliscope_b200.tables``; byte-identical files)."""
    global _original
    from . import _lib

    _lib.load()  # fail loudly now if the CUDA library is missing
    mod = importlib.import_module(_TARGET)
    if _original is None:
        _original = mod.least_squares
    mod.least_squares = solver.least_squares
    if full:
        from . import capture_volume as cv2b

        cls = mod.CaptureVolume
        if not _original_methods:
            _original_methods.update(optimize=cls.optimize, _compute_img_to_obj_map=cls._compute_img_to_obj_map,
                                     reprojection_report=cls.__dict__["reprojection_report"],
                                     _filter_by_reprojection_thresholds=cls._filter_by_reprojection_thresholds,
                                     filter_by_percentile_error=cls.filter_by_percentile_error)
        cls.optimize = cv2b.optimize
        cls._compute_img_to_obj_map = cv2b.fast_img_to_obj_map
        report = functools.cached_property(cv2b.reprojection_report)
        report.__set_name__(cls, "reprojection_report")
        cls.reprojection_report = report
        cls._filter_by_reprojection_thresholds = cv2b.filter_by_reprojection_thresholds
        cls.filter_by_percentile_error = cv2b.filter_by_percentile_error
        from . import triangulation

        pd_mod = importlib.import_module("caliscope.core.point_data")
        if "triangulate_image_points" not in _original_functions:
            _original_functions["triangulate_image_points"] = pd_mod.triangulate_image_points
        pd_mod.triangulate_image_points = triangulation.triangulate_image_points
        if "ImagePoints.triangulate" not in _original_functions:
            _original_functions["ImagePoints.triangulate"] = pd_mod.ImagePoints.triangulate
        pd_mod.ImagePoints.triangulate = triangulation.triangulate
        # seam S4: the extrinsic bootstrap (pose_network_builder.py): the five stage functions PoseNetworkBuilder calls are
        # module attributes, so build_paired_pose_network / PoseNetworkBuilder run unchanged on top of the GPU stages
        from . import bootstrap

        pnb = importlib.import_module("caliscope.core.bootstrap_pose.pose_network_builder")
        for name in _BOOTSTRAP_FUNCTIONS:
            if name not in _original_bootstrap:
                _original_bootstrap[name] = getattr(pnb, name)
            setattr(pnb, name, getattr(bootstrap, name))
        # ... and the entry point above them (capture_volume.py:287-315 imports it at call time): the whole PnP branch as
        # three device calls, no per-pose Python objects
        bpn = importlib.import_module("caliscope.core.bootstrap_pose.build_paired_pose_network")
        if "build_paired_pose_network" not in _original_bootstrap:
            _original_bootstrap["build_paired_pose_network"] = bpn.build_paired_pose_network
        bpn.build_paired_pose_network = bootstrap.build_paired_pose_network


        # seam S5: the numeric CSV tables (ImagePoints / WorldPoints .to_csv / .from_csv) through the native writer / parser
        from . import tables

        pd_mod = importlib.import_module("caliscope.core.point_data")
        for cls_name in ("ImagePoints", "WorldPoints"):
            cls_t = getattr(pd_mod, cls_name)
            if (cls_name, "to_csv") not in _original_tables:
                _original_tables[(cls_name, "to_csv")] = cls_t.__dict__["to_csv"]
                _original_tables[(cls_name, "from_csv")] = cls_t.__dict__["from_csv"]
            cls_t.to_csv = _make_to_csv(cls_name, tables)
            cls_t.from_csv = classmethod(_make_from_csv(tables))


def _make_to_csv(cls_name: str, tables):
    def to_csv(self, path) -> None:
        """Same file, byte for byte, as the reference's to_csv (point_data.py:358-373 / :662-677)."""
        from pathlib import Path

        from caliscope.persistence import PersistenceError

        try:
            tables.write_table_csv(self._df, Path(path))
        except NotImplementedError:
            raise
        except Exception as e:
            what = "image points" if cls_name == "ImagePoints" else "world points"
            raise PersistenceError(f"Failed to save {what} to {path}: {e}") from e

    return to_csv


def _make_from_csv(tables):
    def from_csv(cls, path):
        return cls(tables.read_table_csv(path))

    return from_csv


_original_tables: dict = {}
_BOOTSTRAP_FUNCTIONS = ("compute_camera_to_object_poses_pnp", "compute_relative_poses", "reject_outliers", "aggregate_poses",
                        "estimate_pnp_paired_pose_network")
_original_bootstrap: dict = {}


def uninstall() -> None:
    global _original
    mod = importlib.import_module(_TARGET)
    if _original is not None:
        mod.least_squares = _original
        _original = None
    if _original_methods:
        for name, fn in _original_methods.items():
            setattr(mod.CaptureVolume, name, fn)
        _original_methods.clear()
    if _original_functions:
        pd_mod = importlib.import_module("caliscope.core.point_data")
        for name, fn in _original_functions.items():
            if name == "ImagePoints.triangulate":
                pd_mod.ImagePoints.triangulate = fn
            else:
                setattr(pd_mod, name, fn)
        _original_functions.clear()
    if _original_tables:
        pd_mod = importlib.import_module("caliscope.core.point_data")
        for (cls_name, attr), fn in _original_tables.items():
            setattr(getattr(pd_mod, cls_name), attr, fn)
        _original_tables.clear()
    if _original_bootstrap:
        pnb = importlib.import_module("caliscope.core.bootstrap_pose.pose_network_builder")
        bpn = importlib.import_module("caliscope.core.bootstrap_pose.build_paired_pose_network")
        for name, fn in _original_bootstrap.items():
            setattr(bpn if name == "build_paired_pose_network" else pnb, name, fn)
        _original_bootstrap.clear()


def install_from_env() -> bool:
    """Install the seam when the environment asks for it: ``CALISCOPE_BA_BACKEND=b200`` (seam S1: the solver call only) or
    ``CALISCOPE_BA_BACKEND=b200-full`` (S1-S5).  Returns whether anything was installed.  Meant to be called once by the
    application that owns the process (e.g. at the top of a calibration script); nothing in this package calls it."""
    import os

    mode = os.environ.get("CALISCOPE_BA_BACKEND", "").strip().lower()
    if mode == "b200":
        install(False)
        return True
    if mode in ("b200-full", "b200_full"):
        install(True)
        return True
    return False


@contextlib.contextmanager
def installed(full: bool = False):
    install(full)
    try:
        yield
    finally:
        uninstall()
