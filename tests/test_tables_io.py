"""Native numeric-CSV writer / reader (caliscope_b200/tables.py, csrc/cb_io.h) against pandas on the reference's own
files: byte-identical output, identical frames.  Host code only; runs on CPU."""
from __future__ import annotations

import io
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from caliscope_b200 import tables

REF = Path("/root/reference/tests/sessions")


def pandas_bytes(df) -> bytes:
    buf = io.StringIO(newline="")
    df.to_csv(buf, index=False, float_format="%.6f")
    return buf.getvalue().encode("utf-8")


def synthetic_frame(n=50_000, seed=0) -> pd.DataFrame:
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({
        "sync_index": rng.integers(0, 5000, n), "cam_id": rng.integers(0, 64, n), "object_id": rng.integers(0, 3, n),
        "keypoint_id": rng.integers(0, 88, n), "img_loc_x": rng.uniform(0, 1920, n), "img_loc_y": rng.uniform(0, 1080, n),
        "obj_loc_x": rng.normal(0, 0.3, n), "obj_loc_y": rng.normal(0, 1e-4, n), "obj_loc_z": np.zeros(n),
    })  # fmt: skip
    df.loc[rng.integers(0, n, 500), "obj_loc_z"] = np.nan  # planar trackers leave z empty
    df.loc[0, "img_loc_x"] = 0.0000005  # rounding ties and signs
    df.loc[1, "img_loc_x"] = -0.0000004
    df.loc[2, "img_loc_x"] = 1e15 + 0.5
    df.loc[3, "img_loc_x"] = 2.5e-7
    return df


def test_write_is_byte_identical_to_pandas(tmp_path):
    df = synthetic_frame()
    p = tmp_path / "xy.csv"
    tables.write_table_csv(df, p)
    assert p.read_bytes() == pandas_bytes(df)
    assert not (tmp_path / "xy.csv.tmp").exists()
    for threads in (1, 3):
        tables.write_table_csv(df, p, n_threads=threads)
        assert p.read_bytes() == pandas_bytes(df)


def test_read_equals_pandas_read_csv(tmp_path):
    df = synthetic_frame(20_000, seed=1)
    # pandas' default ("high") float parser is not correctly rounded beyond 17 significant digits; std::from_chars is.  Every
    # value this path writes ("%.6f" of pixel / metre quantities) is far below that, so the frames are identical there
    df.loc[2, "img_loc_x"] = 123456789.123456
    p = tmp_path / "xy.csv"
    p.write_bytes(pandas_bytes(df))
    got = tables.read_table_csv(p)
    ref = pd.read_csv(p)
    pd.testing.assert_frame_equal(got, ref, check_exact=True)
    assert list(got.dtypes) == list(ref.dtypes)


@pytest.mark.skipif(not REF.exists(), reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("rel", [
    "post_optimization/calibration/extrinsic/CHARUCO/xy_CHARUCO.csv",
    "post_optimization/calibration/extrinsic/CHARUCO/xyz_CHARUCO.csv",
    "larger_calibration_post_monocal/calibration/extrinsic/CHARUCO/xy_CHARUCO.csv",
])
def test_reference_session_files_round_trip(tmp_path, rel):
    src = REF / rel
    ref = pd.read_csv(src)
    got = tables.read_table_csv(src)
    pd.testing.assert_frame_equal(got, ref, check_exact=True)
    out = tmp_path / "out.csv"
    tables.write_table_csv(ref, out)
    assert out.read_bytes() == pandas_bytes(ref)


def test_non_numeric_tables_are_refused(tmp_path):
    df = pd.DataFrame({"a": [1, 2], "name": ["x", "y"]})
    with pytest.raises(NotImplementedError):
        tables.write_table_csv(df, tmp_path / "t.csv")
    (tmp_path / "t.csv").write_text("a,name\n1,x\n")
    with pytest.raises(Exception):
        tables.read_table_csv(tmp_path / "t.csv")
