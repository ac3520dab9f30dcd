"""The numeric CSV tables Caliscope keeps its observations and points in.

Same behaviour as the reference's writers / readers, byte for byte and dtype for dtype, with the formatting and parsing
done by all host cores in the native library (``csrc/cb_io.h``):

  write_table_csv(df, path)  ==  persistence._safe_write_csv(df, path, index=False, float_format="%.6f")
                                 (/root/reference/src/caliscope/persistence.py:27-41; called by ImagePoints.to_csv,
                                 core/point_data.py:358-373, and WorldPoints.to_csv, :662-677)
  read_table_csv(path)       ==  pd.read_csv(path)  for purely numeric tables (ImagePoints.from_csv :352-356,
                                 WorldPoints.from_csv :655-660)

Tables with a non-numeric column are not this path: both functions raise ``NotImplementedError`` for them.
``seam.install(full=True)`` swaps them in behind ``ImagePoints`` / ``WorldPoints`` ``.to_csv`` / ``.from_csv``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import _lib as L


def write_table_csv(df, path, n_threads: int = 0) -> None:
    """Byte-identical to ``df.to_csv(path, index=False, float_format="%.6f")`` + fsync + atomic rename."""
    lib = L.load()
    cols, kinds, keep = [], [], []
    for name in df.columns:
        a = df[name].to_numpy()
        if a.dtype.kind in "iu" or a.dtype == bool:
            if a.dtype == bool:
                raise NotImplementedError(f"column {name!r}: boolean columns are not part of this path")
            a = np.ascontiguousarray(a, dtype=np.int64)
            kinds.append(0)
        elif a.dtype.kind == "f":
            a = np.ascontiguousarray(a, dtype=np.float64)
            kinds.append(1)
        else:
            raise NotImplementedError(f"column {name!r} has dtype {a.dtype}: only numeric tables are handled natively")
        keep.append(a)
        cols.append(a.ctypes.data)
    header = ",".join(str(c) for c in df.columns)
    if any(("," in str(c) or '"' in str(c) or "\n" in str(c)) for c in df.columns):
        raise NotImplementedError("column names that need CSV quoting are not part of this path")
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    kinds_a = np.asarray(kinds, dtype=np.int32)
    ptrs = (C.c_void_p * len(cols))(*cols)
    L.check(lib.cb_csv_write_numeric(str(path).encode(), header.encode(), len(df), len(cols), kinds_a.ctypes.data,
                                     C.cast(ptrs, C.c_void_p), int(n_threads)), "csv_write")  # fmt: skip


def read_table_csv(path, n_threads: int = 0):
    """``pd.read_csv(path)`` for a purely numeric table: int64 where every field is an integer literal, float64 otherwise
    (empty field = NaN)."""
    import pandas as pd

    lib = L.load()
    path = str(path)
    with open(path, "rb") as f:
        header = f.readline().decode("utf-8").rstrip("\r\n")
    names = header.split(",")
    n_rows, n_cols = C.c_int64(), C.c_int32()
    L.check(lib.cb_csv_scan(path.encode(), C.byref(n_rows), C.byref(n_cols)), "csv_scan")
    if n_cols.value != len(names):
        raise ValueError(f"{path}: header has {len(names)} columns, scanner saw {n_cols.value}")
    n = int(n_rows.value)
    out = np.empty((len(names), max(n, 1)), dtype=np.float64)
    all_int = np.zeros(len(names), np.int32)
    has_empty = np.zeros(len(names), np.int32)
    if n:
        L.check(lib.cb_csv_parse_numeric(path.encode(), n, len(names), out.ctypes.data, all_int.ctypes.data, has_empty.ctypes.data,
                                         int(n_threads)), "csv_parse")  # fmt: skip
    data = {}
    for i, name in enumerate(names):
        col = out[i, :n]
        data[name] = col.astype(np.int64) if (n and all_int[i]) else col.copy()
        if n == 0:
            data[name] = np.zeros(0, dtype=object)  # pandas: empty body -> object columns
    return pd.DataFrame(data, columns=names)
