"""Minimal stand-in for the `rtoml` package (pinned 0.13.0 by the reference's
uv.lock but not installed in this image) so that the UNMODIFIED reference can be
imported by tests/golden/make_golden.py.  Build-container tooling only."""
from __future__ import annotations

import io
import tomllib
from pathlib import Path

import tomli_w


def _strip_none(obj):
    if isinstance(obj, dict):
        return {k: _strip_none(v) for k, v in obj.items() if v is not None}
    if isinstance(obj, (list, tuple)):
        return [_strip_none(v) for v in obj]
    return obj


def loads(s: str):
    return tomllib.loads(s)


def load(src):
    if isinstance(src, (str, Path)):
        p = Path(src)
        if p.exists():
            return tomllib.loads(p.read_text())
        return tomllib.loads(str(src))
    data = src.read()
    if isinstance(data, bytes):
        data = data.decode()
    return tomllib.loads(data)


def dumps(obj, *, pretty: bool = False, none_value=None) -> str:
    return tomli_w.dumps(_strip_none(obj))


def dump(obj, dst, *, pretty: bool = False, none_value=None) -> int:
    s = dumps(obj)
    if isinstance(dst, (str, Path)):
        Path(dst).write_text(s)
    elif isinstance(dst, io.TextIOBase):
        dst.write(s)
    else:
        dst.write(s.encode())
    return len(s)
