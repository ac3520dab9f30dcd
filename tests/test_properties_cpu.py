"""Property tests (hypothesis) of the CPU-side building blocks: the oracle's analytic Jacobian against central
differences for random cameras of both models, the keep-mask rule against a row-by-row restatement of the
reference loop, and the triangulation key packing."""
from __future__ import annotations

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from caliscope_b200 import filtering
from caliscope_b200 import triangulation as T
from oracle import ba_oracle as O


def _random_rig(seed: int, fisheye: bool, free: bool):
    rng = np.random.default_rng(seed)
    n_cams, n_pts = 3, 12
    flags = np.array([(2 if (fisheye and c == 0) else (1 if free else 0)) for c in range(n_cams)], np.int32)
    const = np.zeros((n_cams, 9))
    for c in range(n_cams):
        fx = rng.uniform(500, 1500)
        const[c, :4] = [fx, fx * rng.uniform(0.97, 1.03), rng.uniform(300, 1000), rng.uniform(200, 600)]
        if flags[c] & 2:
            const[c, 4:8] = rng.normal(0, [0.05, 0.02, 0.005, 0.001])
        else:
            const[c, 4:9] = rng.normal(0, [0.1, 0.1, 0.002, 0.002, 0.05])
    cam = np.repeat(np.arange(n_cams, dtype=np.int32), n_pts)
    pt = np.tile(np.arange(n_pts, dtype=np.int32), n_cams)
    X = rng.uniform(-0.5, 0.5, (n_pts, 3))
    blocks = []
    for c in range(n_cams):
        r, t = rng.normal(0, 0.3, 3), np.array([0, 0, 3.0]) + rng.normal(0, 0.2, 3)
        blocks.append(np.concatenate([r, t, [rng.uniform(0.9, 1.1), const[c, 4], const[c, 5]]]) if flags[c] & 1
                      else np.concatenate([r, t]))  # fmt: skip
    x = np.concatenate(blocks + [X.ravel()])
    xy = rng.uniform(0, 1000, (len(cam), 2))
    return O.Rig(flags, const, n_pts, cam, pt, xy), x


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), fisheye=st.booleans(), free=st.booleans())
def test_oracle_jacobian_matches_central_differences(seed, fisheye, free):
    rig, x = _random_rig(seed, fisheye, free)
    J = O.jacobian(x, rig).toarray()
    h = 1e-6
    num = np.empty_like(J)
    for k in range(len(x)):
        e = np.zeros(len(x))
        e[k] = h
        num[:, k] = (O.residuals(x + e, rig) - O.residuals(x - e, rig)) / (2 * h)
    scale = np.maximum(np.abs(num).max(axis=0), 1e-12)
    assert (np.abs(J - num).max(axis=0) / scale).max() < 2e-6


def _reference_keep_loop(err, cam, thr, floor):
    """capture_volume.py:622-646 restated row by row."""
    keep = err <= thr[cam]
    for c in np.unique(cam):
        idx = cam == c
        n_keep, n_total = keep[idx].sum(), idx.sum()
        if n_keep < floor and n_keep < n_total:
            need = min(floor, n_total) - n_keep
            dropped = np.sort(err[idx & ~keep])
            if len(dropped) >= need:
                keep[idx] = err[idx] <= dropped[need - 1]
    return keep


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 400), n_cams=st.integers(1, 6), floor=st.integers(1, 60),
       q=st.floats(0.5, 99.5))  # fmt: skip
def test_keep_mask_equals_the_reference_loop(seed, n, n_cams, floor, q):
    rng = np.random.default_rng(seed)
    err = np.round(rng.gamma(2.0, 0.5, n), 2)  # rounded: ties at the threshold are the interesting case
    cam = rng.integers(0, n_cams, n)
    thr = np.array([np.percentile(err[cam == c], q) if np.any(cam == c) else np.inf for c in range(n_cams)])
    assert np.array_equal(filtering.keep_mask(err, cam, thr, floor), _reference_keep_loop(err, cam, thr, floor))


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(2, 300), lo=st.integers(-1000, 1000), span=st.integers(1, 2000))
def test_pack_keys_groups_exactly_like_lexsort(seed, n, lo, span):
    rng = np.random.default_rng(seed)
    s = rng.integers(lo, lo + span, n)
    o = rng.integers(0, 4, n)
    k = rng.integers(-3, 50, n)
    key = T.pack_keys(s, o, k)
    assert key.min() >= 0
    assert np.array_equal(np.argsort(key, kind="stable"), np.lexsort((k, o, s)))
    same_key = key[:, None] == key[None, :]
    same_tuple = (s[:, None] == s[None, :]) & (o[:, None] == o[None, :]) & (k[:, None] == k[None, :])
    assert np.array_equal(same_key, same_tuple)


# ---------------------------------------------------------------------------------------------------------------------
# bootstrap bookkeeping (host array versions; the device versions are compared with them in tests/test_gpu_bootstrap.py)
# ---------------------------------------------------------------------------------------------------------------------
@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 400), n_seg=st.integers(1, 9))
def test_segment_percentiles_equal_numpy_per_segment(seed, n, n_seg):
    """bootstrap._segment_percentiles == np.percentile(values[seg == p], [25, 75]) for every segment, bit for bit (the IQR
    thresholds of reject_outliers, pose_network_builder.py:369-400, are compared with < and >)."""
    from caliscope_b200 import bootstrap as B

    rng = np.random.default_rng(seed)
    seg = np.sort(rng.integers(0, n_seg, n))
    vals = rng.normal(size=n) * 10.0
    vals[rng.uniform(size=n) < 0.1] = np.round(vals[rng.uniform(size=n) < 0.1][:1].sum())  # some ties
    q1, q3 = B._segment_percentiles(vals, seg, n_seg)
    for p in range(n_seg):
        v = vals[seg == p]
        if len(v) == 0:
            continue
        a, b = np.percentile(v, [25, 75])
        assert q1[p] == a and q3[p] == b


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(2, 60), n_seg=st.integers(1, 5))
def test_segment_quaternion_average_equals_per_segment_eigenvector(seed, n, n_seg):
    """One batched 4x4 eigen-decomposition per pair == quaternion_average (pose_network_builder.py:416-438) of each pair's
    samples; a lone sample is returned itself."""
    from caliscope_b200 import bootstrap as B
    from oracle.ippe import _rodrigues

    rng = np.random.default_rng(seed)
    seg = np.sort(np.concatenate([np.arange(n_seg), rng.integers(0, n_seg, n)]))  # every segment non-empty
    base = rng.normal(size=3) * 0.5
    R = np.array([_rodrigues(base + rng.normal(size=3) * 0.05) for _ in seg])
    q = B._quat_wxyz(R)
    out = B._segment_quaternion_average(q, seg, n_seg)
    for p in range(n_seg):
        qs = q[seg == p]
        ref = B.quaternion_average(qs)
        assert np.abs(out[p] - ref).max() < 1e-12 or np.abs(out[p] + ref).max() < 1e-12
        if len(qs) == 1:
            assert (out[p] == qs[0]).all()


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 300))
def test_bootstrap_key_packing_is_an_order_isomorphism(seed, n):
    """bootstrap._pack maps (a, b, c) rows to one integer whose order is the lexicographic order of the rows and whose
    equality is row equality (what the device grouping sorts and compares)."""
    from caliscope_b200 import bootstrap as B

    rng = np.random.default_rng(seed)
    cols = [rng.integers(-5, 40, n), rng.integers(1000, 1030, n), rng.integers(0, 4, n)]
    key = B._pack(*cols)
    order = np.lexsort((cols[2], cols[1], cols[0]))
    assert (np.diff(key[order]) >= 0).all()
    rows = np.stack(cols, axis=1)
    same_key = key[:, None] == key[None, :]
    same_row = (rows[:, None, :] == rows[None, :, :]).all(axis=2)
    assert (same_key == same_row).all()
