"""PQ encoder parity.  The reference encodes with scipy.cluster.vq.vq (extraction/descriptor_PQ.py:3,
:26) on float32 descriptors and the float32 codebook; for 6-dimensional sub-vectors scipy evaluates
|o|^2 + |c|^2 - 2 o.c through BLAS, whose summation order is not ours.  The encoders here (Python tooling and the device
kernel, identical fp32 arithmetic: squared differences accumulated in dimension order) must therefore agree with the
reference wherever the nearest centroid is determined beyond fp32 rounding, and EVERY disagreement must be a near-tie:
the two chosen centroids equidistant to within the rounding error of the reference's own formula.

Golden vectors: tests/golden/golden_pq.npz, produced by the reference's TrainedPQEncoder class itself
(tests/golden/make_golden_pq.py), 512 constructed near-ties included."""
import os

import numpy as np
import pytest

from conftest import ROOT


def _golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_pq.npz"))
    cb = np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))["codebook"].astype(np.float32)
    return g["des"], g["codes"], cb


def check_near_ties(des, cb, got, want):
    """Every (point, sub-quantizer) where `got` differs from `want`: both centroids must be equally near up to the fp32
    rounding of |o|^2 + |c|^2 - 2 o.c (a few ulps of the largest term).  Returns the number of disagreements."""
    diff = np.argwhere(got != want)
    for n, m in diff:
        x = des[n, m * 6:(m + 1) * 6].astype(np.float64)
        ca, cbb = cb[m, got[n, m]].astype(np.float64), cb[m, want[n, m]].astype(np.float64)
        da, db = ((x - ca) ** 2).sum(), ((x - cbb) ** 2).sum()
        scale = (x ** 2).sum() + max((ca ** 2).sum(), (cbb ** 2).sum()) + 2 * max(abs(x @ ca), abs(x @ cbb))
        assert abs(da - db) <= 16 * 2.0 ** -24 * scale, (int(n), int(m), da, db, scale)
    # and `got` is a true nearest centroid up to the same rounding
    for n, m in diff:
        x = des[n, m * 6:(m + 1) * 6].astype(np.float64)
        d = ((x[None, :] - cb[m].astype(np.float64)) ** 2).sum(1)
        assert d[got[n, m]] - d.min() <= 16 * 2.0 ** -24 * ((x ** 2).sum() + (cb[m].astype(np.float64) ** 2).sum(1).max() * 2 + 1e-30)
    return len(diff)


def test_python_encoder_vs_reference_golden(pkg):
    des, want, cb = _golden()
    got = pkg.templates.pq_encode(des, cb)
    n_diff = check_near_ties(des, cb, got, want)
    # unconstrained points (rows 600..) are decided far beyond rounding: no disagreement there
    assert np.array_equal(got[600:], want[600:])
    assert n_diff <= 512 * 16


def test_python_encoder_vs_scipy_vq_live(pkg):
    """The same comparison against scipy.cluster.vq.vq called here, the way descriptor_PQ.py:26 calls it."""
    from scipy.cluster.vq import vq
    _, _, cb = _golden()
    rng = np.random.default_rng(11)
    des = (1.73 * rng.standard_normal((3000, 96)) / np.sqrt(96)).astype(np.float32)
    want = np.empty((3000, 16), np.uint8)
    for m in range(16):  # descriptor_PQ.py:25-26
        want[:, m], _ = vq(des[:, m * 6:(m + 1) * 6], cb[m])
    got = pkg.templates.pq_encode(des, cb)
    check_near_ties(des, cb, got, want)
    assert (got != want).mean() < 1e-3


@pytest.mark.gpu
def test_device_encoder_vs_reference_golden(pkg, built):
    des, want, cb = _golden()
    m = pkg.Matcher(codebook=cb, device=0)
    try:
        got = m.pq_encode(des)
        # identical arithmetic on the host and on the device: identical codes
        assert np.array_equal(got, pkg.templates.pq_encode(des, cb))
        check_near_ties(des, cb, got, want)
        assert np.array_equal(got[600:], want[600:])
    finally:
        m.close()
