"""Bit-exact helper code of the CUDA path, compiled for the host (lib/libhostcheck.so) and pinned
against glibc's atan2f and libstdc++'s std::sort — the two library routines whose exact behaviour the
reference's discrete decisions depend on (matcher.cpp:1516, :476)."""
import ctypes as C
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def hc(built):
    import __graft_entry__ as entry
    L = C.CDLL(os.path.join(entry.PKG_DIR, "lib", "libhostcheck.so"))
    L.hc_atan2f.restype = C.c_float
    L.hc_atan2f.argtypes = [C.c_float, C.c_float]
    return L


def _libm_atan2f(y, x):
    libm = C.CDLL("libm.so.6")
    libm.atan2f.restype = C.c_float
    libm.atan2f.argtypes = [C.c_float, C.c_float]
    return np.array([libm.atan2f(float(a), float(b)) for a, b in zip(y, x)], np.float32)


def test_atan2f_matches_glibc_on_integer_differences(hc):
    # the matcher only ever feeds integer coordinate differences (matcher.cpp:1513-1524): pixel
    # differences up to +-800 for minutiae, block differences up to +-50 for texture points
    rng = np.random.default_rng(0)
    dy = np.concatenate([np.arange(-60, 61).repeat(121), rng.integers(-800, 801, 40000)]).astype(np.float32)
    dx = np.concatenate([np.tile(np.arange(-60, 61), 121), rng.integers(-800, 801, 40000)]).astype(np.float32)
    got = np.zeros(len(dy), np.float32)
    hc.hc_atan2f_many(dy.ctypes.data_as(C.c_void_p), dx.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p),
                      C.c_long(len(dy)))
    want = _libm_atan2f(dy, dx)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_atan2f_matches_glibc_on_random_floats(hc):
    rng = np.random.default_rng(1)
    y = (rng.standard_normal(30000) * 10.0 ** rng.integers(-6, 7, 30000)).astype(np.float32)
    x = (rng.standard_normal(30000) * 10.0 ** rng.integers(-6, 7, 30000)).astype(np.float32)
    got = np.zeros(len(y), np.float32)
    hc.hc_atan2f_many(y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p),
                      C.c_long(len(y)))
    assert np.array_equal(got.view(np.uint32), _libm_atan2f(y, x).view(np.uint32))


@pytest.mark.parametrize("n", [1, 5, 16, 17, 40, 120, 200, 400, 2047, 9600, 15360])
def test_sort_prefix_equals_std_sort(hc, n):
    rng = np.random.default_rng(n)
    for levels in (0, 3, 11, 200):
        key = rng.standard_normal(n).astype(np.float32)
        if levels:
            key = (np.round(key * levels) / levels).astype(np.float32)
        if n > 100:
            key[rng.integers(0, n, n // 2)] = 0.0  # half of the similarity matrix is exactly 0 after the ReLU
        full = np.zeros(n, np.int32)
        hc.hc_std_sort(key.ctypes.data_as(C.c_void_p), n, full.ctypes.data_as(C.c_void_p))
        for need in sorted({1, min(n, 16), min(n, 17), min(n, 120), min(n, 200), n}):
            got = np.zeros(need, np.int32)
            hc.hc_sort_prefix(key.ctypes.data_as(C.c_void_p), n, need, got.ctypes.data_as(C.c_void_p))
            assert np.array_equal(got, full[:need]), (n, levels, need)
            gotc = np.zeros(need, np.int32)
            hc.hc_sort_prefix_closed(key.ctypes.data_as(C.c_void_p), n, need, gotc.ctypes.data_as(C.c_void_p))
            assert np.array_equal(gotc, full[:need]), ("closed form", n, levels, need)
            if n < 65536:
                got16 = np.zeros(need, np.int32)
                hc.hc_sort_prefix_u16(key.ctypes.data_as(C.c_void_p), n, need, got16.ctypes.data_as(C.c_void_p))
                assert np.array_equal(got16, full[:need])


def test_sort_prefix_adversarial_patterns(hc):
    # sorted, reversed, constant and organ-pipe inputs exercise the depth limit / heap-sort fallback
    n = 5000
    pats = [np.arange(n), np.arange(n)[::-1], np.zeros(n), np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]),
            np.arange(n) % 7]
    for p in pats:
        key = p.astype(np.float32)
        full = np.zeros(n, np.int32)
        hc.hc_std_sort(key.ctypes.data_as(C.c_void_p), n, full.ctypes.data_as(C.c_void_p))
        for need in (1, 120, 200, 2500, n):
            got = np.zeros(need, np.int32)
            hc.hc_sort_prefix(key.ctypes.data_as(C.c_void_p), n, need, got.ctypes.data_as(C.c_void_p))
            assert np.array_equal(got, full[:need])
            hc.hc_sort_prefix_closed(key.ctypes.data_as(C.c_void_p), n, need, got.ctypes.data_as(C.c_void_p))
            assert np.array_equal(got, full[:need])
