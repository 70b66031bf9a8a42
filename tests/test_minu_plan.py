"""The per-call tile plan of the fast minutiae kernels (csrc/minu_plan.h, via lib/libhostcheck.so): which latent
slots / gallery templates the shared-memory kernels take and which go to the HBM-resident kernels of minu_big.cuh.
Host logic only - the GPU tests check that both sides give the reference's results."""
import ctypes as C
import os

import numpy as np
import pytest

LIMIT = 227 * 1024 - 1024


@pytest.fixture(scope="module")
def plan(built):
    import __graft_entry__ as entry
    hc = C.CDLL(os.path.join(entry.PKG_DIR, "lib", "libhostcheck.so"))
    hc.hc_plan_minu.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p]

    def run(max_slot_n, max_nR, sizes=None):
        out = np.zeros(9, np.int64)
        h = None if sizes is None else np.ascontiguousarray(sizes, np.uint16)
        hc.hc_plan_minu(max_slot_n, max_nR, None if h is None else h.ctypes.data, 0 if h is None else len(h), out.ctypes.data)
        keys = ("l_cap", "r_cap", "b_double", "efficient", "slow_dense", "sim_smem", "sel_smem", "slow_smem", "job_stride")
        return dict(zip(keys, (int(v) for v in out)))
    return run


def test_benchmark_geometry_is_the_efficient_one(plan):
    # 80-minutiae latent slots, gallery templates of up to 150 minutiae: everything on the fast path,
    # double-buffered similarity kernel, four selection CTAs per SM - the numbers ncu reports for the bench
    p = plan(80, 150, np.full(1000, 150))
    assert (p["l_cap"], p["r_cap"], p["b_double"], p["efficient"]) == (80, 152, 1, 1)
    assert p["sim_smem"] == 210368 and p["sel_smem"] == 54960 and p["job_stride"] == 80 * 152


def test_every_plan_fits_the_shared_memory(plan):
    for L in (1, 7, 37, 80, 100, 128, 129, 500, 2000):
        for R in (1, 90, 156, 157, 200, 360, 361, 700, 2000):
            p = plan(L, R)
            assert p["l_cap"] == min(L, 128) and 4 <= p["r_cap"] <= ((R + 3) & ~3)
            assert max(p["sim_smem"], p["sel_smem"], p["slow_smem"]) <= LIMIT
            assert p["l_cap"] * p["r_cap"] < 65536            # 16-bit indices of the shared-memory introsort replay
            if p["r_cap"] < R:                                 # the cap is the LARGEST tile that fits
                assert plan(L, p["r_cap"] + 4)["r_cap"] == p["r_cap"]
    assert plan(80, 2000)["r_cap"] == 360 and plan(128, 2000)["r_cap"] == 216


def test_a_few_outsized_templates_do_not_change_the_geometry(plan):
    sizes = np.full(10000, 120)
    sizes[:40] = 420                                           # 0.4 % of the gallery
    p = plan(80, 420, sizes)
    assert p["efficient"] == 1 and p["r_cap"] == 156 and p["b_double"] == 1
    sizes[:80] = 420                                           # 0.8 %: the gallery is sized for them instead
    p = plan(80, 420, sizes)
    assert p["efficient"] == 0 and p["r_cap"] == 360
    # a gallery whose largest template still fits keeps it on the fast path when many templates are that large
    p = plan(80, 200, np.full(1000, 200))
    assert p["r_cap"] == 200 and p["b_double"] == 0
