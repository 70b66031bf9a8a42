"""The C-ABI shared library: loads without a GPU, exports every entry point include/latentafis_b200.h
declares, fails loudly (no fallback) when no sm_100 device is present, and its host-only entry points
behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def test_library_exports_every_declared_symbol(pkg, built):
    L = pkg.load_library()
    header = open(os.path.join(ROOT, "include", "latentafis_b200.h")).read()
    declared = set(re.findall(r"LAFIS_API[^;(]*?\b(lafis_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(pkg.matcher.EXPORTS), declared ^ set(pkg.matcher.EXPORTS)
    assert b"sm_100a" in L.lafis_version()


def test_no_device_means_error_not_fallback(pkg, built):
    if have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.LafisError) as e:
        pkg.Matcher(codebook=pkg.templates.synthetic_codebook())
    assert e.value.status == pkg.matcher.LAFIS_ERR_CUDA
    bad = np.zeros((4, 256, 6), np.float32)
    ctx = C.c_void_p()
    assert pkg.load_library().lafis_create_from_codebook(bad.ctypes.data, 4, 256, 6, 0, C.byref(ctx)) == \
        pkg.matcher.LAFIS_ERR_CODEBOOK


def test_merge_hits_host(pkg, built):
    L = pkg.load_library()
    dt = pkg.matcher.HIT_DTYPE
    rng = np.random.default_rng(0)
    Q, n_lists, k = 3, 4, 5
    lists = np.zeros((Q, n_lists, k), dt)
    for q in range(Q):
        for s in range(n_lists):
            sc = np.sort(rng.integers(0, 6, k).astype(np.float32))[::-1]  # many ties across shards
            idx = np.sort(rng.choice(100, k, replace=False)) + 100 * s
            order = sorted(range(k), key=lambda i: (-sc[i], idx[i]))
            lists[q, s]["score"] = sc[order]
            lists[q, s]["index"] = idx[order]
    lists[0, 3]["index"][3:] = 0xFFFFFFFF  # short shard: empty slots
    lists[0, 3]["score"][3:] = -np.inf
    out = np.zeros((Q, k), dt)
    assert L.lafis_merge_hits(lists.ctypes.data, Q, n_lists, k, out.ctypes.data) == 0
    for q in range(Q):
        flat = [(float(h["score"]), int(h["index"])) for h in lists[q].reshape(-1) if h["index"] != 0xFFFFFFFF]
        want = sorted(flat, key=lambda t: (-t[0], t[1]))[:k]
        assert [(float(h["score"]), int(h["index"])) for h in out[q]] == want


def test_cli_binary_reports_missing_device(built, tmp_path):
    import subprocess
    import __graft_entry__ as entry
    exe = os.path.join(entry.PKG_DIR, "bin", "match")
    assert os.path.isfile(exe)
    if have_gpu():
        pytest.skip("a GPU is present")
    cb = tmp_path / "cb.dat"
    entry.load_package().templates.write_codebook(str(cb), entry.load_package().templates.synthetic_codebook())
    work = tmp_path / "cwd"
    work.mkdir()
    r = subprocess.run([exe, "-c", str(cb), "-s", str(tmp_path / "s") + "/", "-g", str(tmp_path), "-l", "x.dat"],
                       cwd=str(work), capture_output=True, text=True)
    assert r.returncode == 2 and "sm_100" in r.stderr
