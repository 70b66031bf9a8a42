"""The plain-C oracle (oracle/lafis_oracle.c) against the REFERENCE matcher.

Always: the golden fixture (tests/golden/golden_small.npz, produced by tests/golden/make_golden.py from
oracle/_ref = the reference's own matcher.cpp) must be reproduced bit for bit.
When oracle/_ref/libref_matcher.so is present (build container, or shipped to the GPU box): additional
seeded pairs, every pruning stage separately, and the std::sort permutation emulation."""
import os

import numpy as np
import pytest

from helpers import UB_GALLERY, write_golden_files


def _refbind():
    import refbind
    return refbind if refbind.available() else None


def test_oracle_reproduces_golden_pairs(pkg, oracle, golden, tmp_path):
    T = pkg.templates
    gdir, ldir = write_golden_files(golden, str(tmp_path))
    cb = golden["codebook"]
    gnames = [str(g) for g in golden["gallery_names"]]
    rolled = [T.read_template(os.path.join(gdir, g + ".dat"), latent=False) for g in gnames]
    for i, l in enumerate(golden["latent_names"]):
        lat = T.read_template(os.path.join(ldir, f"{l}.dat"), latent=True)
        OL = oracle.OracleLatent(lat, cb)
        for j, g in enumerate(gnames):
            if g in UB_GALLERY:
                continue
            rc, comp, fin = oracle.score_pair(OL, oracle.OracleRolled(rolled[j]))
            want_rc = int(golden["pair_rc"][i, j])
            if want_rc != 0:
                assert rc == want_rc and fin == -1.0, (l, g)
                continue
            # (lG_30templates: 30 minutiae templates, score[28] is a never-written minutiae slot, matcher.cpp:188)
            assert rc == 0
            assert np.array_equal(comp, golden["pair_comp"][i, j]), (l, g, comp, golden["pair_comp"][i, j])
            assert fin == golden["pair_final"][i, j], (l, g)


def test_oracle_vs_reference_library_random_pairs(pkg, oracle, golden, tmp_path):
    rb = _refbind()
    if rb is None:
        pytest.skip("oracle/_ref not built")
    T = pkg.templates
    cb = golden["codebook"]
    cbp = str(tmp_path / "cb.dat")
    T.write_codebook(cbp, cb)
    R = rb.RefMatcher(cbp)
    raws = [T.synth_rolled_raw(2000 + g, n_minu=30 + 9 * g, n_tex=100 + 60 * g) for g in range(12)]
    lat = T.synth_latent(77, raws[4], n_minu=50, n_tex_pts=130)
    lp = str(tmp_path / "l.dat")
    T.write_template(lp, lat)
    lh, _ = R.load_latent(lp)
    OL = oracle.OracleLatent(lat, cb)
    for g, raw in enumerate(raws):
        r = T.rolled_from_raw(raw, cb)
        rp = str(tmp_path / f"r{g}.dat")
        T.write_template(rp, r)
        rh, _ = R.load_rolled(rp)
        rc1, comp1, fin1 = R.score_pair(lh, rh)
        rc2, comp2, fin2 = oracle.score_pair(OL, oracle.OracleRolled(r))
        assert rc1 == rc2 == 0
        assert np.array_equal(comp1, comp2) and fin1 == fin2, (g, comp1, comp2)
        # stage level: the three pruning routines on arbitrary candidate lists
        rng = np.random.default_rng(g)
        n = 60
        li = rng.integers(0, lat.minu[26].n, n).astype(np.int32)
        rj = rng.integers(0, r.minu[0].n, n).astype(np.int32)
        v = rng.uniform(0.1, 1.5, n).astype(np.float32)
        Lk = oracle.Keep(lat.minu[26].x, lat.minu[26].y, lat.minu[26].ori, des=lat.minu[26].des)
        Rk = oracle.Keep(r.minu[0].x, r.minu[0].y, r.minu[0].ori, des=r.minu[0].des)
        for which_ref, which_or in ((1, "dist_euclid"), (3, "angle")):
            a = R.prune(which_ref, lh, 26, rh, v, li, rj)
            b = oracle.prune(which_or, v, li, rj, Lk, Rk)
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), which_or
        lt, rt = lat.tex[0], r.tex[0]
        li = rng.integers(0, lt.n, n).astype(np.int32)
        rj = rng.integers(0, rt.n, n).astype(np.int32)
        Lk = oracle.Keep(lt.x, lt.y, lt.ori, des=lt.des)
        Rk = oracle.Keep(rt.x, rt.y, rt.ori, codes=rt.des)
        a = R.prune(0, lh, 0, rh, v, li, rj)
        b = oracle.prune("dist_lookup", v, li, rj, Lk, Rk, table=oracle.make_table())
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    R.close()


def test_oracle_sort_permutation_matches_libstdcxx(oracle):
    rb = _refbind()
    if rb is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for n in (1, 2, 15, 16, 17, 33, 120, 200, 1000, 9600):
        for levels in (0, 2, 7, 50):
            key = rng.standard_normal(n).astype(np.float32)
            if levels:
                key = np.round(key * levels) / levels  # many exact ties
            assert np.array_equal(oracle.std_sort_desc(key), rb.std_sort_desc(key)), (n, levels)


def _odd_layout_latents(T, raws):
    """Latents whose template counts put the texture score somewhere else than score[28] (matcher.cpp:414, :188):
    (minutiae templates, texture templates) = (1, 28), (2, 27) -> the texture score is score[1], score[2] and is
    fused with weight 1; (5, 24) -> it is score[5], which the fusion never reads.  (A latent file without any
    minutiae template cannot be written by the reference's writer: descriptor_PQ.py:92-95 emits the empty file.)"""
    out = []
    for k, (nm, nt) in enumerate(((1, 28), (2, 27), (5, 24))):
        full = T.synth_latent(300 + k, raws[k], n_minu=30, n_tex_pts=60)
        tex = full.tex[0]
        out.append(T.FPTemplate(h=full.h, w=full.w, blkH=full.blkH, blkW=full.blkW, minu=full.minu[:nm],
                                tex=[tex] + [T.TextureTemplate(tex.x[:3], tex.y[:3], tex.ori[:3], tex.des[:3])] * (nt - 1)))
    return out


def test_texture_score_in_an_unweighted_slot(oracle, golden, tmp_path):
    """Oracle port against the reference itself on latents with 1, 2 and 5 minutiae templates."""
    rb = _refbind()
    if rb is None:
        pytest.skip("oracle/_ref not built")
    import __graft_entry__ as entry
    T = entry.load_package().templates
    cb = golden["codebook"]
    cbp = str(tmp_path / "cb.dat")
    T.write_codebook(cbp, cb)
    R = rb.RefMatcher(cbp)
    raws = [T.synth_rolled_raw(4100 + g, n_minu=50, n_tex=150) for g in range(3)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    for k, lat in enumerate(_odd_layout_latents(T, raws)):
        lp = str(tmp_path / f"l{k}.dat")
        T.write_template(lp, lat)
        lh, _ = R.load_latent(lp)
        OL = oracle.OracleLatent(lat, cb)
        for g, r in enumerate(rolled):
            rp = str(tmp_path / f"r{g}.dat")
            T.write_template(rp, r)
            rh, _ = R.load_rolled(rp)
            rc1, comp1, fin1 = R.score_pair(lh, rh)
            rc2, comp2, fin2 = oracle.score_pair(OL, oracle.OracleRolled(r))
            assert rc1 == rc2 == 0, (k, g, rc1, rc2)
            assert np.array_equal(comp1, comp2) and fin1 == fin2, (k, g, comp1, comp2)
            if k < 2:
                assert comp1[3] == 0 and fin1 == comp1[k + 1] and (g != k or fin1 > 0)
            else:  # 5 minutiae templates: template 2 is matched (score[1]), the texture score (score[5]) is never read
                assert comp1[3] == 0 and comp1[0] == 0 and comp1[2] == 0 and fin1 == comp1[1]
    R.close()
