"""GPU parity on the inputs that exercise the rare paths of the CUDA kernels: exact ties (introsort
replays), the integer filter's overflow / unquantised rows, odd template sizes and tile shapes, batches of
unequal latents.  Everything is compared bit for bit with the CPU oracle."""
import numpy as np
import pytest

from helpers import oracle_scores, rank_list

pytestmark = pytest.mark.gpu


def _run(pkg, cb, latents, rolled, oracle, topk=5):
    T = pkg.templates
    m = pkg.Matcher(codebook=cb, device=0)
    try:
        m.set_gallery(pkg.pack_rolled(rolled))
        out = m.match(m.latents_from_packed(pkg.pack_latents(latents)), topk=min(topk, len(rolled)), want_components=True)
        st = m.stats()
    finally:
        m.close()
    # a latent with more than 28 minutiae templates is outside the oracle's domain: the reference then fuses a
    # never-written minutiae slot as score[28] (matcher.cpp:188), i.e. 0 - emulate with the first 28 templates
    # and the texture component zeroed
    trimmed = [T.FPTemplate(minu=l.minu[:28], tex=l.tex) if len(l.minu) > 28 else l for l in latents]
    rc, comp, fin = oracle_scores(oracle, T, trimmed, rolled, cb)
    assert (rc == 0).all()
    for q, l in enumerate(latents):
        if len(l.minu) > 28:
            comp[q, :, 3] = 0.0
            for g in range(len(rolled)):
                fin[q, g] = oracle.lib().lo_fuse(float(comp[q, g, 0]), float(comp[q, g, 1]), float(comp[q, g, 2]), 0.0)
    bad = np.argwhere(out["components"] != comp)
    assert len(bad) == 0, (bad[:8], out["components"][tuple(bad[0][:2])], comp[tuple(bad[0][:2])])
    assert np.array_equal(out["scores"], fin)
    for q in range(len(latents)):
        assert list(out["hits"][q]["index"]) == rank_list(fin[q], min(topk, len(rolled)))
    return st


def test_template_sizes_and_tile_shapes(pkg, built, golden, oracle):
    T = pkg.templates
    cb = golden["codebook"]
    sizes = [(1, 1), (3, 7), (16, 16), (33, 100), (90, 600), (128, 800), (129, 810), (145, 999), (160, 1000), (161, 400),
             (200, 300), (257, 64)]
    raws = [T.synth_rolled_raw(3000 + k, n_minu=nm, n_tex=nt) for k, (nm, nt) in enumerate(sizes)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(80, raws[4], n_minu=37, n_tex_pts=57),     # nL not a multiple of 16, nLt = 114 < 200
               T.synth_latent(81, raws[7], n_minu=100, n_tex_pts=208),   # nLt = 416 > 200
               T.synth_latent(82, raws[10], n_minu=7, n_tex_pts=5)]
    _run(pkg, cb, latents, rolled, oracle)


def test_exact_ties_trigger_the_introsort_replays(pkg, built, golden, oracle):
    """Duplicated gallery minutiae / latent texture points give bit-identical normalised similarities and
    row maxima: the order std::sort leaves them in decides the candidate lists."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(3100 + k, n_minu=60, n_tex=300) for k in range(6)]
    for r in raws:  # duplicate descriptors (different coordinates): identical columns of S
        r.minu.des[1::2] = r.minu.des[0::2]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    lat = T.synth_latent(90, raws[2], n_minu=50, n_tex_pts=130)
    tex = lat.tex[0]
    tex.des[1::2] = tex.des[0::2]  # identical latent texture rows -> tied row maxima among the top 200
    st = _run(pkg, cb, [lat], rolled, oracle)
    assert st["minu_replays"] > 0 and st["tex_replays"] > 0, st


def test_unquantised_rows_and_queue_overflow(pkg, built, oracle):
    """A codebook of identical centroids makes every PQ distance-table row constant: rows whose entries are
    all ~0 are not quantised (every column is a candidate), other rows put every column inside the filter
    window - both overflow the candidate queue and fall back to the full exact evaluation; all similarities
    tie, so the FIRST column must win (std::max_element)."""
    T = pkg.templates
    cb = np.zeros((16, 256, 6), np.float32)
    raws = [T.synth_rolled_raw(3200 + k, n_minu=40, n_tex=120 + 30 * k) for k in range(4)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    lat = T.synth_latent(91, raws[1], n_minu=30, n_tex_pts=40)
    lat.tex[0].des[: 20] = 0.0  # LUT rows exactly 0 -> scale 0
    st = _run(pkg, cb, [lat], rolled, oracle)
    assert st["tex_overflow"] > 0, st


def test_batch_of_unequal_latents(pkg, built, golden, oracle):
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(3300 + k) for k in range(10)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(100 + q, raws[q], n_minu=20 + 13 * q, n_tex_pts=30 + 45 * q) for q in range(5)]
    latents.append(T.synth_latent(106, raws[6], n_minu=10, n_tex_pts=20, n_minu_templates=30))  # texture unweighted
    _run(pkg, cb, latents, rolled, oracle)


def test_cli_binary_matches_reference_score_files(pkg, built, golden, tmp_path):
    """The drop-in `match` executable, run the way the reference CLI is run (matching/main.cpp)."""
    import os
    import subprocess
    import __graft_entry__ as entry
    from helpers import UB_GALLERY, write_golden_files
    T = pkg.templates
    gdir, ldir = write_golden_files(golden, str(tmp_path))
    cbp = str(tmp_path / "codebook.dat")
    T.write_codebook(cbp, golden["codebook"])
    work = tmp_path / "cwd"
    work.mkdir()
    sdir = str(tmp_path / "scores") + "/"
    exe = os.path.join(entry.PKG_DIR, "bin", "match")
    r = subprocess.run([exe, "-c", cbp, "-s", sdir, "-g", gdir, "-ldir", ldir], cwd=str(work), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for fname, text in zip(golden["n2n_files"], golden["n2n_text"]):
        got = open(os.path.join(sdir, str(fname))).read().replace(gdir, "@G@")
        rows = lambda s: sorted(x for x in s.strip().split("\n") if not any(u in x for u in UB_GALLERY))
        assert rows(got) == rows(str(text)), fname
    r = subprocess.run([exe, "-c", cbp, "-s", sdir, "-g", gdir, "-l", os.path.join(ldir, "lB.dat")], cwd=str(work),
                       capture_output=True, text=True)
    assert r.returncode == 0 and "Match Results" in r.stdout
    first = open(os.path.join(sdir, "lB.csv")).read().split("\n")[1]
    assert first.startswith('1"') and "r01_mateB.dat" in first


def test_large_images_take_the_dense_graph_kernel(pkg, built, golden, oracle):
    """Minutiae coordinates beyond 2048 px (and negative ones): the sparse graph kernel's fp32 pre-test is only exact
    below 2048, such jobs must be routed to the dense kernel and still match the oracle bit for bit."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(5200 + k) for k in range(6)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(95, raws[1]), T.synth_latent(96, raws[4], n_minu=50, n_tex_pts=90)]

    def stretch(t, f, shift):
        for m in t.minu:
            m.x = (m.x.astype(np.int32) * f + shift).astype(np.int16)
            m.y = (m.y.astype(np.int32) * f + shift).astype(np.int16)
        return t

    rolled = [stretch(r, 3, 0) if k != 2 else stretch(r, 1, -500) for k, r in enumerate(rolled)]  # one with negative x, y
    latents = [stretch(l, 3, 0) for l in latents]
    assert max(int(r.minu[0].x.max()) for r in rolled) > 2048
    st = _run(pkg, cb, latents, rolled, oracle)
    assert st["kernel_launches"] > 0


def test_clustered_minutiae_take_the_second_chance_graph_kernel(pkg, built, golden, oracle):
    """Minutiae concentrated in a small area give distance-consistency graphs far denser than the ~9 % of spread-out
    impostor prints: they overflow the first sparse graph kernel's 2,560 non-zeros and are retried by
    graph_minu_mid_kernel (6,656 non-zeros) before anything goes to the dense kernel - same scores bit for bit."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(5300 + k, n_minu=120) for k in range(8)]
    rng = np.random.default_rng(5300)
    for k, r in enumerate(raws):  # sigma of 45..80 px around the centre (uniform over the image otherwise)
        sig = 45 + 5 * k
        r.minu.x = np.clip(np.rint(rng.normal(400, sig, r.minu.n)), 0, 799).astype(np.int16)
        r.minu.y = np.clip(np.rint(rng.normal(384, sig, r.minu.n)), 0, 767).astype(np.int16)
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(97, raws[2]), T.synth_latent(98, raws[6], n_minu=60, n_tex_pts=90)]
    st = _run(pkg, cb, latents, rolled, oracle)
    assert st["graph_minu_mid_jobs"] > 0, st
    assert st["graph_minu_dense_jobs"] < st["graph_minu_mid_jobs"], st  # most of them fit the second chance


def test_oversized_minutiae_templates(pkg, built, golden, oracle):
    """Templates beyond the shared-memory tiles of the fast minutiae kernels - up to the reference's own limit of 2000
    minutiae (matcher.cpp:788-841) - run through the HBM-resident kernels of minu_big.cuh: same candidate lists, same
    scores.  Includes a latent whose slots exceed 128 minutiae (every pair of it is oversized), exact ties (duplicate
    descriptors) and a template whose similarities are almost all zero (fewer than 120 positive values: the order is
    the introsort's, replayed over ~50,000 equal keys)."""
    T = pkg.templates
    cb = golden["codebook"]
    sizes = [(120, 300), (300, 200), (450, 150), (700, 120), (2000, 64), (600, 100), (130, 400)]
    raws = [T.synth_rolled_raw(3300 + k, n_minu=nm, n_tex=nt) for k, (nm, nt) in enumerate(sizes)]
    raws[2].minu.des[1::2] = raws[2].minu.des[0::2]   # ties among the normalised similarities
    raws[5].minu.des[1:] = 0.0                         # 599 all-zero columns of S
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(85, raws[1], n_minu=80, n_tex_pts=60),
               T.synth_latent(86, raws[3], n_minu=150, n_tex_pts=40),   # slots of 150 > 128: oversized against everything
               T.synth_latent(87, raws[4], n_minu=37, n_tex_pts=30)]
    st = _run(pkg, cb, latents, rolled, oracle)
    assert st["minu_big_jobs"] >= 3 * 7 + 2 * 3 * 4   # latent 1 x all templates, latents 0 and 2 x the templates beyond the tiles
    assert st["minu_replays"] > 0


def test_one_outsized_template_does_not_change_the_geometry(pkg, built, golden, oracle):
    """A gallery of ordinary templates plus ONE of 420 minutiae: the fast kernels keep their efficient tiles (at most
    0.5 % of the templates may exceed them) and the outlier alone goes through the oversized path."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(3400 + k, n_minu=100 + (k * 7) % 50, n_tex=80 + k % 40) for k in range(220)]
    raws.insert(57, T.synth_rolled_raw(3999, n_minu=420, n_tex=90))
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(88, raws[57], n_minu=80, n_tex_pts=50)]
    st = _run(pkg, cb, latents, rolled, oracle)
    assert st["minu_big_jobs"] == 3


def test_correspondences_of_an_oversized_pair(pkg, built, golden, tmp_path):
    """lafis_correspondences (save_corr, matcher.cpp:497-505) when the pair goes through minu_big.cuh: the lists the
    reference itself writes, and their similarities add up to the component scores."""
    import os
    import refbind
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(3500, n_minu=640, n_tex=100), T.synth_rolled_raw(3501, n_minu=110, n_tex=100)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(89, raws[0], n_minu=140, n_tex_pts=40)]
    m = pkg.Matcher(codebook=cb, device=0)
    try:
        m.set_gallery(pkg.pack_rolled(rolled))
        L = m.latents_from_packed(pkg.pack_latents(latents))
        out = m.match(L, topk=2, want_components=True)
        assert out["hits"][0]["index"][0] == 0 and m.stats()["minu_big_jobs"] == 6
        got = [m.correspondences(L, 0, g) for g in range(2)]
        for g in range(2):
            assert all(len(c) <= 120 for c in got[g])
        if refbind.available():
            cbp = os.path.join(str(tmp_path), "cb.dat")
            T.write_codebook(cbp, cb)
            R = refbind.RefMatcher(cbp)
            lp = os.path.join(str(tmp_path), "l.dat")
            T.write_template(lp, latents[0])
            lh, _ = R.load_latent(lp)
            for g, r in enumerate(rolled):
                rp = os.path.join(str(tmp_path), f"r{g}.dat")
                T.write_template(rp, r)
                rh, rc = R.load_rolled(rp)
                assert rc == 0
                _, comp, fin = R.score_pair(lh, rh)
                assert np.array_equal(out["components"][0, g], comp) and out["scores"][0, g] == fin
                _, want = R.correspondences(lh, rh)
                for s in range(3):
                    assert np.array_equal(got[g][s], want[s]), (g, s)
            R.close()
    finally:
        m.close()


def test_oversized_templates_across_pipeline_chunks(pkg, built, golden, oracle, monkeypatch):
    """The oversized-pair work list is built per pipeline chunk and per latent of a batch: a tiny work budget
    (several chunks, two streams) and a batch of two latents, one of them oversized itself, against the oracle."""
    T = pkg.templates
    cb = golden["codebook"]
    sizes = [(60, 80), (330, 80), (70, 90), (50, 70), (410, 60), (90, 75), (65, 85), (500, 50), (80, 64)]
    raws = [T.synth_rolled_raw(3600 + k, n_minu=nm, n_tex=nt) for k, (nm, nt) in enumerate(sizes)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(92, raws[4], n_minu=60, n_tex_pts=40), T.synth_latent(93, raws[7], n_minu=135, n_tex_pts=30)]
    monkeypatch.setenv("LAFIS_WORK_BYTES", "1500000")
    st = _run(pkg, cb, latents, rolled, oracle)
    assert st["minu_big_jobs"] == 3 * 9 + 3 * 3   # the oversized latent against everything + the other against 3 templates


def test_texture_score_in_an_unweighted_slot(pkg, built, golden, oracle):
    """Latents with 1 or 2 minutiae templates: the reference stores the texture score in score[1], score[2] (matcher.cpp:414)
    and fuses it with weight 1 (:188); with 5 minutiae templates it is never read.  Library vs oracle, bit for bit."""
    from test_oracle_vs_reference import _odd_layout_latents
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(4100 + g, n_minu=50, n_tex=150) for g in range(3)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = _odd_layout_latents(T, raws)
    matcher = pkg.Matcher(codebook=cb, device=0)
    matcher.set_gallery(pkg.pack_rolled(rolled))
    L = matcher.latents_from_packed(pkg.pack_latents(latents))
    assert [L.status(q) for q in range(3)] == [0, 0, 0]
    out = matcher.match(L, topk=2, want_components=True)
    L.free()
    matcher.close()
    rc, comp, fin = oracle_scores(oracle, T, latents, rolled, cb)
    assert (rc == 0).all()
    assert np.array_equal(out["components"], comp), (out["components"], comp)
    assert np.array_equal(out["scores"], fin)
    for k in range(2):
        assert out["hits"][k]["index"][0] == k and fin[k, k] > 0
    assert np.array_equal(fin[2], comp[2, :, 1]) and (comp[2, :, 3] == 0).all()  # texture score never read
