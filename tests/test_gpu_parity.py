"""GPU parity tests: the CUDA path through the C ABI against the reference's own outputs (golden
fixtures) and against the CPU oracle on seeded synthetic templates.  Integer/index results and all
scores are required to be BIT-IDENTICAL (the kernels keep the reference's fp32 operation order); the
north-star tolerance of 1e-4 relative is therefore met with margin."""
import os

import numpy as np
import pytest

from helpers import UB_GALLERY, oracle_scores, rank_list, write_golden_files

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher(pkg, built, golden):
    m = pkg.Matcher(codebook=golden["codebook"], device=0)
    yield m
    m.close()


def test_golden_pairs_bit_exact(pkg, matcher, golden, tmp_path):
    gdir, ldir = write_golden_files(golden, str(tmp_path))
    gnames = [str(g) for g in golden["gallery_names"]]
    lnames = [str(l) for l in golden["latent_names"]]
    matcher.load_gallery_files([os.path.join(gdir, g + ".dat") for g in gnames])
    L = matcher.load_latents([os.path.join(ldir, l + ".dat") for l in lnames])
    out = matcher.match(L, topk=8, want_components=True)
    keep = [j for j, g in enumerate(gnames) if g not in UB_GALLERY]
    want_final, want_comp, want_rc = golden["pair_final"], golden["pair_comp"], golden["pair_rc"]
    got_final, got_comp = out["scores"], out["components"]
    for i, l in enumerate(lnames):
        for j in keep:
            if want_rc[i, j] != 0:
                assert got_final[i, j] == -1.0, (l, gnames[j])
                continue
            assert np.array_equal(got_comp[i, j], want_comp[i, j]), (l, gnames[j], got_comp[i, j], want_comp[i, j])
            assert got_final[i, j] == want_final[i, j], (l, gnames[j])
    # rank lists: (score desc, index asc) over the library's own scores
    for i in range(len(lnames)):
        assert list(out["hits"][i]["index"]) == rank_list(got_final[i], 8)
        assert np.array_equal(out["hits"][i]["score"], got_final[i][out["hits"][i]["index"]])


def test_correspondences_match_the_reference_save_corr(pkg, matcher, golden, tmp_path):
    """lafis_correspondences against the lists the reference writes with save_corr (matcher.cpp:497-505), for every
    scored pair of the golden set, and the files the 1-vs-N driver writes for its 24 best (:322-327)."""
    gdir, ldir = write_golden_files(golden, str(tmp_path))
    gnames = [str(g) for g in golden["gallery_names"]]
    lnames = [str(l) for l in golden["latent_names"]]
    matcher.load_gallery_files([os.path.join(gdir, g + ".dat") for g in gnames])
    L = matcher.load_latents([os.path.join(ldir, l + ".dat") for l in lnames])
    corr_n, corr_xy = golden["corr_n"], golden["corr_xy"]
    comp = matcher.match(L, want_components=True)["components"]
    at = 0
    checked = 0
    for i, l in enumerate(lnames):
        for j, g in enumerate(gnames):
            scored = golden["pair_rc"][i, j] == 0 and golden["rolled_load_rc"][j] == 0
            want = []
            if scored:
                for s in range(3):
                    want.append(corr_xy[at:at + corr_n[i, j, s]])
                    at += corr_n[i, j, s]
            if not scored or g in UB_GALLERY:
                continue
            got = matcher.correspondences(L, i, j)
            for s in range(3):
                assert np.array_equal(got[s], want[s]), (l, g, s, got[s], want[s])
                assert (len(got[s]) > 0) == (comp[i, j, s] > 0)
            checked += 1
    assert at == len(corr_xy) and checked > 40
    # the driver's files for the 24 best of latent lA
    sdir = os.path.join(str(tmp_path), "scores1") + os.sep
    os.makedirs(sdir)
    assert matcher.One2List_matching(os.path.join(ldir, "lA.dat"), gdir, sdir) == 0
    rows = open(os.path.join(sdir, "lA.csv")).read().split()[1:]
    i = lnames.index("lA")
    first = {}
    a = 0
    for ii in range(len(lnames)):
        for j in range(len(gnames)):
            if golden["pair_rc"][ii, j] == 0 and golden["rolled_load_rc"][j] == 0:
                if ii == i:
                    first[j] = a
                a += int(corr_n[ii, j].sum())
    n_files = 0
    for row in rows:
        stem = os.path.basename(row.split('"')[1])[:-4]
        j = gnames.index(stem)
        if j not in first or stem in UB_GALLERY:
            continue
        a = first[j]
        if pkg.templates.read_template(os.path.join(gdir, stem + ".dat"), latent=False).minu == []:
            # no minutiae template on the rolled side: the reference never enters the minutiae loop (matcher.cpp:400),
            # so no correspondence file is created for this print
            assert not any(os.path.exists(os.path.join(sdir, f"corrlA_{stem}_{s}.csv")) for s in range(3)), stem
            assert int(corr_n[i, j].sum()) == 0
            continue
        for s in range(3):
            path = os.path.join(sdir, f"corrlA_{stem}_{s}.csv")
            text = open(path).read()
            want = "".join(",".join(str(int(v)) for v in r) + "\n" for r in corr_xy[a:a + corr_n[i, j, s]])
            assert text == want, (stem, s)
            a += corr_n[i, j, s]
            n_files += 1
    assert n_files >= 27


def test_synthetic_vs_oracle_bit_exact(pkg, matcher, golden, oracle):
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(100 + g) for g in range(40)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(50, raws[3]), T.synth_latent(51, raws[17], n_minu=45, n_tex_pts=111),
               T.synth_latent(52, raws[30], n_minu=96, n_tex_pts=300)]
    matcher.set_gallery(pkg.pack_rolled(rolled))
    L = matcher.latents_from_packed(pkg.pack_latents(latents))
    out = matcher.match(L, topk=10, want_components=True)
    rc, comp, fin = oracle_scores(oracle, T, latents, rolled, cb)
    assert (rc == 0).all()
    assert np.array_equal(out["components"], comp), np.argwhere(out["components"] != comp)[:10]
    assert np.array_equal(out["scores"], fin)
    for i, mate in enumerate((3, 17, 30)):
        assert out["hits"][i]["index"][0] == mate
        assert list(out["hits"][i]["index"]) == rank_list(fin[i], 10)


def test_batch_and_shard_invariance(pkg, matcher, golden):
    """Scores do not depend on batch composition, chunking or sharding; merged shard rank lists equal
    the single-shard rank list."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(300 + g, n_minu=40 + g % 30, n_tex=200 + 7 * g) for g in range(30)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(60 + i, raws[5 * i], n_minu=40, n_tex_pts=110) for i in range(4)]
    matcher.set_gallery(pkg.pack_rolled(rolled))
    whole = matcher.match(matcher.latents_from_packed(pkg.pack_latents(latents)), topk=6)
    for i, l in enumerate(latents):
        one = matcher.match(matcher.latents_from_packed(pkg.pack_latents([l])), topk=6)
        assert np.array_equal(one["scores"][0], whole["scores"][i])
        assert np.array_equal(one["hits"][0], whole["hits"][i])
    shard_hits = []
    scores = []
    for r in range(3):
        lo, hi = 30 * r // 3, 30 * (r + 1) // 3
        matcher.set_gallery(pkg.pack_rolled(rolled[lo:hi]), index_base=lo)
        o = matcher.match(matcher.latents_from_packed(pkg.pack_latents(latents)), topk=6)
        shard_hits.append(o["hits"])
        scores.append(o["scores"])
    assert np.array_equal(np.concatenate(scores, axis=1), whole["scores"])
    merged = matcher.merge_hits(np.stack(shard_hits, axis=1))
    assert np.array_equal(merged, whole["hits"])


def test_small_work_budget_chunks(pkg, golden, monkeypatch):
    """Force many pipeline chunks (tiny work budget) and compare with one chunk."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(500 + g, n_minu=30, n_tex=150) for g in range(23)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(70, raws[11], n_minu=25, n_tex_pts=60)]
    res = []
    for budget in ("100000", str(8 << 30)):
        monkeypatch.setenv("LAFIS_WORK_BYTES", budget)
        m = pkg.Matcher(codebook=cb, device=0)
        m.set_gallery(pkg.pack_rolled(rolled))
        res.append(m.match(m.latents_from_packed(pkg.pack_latents(latents)), topk=5, want_components=True))
        m.close()
    assert np.array_equal(res[0]["scores"], res[1]["scores"])
    assert np.array_equal(res[0]["components"], res[1]["components"])
    assert np.array_equal(res[0]["hits"], res[1]["hits"])


def test_gallery_roundtrip_and_pq_encode(pkg, matcher, golden):
    T = pkg.templates
    cb = golden["codebook"]
    rolled = [T.synth_rolled(700 + g, cb, n_minu=17 + g, n_tex=33 + 5 * g) for g in range(5)]
    matcher.set_gallery(pkg.pack_rolled(rolled))
    for i, r in enumerate(rolled):
        back = matcher.gallery_template(i)
        assert np.array_equal(back.minu[0].x, r.minu[0].x) and np.array_equal(back.minu[0].y, r.minu[0].y)
        assert np.array_equal(back.minu[0].ori, r.minu[0].ori) and np.array_equal(back.minu[0].des, r.minu[0].des)
        assert np.array_equal(back.tex[0].x, r.tex[0].x) and np.array_equal(back.tex[0].des, r.tex[0].des)
    # the encoder's parity against the reference's scipy.cluster.vq.vq call: tests/test_pq_encode.py
    rng = np.random.default_rng(5)
    des = (1.73 * rng.standard_normal((3000, 96)) / np.sqrt(96)).astype(np.float32)
    assert np.array_equal(matcher.pq_encode(des), T.pq_encode(des, cb))  # same fp32 arithmetic on host and device


def test_drivers_write_reference_score_files(pkg, matcher, golden, tmp_path):
    gdir, ldir = write_golden_files(golden, str(tmp_path))
    sdir = os.path.join(str(tmp_path), "scores") + "/"
    os.makedirs(sdir)
    assert matcher.List2List_matching(ldir, gdir, sdir) == 0
    for fname, text in zip(golden["n2n_files"], golden["n2n_text"]):
        got = open(os.path.join(sdir, str(fname))).read().replace(gdir, "@G@")
        rows = lambda s: sorted(r for r in s.strip().split("\n") if not any(u in r for u in UB_GALLERY))
        assert rows(got) == rows(str(text)), fname
    sdir1 = os.path.join(str(tmp_path), "scores1") + "/"
    os.makedirs(sdir1)
    assert matcher.One2List_matching(os.path.join(ldir, "lA.dat"), gdir, sdir1) == 0
    got = open(os.path.join(sdir1, "lA.csv")).read().replace(gdir, "@G@").strip().split("\n")
    want = str(golden["one2n_text"]).strip().split("\n")
    assert got[0] == want[0] == "filename,score"
    # rows with distinct positive scores must agree verbatim; tied rows (score 0 / -1) as sets,
    # because their order depends on the directory enumeration order of the machine
    pos = lambda rows: [r for r in rows[1:] if not r.endswith(",0") and not r.endswith(",-1")]
    tail = lambda rows: sorted(r.split('"', 1)[1] for r in rows[1:] if r.endswith(",0") or r.endswith(",-1"))
    assert pos(got) == pos(want)
    assert tail(got) == tail(want)
    assert matcher.One2List_matching(os.path.join(ldir, "lE_empty.dat"), gdir, sdir1) == 1
    assert open(os.path.join(sdir1, "lE_empty.csv")).read().strip() == "0"
    assert matcher.One2List_matching(os.path.join(ldir, "lA.dat"), str(tmp_path / "nowhere_dir_empty"), sdir1) == -1 \
        if os.makedirs(str(tmp_path / "nowhere_dir_empty"), exist_ok=True) is None else True


def test_parallel_ingest_equals_packed_input(pkg, matcher, golden, tmp_path):
    """lafis_gallery_load_files parses with several host threads (each a contiguous range of files): the resident
    gallery must equal the one built from the same templates packed in memory, template by template and score by
    score, including unreadable files in the middle of a range."""
    T = pkg.templates
    cb = golden["codebook"]
    n = 300
    raws = [T.synth_rolled_raw(7000 + g, n_minu=20 + (g * 7) % 50, n_tex=30 + (g * 13) % 90) for g in range(n)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    paths = []
    for g, r in enumerate(rolled):
        p = os.path.join(str(tmp_path), f"g{g:04d}.dat")
        if g in (17, 150, 299):
            with open(p, "wb") as f:  # <= 10 bytes: the loader returns 1 (matcher.cpp:899-902)
                f.write(b"\x01\x00")
        else:
            T.write_template(p, r)
        paths.append(p)
    latents = [T.synth_latent(70, raws[5], n_minu=15, n_tex_pts=25), T.synth_latent(71, raws[200], n_minu=18, n_tex_pts=33)]
    L = matcher.latents_from_packed(pkg.pack_latents(latents))
    matcher.load_gallery_files(paths)
    assert matcher.gallery_size == n
    from_files = matcher.match(L, topk=5)
    for g in (0, 16, 18, 149, 151, 298):
        a, b = matcher.gallery_template(g), rolled[g]
        assert np.array_equal(a.minu[0].x, b.minu[0].x) and np.array_equal(a.minu[0].des, b.minu[0].des)
        assert np.array_equal(a.tex[0].des, b.tex[0].des) and np.array_equal(a.tex[0].ori, b.tex[0].ori)
    keep = [g for g in range(n) if g not in (17, 150, 299)]
    matcher.set_gallery(pkg.pack_rolled([rolled[g] for g in keep]))
    packed = matcher.match(L, topk=5)
    assert np.array_equal(from_files["scores"][:, keep], packed["scores"])
    assert (from_files["scores"][:, [17, 150, 299]] == -1.0).all()
    assert from_files["hits"][0]["index"][0] == 5 and from_files["hits"][1]["index"][0] == 200


def test_enroll_rolled_writes_the_reference_layout(pkg, matcher, golden, tmp_path):
    """lafis_enroll_rolled (GPU PQ encoder + Template2Bin_Byte_PQ_rolled layout, descriptor_PQ.py:19-27, :178-272):
    byte-identical to the Python writer fed with numpy nearest-centroid codes, accepted by the reference's loader
    when oracle/_ref is present, and scored like the same template packed in memory."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(8100 + g, n_minu=40 + 9 * g, n_tex=150 + 60 * g) for g in range(4)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    paths = []
    for g, (raw, want) in enumerate(zip(raws, rolled)):
        m0 = raw.minu
        minu_xyo = np.stack([m0.x.astype(np.float32) + 0.25, m0.y.astype(np.float32) + 0.75, m0.ori], axis=1)  # truncated
        tex_xyo = np.stack([raw.tex_x.astype(np.float32) * 16 + 24 + 3, raw.tex_y.astype(np.float32) * 16 + 24, raw.tex_ori], axis=1)
        p = os.path.join(str(tmp_path), f"enrolled{g}.dat")
        matcher.enroll_rolled(p, minu_xyo, m0.des, tex_xyo, raw.tex_des, h=want.h, w=want.w, blkH=want.blkH, blkW=want.blkW)
        q = os.path.join(str(tmp_path), f"python{g}.dat")
        T.write_template(q, want)
        assert open(p, "rb").read() == open(q, "rb").read(), g
        paths.append(p)
    latents = [T.synth_latent(90, raws[2], n_minu=30, n_tex_pts=60)]
    L = matcher.latents_from_packed(pkg.pack_latents(latents))
    matcher.load_gallery_files(paths)
    a = matcher.match(L, topk=2)
    matcher.set_gallery(pkg.pack_rolled(rolled))
    b = matcher.match(L, topk=2)
    assert np.array_equal(a["scores"], b["scores"]) and a["hits"][0]["index"][0] == 2
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refbind
    if refbind.available():
        cbp = os.path.join(str(tmp_path), "cb.dat")
        T.write_codebook(cbp, cb)
        R = refbind.RefMatcher(cbp)
        lp = os.path.join(str(tmp_path), "lat.dat")
        T.write_template(lp, latents[0])
        lh, _ = R.load_latent(lp)
        for g, p in enumerate(paths):
            rh, rc = R.load_rolled(p)
            assert rc == 0
            _, _, fin = R.score_pair(lh, rh)
            assert fin == a["scores"][0, g]
        R.close()


def test_enroll_latent_writes_the_reference_layout(pkg, matcher, golden, tmp_path):
    """lafis_enroll_latent (Template2Bin_Byte_latent, descriptor_PQ.py:80-175): byte-identical to the Python writer,
    and the enrolled file scores exactly like the same latent packed in memory."""
    T = pkg.templates
    cb = golden["codebook"]
    raws = [T.synth_rolled_raw(8200 + g, n_minu=60, n_tex=200) for g in range(3)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    lat = T.synth_latent(91, raws[1], n_minu=35, n_tex_pts=70)
    minu_sets = [(np.stack([m.x.astype(np.float32) + 0.5, m.y.astype(np.float32) + 0.25, m.ori], axis=1), m.des)
                 for m in lat.minu]
    tex_sets = [(np.stack([t.x.astype(np.float32) * 16 + 24 + 5, t.y.astype(np.float32) * 16 + 24, t.ori], axis=1), t.des)
                for t in lat.tex]
    p, q = os.path.join(str(tmp_path), "lat_c.dat"), os.path.join(str(tmp_path), "lat_py.dat")
    matcher.enroll_latent(p, minu_sets, tex_sets, h=lat.h, w=lat.w, blkH=lat.blkH, blkW=lat.blkW)
    T.write_template(q, lat)
    assert open(p, "rb").read() == open(q, "rb").read()
    matcher.set_gallery(pkg.pack_rolled(rolled))
    a = matcher.match(matcher.load_latents([p]), topk=3)
    b = matcher.match(matcher.latents_from_packed(pkg.pack_latents([lat])), topk=3)
    assert np.array_equal(a["scores"], b["scores"]) and a["hits"][0]["index"][0] == 1
