"""The C++ .dat / codebook parsers of the host library (csrc/dat_format.cpp, via lib/libhostcheck.so)
against the byte layouts of SURVEY.md §8b as written by templates.py, including the loader quirks the
reference has (matcher.cpp:785-983)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from helpers import write_golden_files


@pytest.fixture(scope="module")
def hc(built):
    import __graft_entry__ as entry
    return C.CDLL(os.path.join(entry.PKG_DIR, "lib", "libhostcheck.so"))


def _rolled_counts(hc, path):
    v = [C.c_int(0) for _ in range(5)]
    rc = hc.hc_read_rolled(path.encode(), *[C.byref(x) for x in v])
    return rc, [x.value for x in v]


def _latent_counts(hc, path):
    a, b, t = C.c_int(0), C.c_int(0), C.c_int(0)
    slots = (C.c_int * 3)()
    rc = hc.hc_read_latent(path.encode(), C.byref(a), C.byref(b), slots, C.byref(t))
    return rc, a.value, b.value, list(slots), t.value


def test_rolled_parser_matches_reference_loader_codes(pkg, hc, golden, tmp_path):
    T = pkg.templates
    gdir, _ = write_golden_files(golden, str(tmp_path))
    for g, ref_rc in zip(golden["gallery_names"], golden["rolled_load_rc"]):
        p = os.path.join(gdir, f"{g}.dat")
        rc, (status, nmt, ntt, nm, nt) = _rolled_counts(hc, p)
        py = T.read_template(p, latent=False)
        if str(g) == "r18_empty":      # header-only file: undefined behaviour in the reference, EMPTY here
            assert (rc, status) == (1, 1)
            continue
        assert rc == int(ref_rc), g
        assert nmt == len(py.minu) and ntt == len(py.tex)
        assert nm == (py.minu[0].n if py.minu else 0)
        assert nt == min(py.tex[0].n if py.tex else 0, 1000)  # matcher.cpp:546-547


def test_rolled_parser_arrays_round_trip(pkg, hc, tmp_path):
    T = pkg.templates
    cb = T.synthetic_codebook()
    t = T.synth_rolled(5, cb, n_minu=33, n_tex=77)
    p = str(tmp_path / "r.dat")
    T.write_template(p, t)
    mx = np.zeros(33, np.int16); my = np.zeros(33, np.int16); mo = np.zeros(33, np.float32)
    md = np.zeros((33, 96), np.float32)
    tx = np.zeros(77, np.int16); ty = np.zeros(77, np.int16); to = np.zeros(77, np.float32)
    tc = np.zeros((77, 16), np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert hc.hc_rolled_arrays(p.encode(), vp(mx), vp(my), vp(mo), vp(md), vp(tx), vp(ty), vp(to), vp(tc)) == 0
    m, x = t.minu[0], t.tex[0]
    assert np.array_equal(mx, m.x) and np.array_equal(my, m.y) and np.array_equal(mo, m.ori) and np.array_equal(md, m.des)
    assert np.array_equal(tx, x.x) and np.array_equal(ty, x.y) and np.array_equal(to, x.ori) and np.array_equal(tc, x.des)


def test_latent_parser_selected_slots_and_index_shift(pkg, hc, tmp_path):
    T = pkg.templates
    raw = T.synth_rolled_raw(9, n_minu=40, n_tex=90)
    lat = T.synth_latent(1, raw, n_minu=20, n_tex_pts=30)
    for i, m in enumerate(lat.minu):  # give every template a distinctive size
        k = 5 + i
        lat.minu[i] = T.MinutiaeTemplate(m.x[:k % 20 + 1], m.y[:k % 20 + 1], m.ori[:k % 20 + 1], m.des[:k % 20 + 1])
    p = str(tmp_path / "l.dat")
    T.write_template(p, lat)
    rc, nm, nt, slots, ntex = _latent_counts(hc, p)
    assert (rc, nm, nt, ntex) == (0, 28, 1, 60)
    assert slots == [lat.minu[26].n, lat.minu[2].n, lat.minu[11].n]
    # an EMPTY minutiae record is skipped without a slot (matcher.cpp:834-836): later templates shift down
    lat.minu[1] = T.MinutiaeTemplate(np.zeros(0, np.int16), np.zeros(0, np.int16), np.zeros(0, np.float32),
                                     np.zeros((0, 96), np.float32))
    T.write_template(p, lat)
    rc, nm, nt, slots, ntex = _latent_counts(hc, p)
    assert (rc, nm, nt) == (0, 27, 1)
    assert slots == [lat.minu[27].n, lat.minu[3].n, lat.minu[12].n]  # positions 26, 2, 11 -> file records 27, 3, 12


def test_parser_limits_and_short_files(pkg, hc, tmp_path):
    T = pkg.templates
    p = str(tmp_path / "x.dat")
    open(p, "wb").write(b"")
    assert _rolled_counts(hc, p)[0] == 1 and _latent_counts(hc, p)[0] == 1
    open(p, "wb").write(b"\0" * 10)
    assert _rolled_counts(hc, p)[0] == 1            # length <= 10 (matcher.cpp:899-902)
    # more than 2000 minutiae in a record: loaders return 2 (matcher.cpp:837-841)
    hdr = struct.pack("<12H4H", 1, *([0] * 11), 800, 768, 50, 48)
    open(p, "wb").write(hdr + struct.pack("<B", 1) + struct.pack("<h", 2001))
    assert _rolled_counts(hc, p)[0] == 2 and _latent_counts(hc, p)[0] == 2
    # more than 2000 texture points: -1 (matcher.cpp:865-869)
    open(p, "wb").write(hdr + struct.pack("<B", 0) + struct.pack("<B", 1) + struct.pack("<h", 2001))
    assert _rolled_counts(hc, p)[0] == -1 and _latent_counts(hc, p)[0] == -1
    assert _rolled_counts(hc, str(tmp_path / "missing.dat"))[0] == -3


def _hc_write_latent(hc, path, lat):
    sets = list(lat.minu) + list(lat.tex)
    counts = np.array([s.n for s in sets], np.int32)
    cat = lambda f, dt, w=None: np.ascontiguousarray(np.concatenate([np.asarray(f(s), dt).reshape((-1,) if w is None else (-1, w)) for s in sets]))
    x, y, ori, des = cat(lambda s: s.x, np.int16), cat(lambda s: s.y, np.int16), cat(lambda s: s.ori, np.float32), \
        cat(lambda s: s.des, np.float32, 96)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    return hc.hc_write_latent(path.encode(), lat.h, lat.w, lat.blkH, lat.blkW, len(lat.minu), len(lat.tex), vp(counts), vp(x),
                              vp(y), vp(ori), vp(des))


def test_latent_writer_is_byte_identical_to_the_python_writer(pkg, hc, tmp_path):
    """write_latent_dat (csrc/dat_format.cpp) against templates.write_template, i.e. the layout of
    Template2Bin_Byte_latent (descriptor_PQ.py:80-175), including an empty minutiae template in the middle, and the
    reference's own loader accepting the file when oracle/_ref is present."""
    T = pkg.templates
    raw = T.synth_rolled_raw(12, n_minu=50, n_tex=120)
    lat = T.synth_latent(3, raw, n_minu=17, n_tex_pts=40)
    lat.blkH = 77  # clamped to 50 by the writer
    a, b = str(tmp_path / "c.dat"), str(tmp_path / "py.dat")
    assert _hc_write_latent(hc, a, lat) == 0
    T.write_template(b, lat)
    assert open(a, "rb").read() == open(b, "rb").read()
    rc, nm, nt, slots, ntex = _latent_counts(hc, a)
    assert (rc, nm, nt, slots, ntex) == (0, 28, 1, [17, 17, 17], 80)
    lat.minu[4] = T.MinutiaeTemplate(np.zeros(0, np.int16), np.zeros(0, np.int16), np.zeros(0, np.float32),
                                     np.zeros((0, 96), np.float32))
    assert _hc_write_latent(hc, a, lat) == 0
    T.write_template(b, lat)
    assert open(a, "rb").read() == open(b, "rb").read()
    import refbind
    if refbind.available():
        cbp = str(tmp_path / "cb.dat")
        T.write_codebook(cbp, T.synthetic_codebook())
        R = refbind.RefMatcher(cbp)
        _, rc = R.load_latent(a)
        assert rc == 0
        R.close()


def test_score_rows_are_what_the_reference_stream_writes(hc):
    """N-vs-N score file rows (matcher.cpp:198-205): "<path>",<score %.3f> per gallery file, the path quoted the way
    boost::filesystem's inserter does it ('&' escapes '"' and '&'), the score as iostream's fixed / setprecision(3)
    prints a float.  The driver formats whole files in one buffer instead of flushing every row."""
    rng = np.random.default_rng(3)
    paths = ["/data/rolled/0001.dat", "rel/dir with space/a.dat", 'odd"quote&amp.dat', "/x/y\\z.dat", ""]
    paths += [f"/g/{i:07d}.dat" for i in range(300)]
    scores = np.concatenate([np.array([-1.0, 0.0, -0.0, 0.0005, 0.0015, 2.5e-4, 1234567.875, 9.9995, 1e-9, -1e-9], np.float32),
                             rng.uniform(0, 40, len(paths) - 10).astype(np.float32)])
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    cap = 1 << 20
    buf = C.create_string_buffer(cap)
    hc.hc_format_score_rows.restype = C.c_long
    n = hc.hc_format_score_rows(arr, scores.ctypes.data_as(C.c_void_p), len(paths), buf, cap)
    got = buf.raw[:n].decode()
    want = "".join('"' + p.replace("&", "&&").replace('"', '&"') + '",' + "%.3f" % float(s) + "\n" for p, s in zip(paths, scores))
    assert got == want
    assert '"/data/rolled/0001.dat",-1.000\n' in got and ',-0.000\n' in got
