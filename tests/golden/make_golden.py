"""Generates tests/golden/golden_small.npz from the REFERENCE matcher itself.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/build_ref.sh):

    python tests/golden/make_golden.py

Inputs: synthetic templates in the reference's .dat layout (SURVEY.md §8d generators, written by
msu-latentafis_b200/templates.py) and the shipped PQ codebook.  Outputs recorded: for every
(latent, rolled) pair the return code, the four component scores, the surviving minutiae correspondences
(save_corr, matcher.cpp:497-505) and the fused score of
PQ::Matcher::One2One_matching_selected_templates (matching/matcher.cpp:376-417, fusion :188), obtained
through oracle/_ref/libref_matcher.so; and the score files the reference CLI oracle/_ref/match writes
in both modes.  The .npz stores the template files as raw bytes so that every consumer goes through
its own parser.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as entry  # noqa: E402
import refbind  # noqa: E402

T = entry.load_package().templates


def main():
    cb_path = T.find_codebook()
    assert cb_path, "the shipped codebook is needed (reference checkout)"
    cb = T.load_codebook(cb_path)
    tmp = tempfile.mkdtemp(prefix="lafis_golden_")
    gdir, ldir, sdir = (os.path.join(tmp, d) for d in ("gallery", "latent", "scores"))
    for d in (gdir, ldir, sdir):
        os.makedirs(d)

    # ---- gallery: 3 mates + varied impostors + edge cases ----
    raws = {}
    rolled = {}

    def add(name, raw):
        raws[name] = raw
        rolled[name] = T.rolled_from_raw(raw, cb)

    add("r00_mateA", T.synth_rolled_raw(0))
    add("r01_mateB", T.synth_rolled_raw(1, n_minu=100, n_tex=700))
    add("r02_mateC", T.synth_rolled_raw(2, n_minu=60, n_tex=400))
    for g in range(3, 12):
        add(f"r{g:02d}", T.synth_rolled_raw(g))
    add("r12_small", T.synth_rolled_raw(12, n_minu=5, n_tex=10))
    add("r13_tiny", T.synth_rolled_raw(13, n_minu=1, n_tex=1))
    add("r14_bigtex", T.synth_rolled_raw(14, n_minu=150, n_tex=1100))   # > 1000: truncated by the matcher
    add("r15_dense", T.synth_rolled_raw(15, n_minu=158, n_tex=1000))
    no_tex = T.rolled_from_raw(T.synth_rolled_raw(16), cb)
    no_tex.tex = []
    rolled["r16_notex"] = no_tex
    no_minu = T.rolled_from_raw(T.synth_rolled_raw(17), cb)
    no_minu.minu = [T.MinutiaeTemplate(np.zeros(0, np.int16), np.zeros(0, np.int16), np.zeros(0, np.float32),
                                       np.zeros((0, 96), np.float32))]
    rolled["r17_nominu"] = no_minu
    for name, t in rolled.items():
        T.write_template(os.path.join(gdir, name + ".dat"), t)
    T.write_template(os.path.join(gdir, "r18_empty.dat"), None)       # header-only "empty" file (PQ.py:190-193)
    with open(os.path.join(gdir, "r19_short.dat"), "wb") as f:        # <= 10 bytes: loader returns 1
        f.write(b"\x01\x00\x00\x00")

    # ---- latents ----
    latents = {
        "lA": T.synth_latent(0, raws["r00_mateA"]),
        "lB": T.synth_latent(1, raws["r01_mateB"], n_minu=30, n_tex_pts=120),
        "lC": T.synth_latent(2, raws["r02_mateC"], n_minu=12, n_tex_pts=75),
        "lG_30templates": T.synth_latent(3, raws["r03"], n_minu=10, n_tex_pts=40, n_minu_templates=30),
    }
    for name, t in latents.items():
        T.write_template(os.path.join(ldir, name + ".dat"), t)
    T.write_template(os.path.join(ldir, "lE_empty.dat"), None)

    # ---- pair scores from the reference library ----
    R = refbind.RefMatcher(cb_path)
    gnames = sorted(f[:-4] for f in os.listdir(gdir))
    lnames = sorted(latents)
    rhandles = []
    rolled_rc = []
    for g in gnames:
        h, rc = R.load_rolled(os.path.join(gdir, g + ".dat"))
        rhandles.append(h)
        rolled_rc.append(rc)
    pair_rc = np.zeros((len(lnames), len(gnames)), np.int32)
    pair_comp = np.zeros((len(lnames), len(gnames), 4), np.float32)
    pair_final = np.zeros((len(lnames), len(gnames)), np.float32)
    # surviving correspondences (the save_corr output, matcher.cpp:497-505) of every scored pair
    corr_n = np.zeros((len(lnames), len(gnames), 3), np.int32)
    corr_rows = []
    for i, l in enumerate(lnames):
        lh, _ = R.load_latent(os.path.join(ldir, l + ".dat"))
        for j, rh in enumerate(rhandles):
            rc, comp, fin = R.score_pair(lh, rh)
            pair_rc[i, j], pair_comp[i, j], pair_final[i, j] = rc, comp, fin
            if rc == 0 and rolled_rc[j] == 0:
                _, lists = R.correspondences(lh, rh)
                for s in range(3):
                    corr_n[i, j, s] = len(lists[s])
                    corr_rows.append(lists[s])
    corr_xy = np.concatenate(corr_rows).astype(np.int16) if corr_rows else np.zeros((0, 4), np.int16)
    R.close()

    # ---- score files from the reference CLI (it insists on ../afis.config relative to the cwd) ----
    work = os.path.join(tmp, "cwd")
    os.makedirs(work)
    with open(os.path.join(tmp, "afis.config"), "w") as f:
        f.write("{}")
    env = dict(os.environ, OMP_STACKSIZE="32M")
    subprocess.check_call([refbind.CLI_PATH, "-c", cb_path, "-s", sdir + "/", "-g", gdir, "-ldir", ldir], cwd=work, env=env,
                          stdout=subprocess.DEVNULL)
    n2n = {}
    for f in sorted(os.listdir(sdir)):
        n2n[f] = open(os.path.join(sdir, f)).read().replace(gdir, "@G@")
    sdir1 = os.path.join(tmp, "scores1")
    os.makedirs(sdir1)
    subprocess.check_call([refbind.CLI_PATH, "-c", cb_path, "-s", sdir1 + "/", "-g", gdir, "-l", os.path.join(ldir, "lA.dat")],
                          cwd=work, env=env, stdout=subprocess.DEVNULL)
    one2n = open(os.path.join(sdir1, "lA.csv")).read().replace(gdir, "@G@")

    blob = lambda p: np.frombuffer(open(p, "rb").read(), np.uint8)
    out = {
        "codebook": cb,
        "gallery_names": np.array(gnames),
        "latent_names": np.array(lnames),
        "all_latent_names": np.array(sorted(f[:-4] for f in os.listdir(ldir))),
        "rolled_load_rc": np.array(rolled_rc, np.int32),
        "pair_rc": pair_rc, "pair_comp": pair_comp, "pair_final": pair_final,
        "n2n_files": np.array(sorted(n2n)), "n2n_text": np.array([n2n[k] for k in sorted(n2n)]),
        "one2n_text": np.array(one2n),
        # corr_xy: rows "latent x, latent y, rolled x, rolled y" of all pairs with corr_n > 0, concatenated in
        # (latent, gallery, slot) order; corr_n[latent, gallery, slot] = rows of that list
        "corr_n": corr_n, "corr_xy": corr_xy,
    }
    for g in gnames:
        out["gal_" + g] = blob(os.path.join(gdir, g + ".dat"))
    for l in out["all_latent_names"]:
        out["lat_" + str(l)] = blob(os.path.join(ldir, str(l) + ".dat"))
    dst = os.path.join(ROOT, "tests", "golden", "golden_small.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    print("pair_final:\n", np.round(pair_final, 3))
    print("n2n files:", sorted(n2n))
    print(one2n)


if __name__ == "__main__":
    main()
