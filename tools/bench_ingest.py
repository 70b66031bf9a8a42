"""Gallery ingest throughput and file-to-file end to end at gallery scale.

  python tools/bench_ingest.py [n_distinct] [n_files]

n_distinct synthetic rolled prints are enrolled (lafis_enroll_rolled: GPU PQ encoder + writer) as .dat files; a second
directory of n_files entries is made of HARD LINKS to them (100,000 x 66 KB of distinct files is not practical to
write here; the links give the loader 100,000 directory entries and 6.6 GB to read, page cache warm).  Measured:
  * enrollment rate;
  * lafis_gallery_load_dir on both directories (parallel parse on all host cores, pinned staging rings, device pool,
    re-layout into the resident layout), best of 3;
  * bin/match -l on the large directory: process start, ingest, one 1-vs-N search, score + correspondence files.
Prints one JSON object."""
import json, os, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
T = pkg.templates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_big = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
g = os.path.join(ROOT, "tests", "golden", "golden_small.npz")
cb = np.load(g)["codebook"] if os.path.isfile(g) else T.synthetic_codebook()
m = pkg.Matcher(codebook=cb, device=0)
tmp = tempfile.mkdtemp(prefix="lafis_ingest_")
out = {"host_cores": os.cpu_count()}
try:
    small, big = os.path.join(tmp, "small"), os.path.join(tmp, "big")
    os.makedirs(small); os.makedirs(big)
    raws = [T.synth_rolled_raw(k) for k in range(n)]
    t0 = time.perf_counter()
    for k, r in enumerate(raws):
        mm = r.minu
        minu_xyo = np.stack([mm.x.astype(np.float32), mm.y.astype(np.float32), mm.ori], axis=1)
        tex_xyo = np.stack([r.tex_x.astype(np.float32) * 16 + 24, r.tex_y.astype(np.float32) * 16 + 24, r.tex_ori], axis=1)
        m.enroll_rolled(os.path.join(small, f"{k:07d}.dat"), minu_xyo, mm.des, tex_xyo, r.tex_des)
    t_enroll = time.perf_counter() - t0
    out["enroll_rolled"] = {"prints": n, "seconds": t_enroll, "prints_per_s": n / t_enroll,
                            "what": "one print at a time: GPU PQ encode + .dat write"}
    for k in range(n_big):
        os.link(os.path.join(small, f"{k % n:07d}.dat"), os.path.join(big, f"{k:07d}.dat"))
    for name, d, cnt in (("ingest_small", small, n), ("ingest_large", big, n_big)):
        size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
        best = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            got = m.load_gallery_dir(d)
            best = min(best, time.perf_counter() - t0)
        assert got == cnt
        out[name] = {"files": cnt, "distinct_files": n, "bytes": size, "seconds": best, "templates_per_s": cnt / best,
                     "GB_per_s": size / best / 1e9, "page_cache": "warm"}
    # the mate of latent 0 is file 0 (and every n-th link): rank 1 expected
    lat = os.path.join(tmp, "q0.dat")
    T.write_template(lat, T.synth_latent(0, raws[0]))
    cbp = os.path.join(tmp, "cb.dat")
    T.write_codebook(cbp, cb)
    sd = os.path.join(tmp, "scores") + os.sep
    os.makedirs(sd)
    work = os.path.join(tmp, "cwd")
    os.makedirs(work)
    m.close()
    exe = os.path.join(entry.PKG_DIR, "bin", "match")
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-c", cbp, "-s", sd, "-g", big, "-l", lat], cwd=work, capture_output=True, text=True,
                       env=dict(os.environ, LAFIS_INGEST_TIMING="1"))
    dt = time.perf_counter() - t0
    rows = open(os.path.join(sd, "q0.csv")).read().split("\n")
    out["cli_one2list_large"] = {"files": n_big, "seconds": dt, "matches_per_s": n_big / dt, "rc": r.returncode,
                                 "first_row": rows[1] if len(rows) > 1 else None,
                                 "ingest": [l for l in r.stderr.split("\n") if l.startswith("lafis ingest:")][-1:],
                                 "reported_matching_ms": [l for l in r.stdout.split("\n") if l.startswith("Total matching")][-1:],
                                 "what": "bin/match -l: process start, CUDA context, ingest of the directory, one 1-vs-N "
                                         "search, score file + 24 x 3 correspondence files; wall clock of the process"}
    print(json.dumps(out))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
