"""Dev helper: gallery ingest throughput - N synthetic rolled .dat files in a directory ->
lafis_gallery_load_dir (parallel parse on all host cores + re-layout into HBM), page cache warm; and end-to-end
enrollment (lafis_enroll_rolled: GPU PQ encoder + writer) of the same prints.
usage: python tools/bench_ingest.py [n_files]"""
import os, shutil, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
T = pkg.templates
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cb = T.synthetic_codebook()
m = pkg.Matcher(codebook=cb, device=0)
tmp = tempfile.mkdtemp(prefix="lafis_ingest_")
try:
    raws = [T.synth_rolled_raw(g) for g in range(n)]
    t0 = time.perf_counter()
    for g, r in enumerate(raws):
        mm = r.minu
        minu_xyo = np.stack([mm.x.astype(np.float32), mm.y.astype(np.float32), mm.ori], axis=1)
        tex_xyo = np.stack([r.tex_x.astype(np.float32) * 16 + 24, r.tex_y.astype(np.float32) * 16 + 24, r.tex_ori], axis=1)
        m.enroll_rolled(os.path.join(tmp, f"{g:07d}.dat"), minu_xyo, mm.des, tex_xyo, r.tex_des)
    t_enroll = time.perf_counter() - t0
    size = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp))
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        got = m.load_gallery_dir(tmp)
        best = min(best, time.perf_counter() - t0)
    assert got == n
    print(f"enroll_rolled: {n} prints in {t_enroll:.2f} s = {n / t_enroll:.0f} prints/s (one at a time, GPU PQ encode + file write)")
    print(f"gallery ingest: {n} files, {size / 1e6:.1f} MB in {best * 1e3:.1f} ms = {n / best:.0f} templates/s, "
          f"{size / best / 1e9:.2f} GB/s, {os.cpu_count()} host cores (page cache warm)")
finally:
    shutil.rmtree(tmp, ignore_errors=True)
    m.close()
