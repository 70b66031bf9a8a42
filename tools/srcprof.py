"""Dev helper: per-CUDA-source-line hot spots from an ncu report (needs -lineinfo).
usage: python tools_srcprof.py <report.ncu-rep> <kernel regex> [top]"""
import csv, subprocess, sys, io, os
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
data = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        hdr = r
        iN = r.index('# Samples'); iI = r.index('Instructions Executed')
        iX = r.index('L1 Wavefronts Shared Excessive'); iB = r.index('stall_barrier'); iL = r.index('stall_long_sb')
        iS = r.index('stall_short_sb'); iM = r.index('stall_mio'); iW=r.index('stall_wait'); iMa=r.index('stall_math')
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    try:
        data.append((int(r[iN]), int(r[iI]), f"{fname}:{r[0]}", r[1].strip()[:100], int(r[iX] or 0), int(r[iB] or 0), int(r[iL] or 0), int(r[iS] or 0), int(r[iM] or 0)))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print("total samples", tot, "warp-inst", toti)
print(" smp%  inst%  xs-wavefronts  barrier long_sb short_sb mio")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/tot:5.1f} {100*d[1]/toti:5.1f} {d[4]:>11} {d[5]:>6} {d[6]:>6} {d[7]:>6} {d[8]:>6}  {d[2]:>22} {d[3]}")
