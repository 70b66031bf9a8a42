"""Dev helper: turns the artefacts of a profiling run under gpurun into the markdown summaries kept in profiles/.

  python tools/profile_summary.py launches <launches.csv> <out.md> "<command line that produced it>"
  python tools/profile_summary.py full <report.ncu-rep> <out.md> "<command line that produced it>"
  python tools/profile_summary.py traffic <report.ncu-rep> <pairs per launch> <out.json>

(`ncu -i` reads reports without a GPU.)"""
import csv
import io
import json
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def launches(path, out_md, cmd):
    tot, cnt = {}, {}
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("lafis::", "")
        ours = ("tex_", "minu_", "graph_", "fuse_", "topk_", "keys_to_hits", "merge_hits", "fill_empty")
        if not name.startswith(ours):
            continue  # gallery synthesis / ingest of the bench set-up, not part of a match
        tot[name] = tot.get(name, 0) + float(r["Metric Value"])
        cnt[name] = cnt.get(name, 0) + 1
    total = sum(tot.values())
    with open(out_md, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none), matching kernels only\n\n")
        f.write(f"`{cmd}`\n\nPer-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event "
                "numbers of the bench line next to this file (`roofline.kernel_ms_per_step`).\n\n")
        f.write("| kernel | launches | total ns | share of matching time |\n|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| {k} | {cnt[k]} | {int(v)} | {100 * v / total:.1f}% |\n")


def full(rep, out_md, cmd):
    hdr, units, rows = raw_rows(rep)
    stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    with open(out_md, "w") as f:
        f.write("# ncu --set full, one launch per kernel\n\n")
        f.write(f"Command: `{cmd}`\n\nUnder ncu every kernel is serialised and replayed ~40 times: durations are cold-cache, "
                "compare SHARES with the live CUDA-event numbers of the bench line.\n\n")
        for r in rows:
            f.write(f"## {r[hdr.index('Kernel Name')]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
            st = sorted(((float(r[hdr.index(h)] or 0), h.split("issue_stalled_")[1].split("_per_")[0]) for h in stall), reverse=True)
            f.write("| warps stalled per issue-active cycle (top 6) | " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]) + " | |\n\n")


def traffic(rep, pairs, out_json):
    hdr, units, rows = raw_rows(rep)
    names = {"tex_rowmax_kernel": "tex_rowmax_kernel", "minu_sim_kernel": "minu_sim_kernel", "minu_select_kernel": "minu_select_kernel",
             "graph_minu_sparse_kernel": "graph_minu_sparse_kernel", "graph_tex_sparse_kernel": "graph_tex_sparse_kernel(+dense)"}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows:
        k = r[hdr.index("Kernel Name")].split("(")[0]
        if k not in names:
            continue
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            b += float(r[i]) * scale[units[i]]
        out[names[k]] = round(b / pairs, 2)
    with open(out_json, "w") as f:
        json.dump({"source": f"{rep} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch, "
                             f"{pairs} (latent, gallery) pairs)", "bytes_per_pair": out}, f, indent=1)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif mode == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4])
    elif mode == "traffic":
        traffic(sys.argv[2], int(sys.argv[3]), sys.argv[4])
