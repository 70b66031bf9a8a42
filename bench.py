#!/usr/bin/env python
"""Benchmark of the 1-vs-N gallery matcher hot path (BASELINE.json metric: gallery matches/sec per latent).

  python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA path (one process per GPU)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference CPU matcher on the host cores

A step is one pass of the hot path: `--latents` latent prints (default 1) scored against the gallery shard resident
on every rank (default: 100,000 synthetic rolled prints on one GPU = BASELINE.json configs[1]; 125,000 per GPU on
several = configs[3] at 8 GPUs, 1,000,000 prints), rank lists of `--topk` entries, and for N > 1 the library's own
exchange - one ncclAllGather of the per-shard rank lists + the device-side merge (lafis_match_sharded_device).
`value` is timed with the latent batch already in HBM and the results left in HBM; `e2e` times the same step through
the public call with the latent batch in pinned host memory (host-to-device copy inside) and the merged rank lists
copied back to the host.  torch.distributed is only the launcher plumbing (barrier, max over ranks, passing the NCCL
id); the data path's collective is issued by liblatentafis_b200.so.

Beside the headline the default run adds (`--no-sub` skips them):
  sub.config3_27_latents    BASELINE configs[2]: 27 latents per step against the same shard
  sub.config5_256_latents   BASELINE configs[4]: 256 latents per step, with the SURVEY §8d parity protocol on probe latents
  cpu_baseline.as_shipped   the reference CLI itself (8 OpenMP threads, per-pair file loads) on a 2,000-file directory
  e2e_cli                   bin/match on the same directory, ingest included
`--profile hard` swaps the i.i.d. synthetic gallery for the correlated one of msu-latentafis_b200/synth.py.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "gallery matches/sec per latent"
UNIT = "matches/s"
KERNEL_NAMES = ["tex_rowmax_kernel", "minu_sim_kernel", "minu_select_kernel", "graph_minu_sparse_kernel",
                "graph_tex_sparse_kernel(+dense)", "fuse+topk kernels", "minu_select_slow_kernel", "graph_minu_dense_kernel"]


def load_codebook(T):
    """The shipped codebook when a reference checkout is reachable, else the copy recorded in the golden
    fixture (tests/golden/make_golden.py), else a synthetic one."""
    p = T.find_codebook()
    if p:
        return T.load_codebook(p), "shipped"
    g = os.path.join(ROOT, "tests", "golden", "golden_small.npz")
    if os.path.isfile(g):
        return np.load(g)["codebook"], "shipped (golden fixture copy)"
    return T.synthetic_codebook(), "synthetic"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                      "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU side: the reference matcher (oracle/_ref) or the oracle port on a bounded sample
# --------------------------------------------------------------------------------------------------
def _refbind():
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)
    os.environ.setdefault("OMP_STACKSIZE", "32M")
    import refbind
    return refbind


class CpuChecker:
    """The reference matcher (oracle/_ref, else the plain-C oracle port) holding a set of rolled templates in RAM."""

    def __init__(self, T, cb, rolled):
        self.T, self.cb = T, cb
        self.refbind = _refbind()
        self.tmp = tempfile.mkdtemp(prefix="lafis_cpu_")
        self.kind = "reference" if self.refbind.available() else "port"
        if self.kind == "reference":
            cbp = os.path.join(self.tmp, "codebook.dat")
            T.write_codebook(cbp, cb)
            self.R = self.refbind.RefMatcher(cbp)
            self.handles = []
            p = os.path.join(self.tmp, "r.dat")
            for r in rolled:
                T.write_template(p, r)
                self.handles.append(self.R.load_rolled(p)[0])
        else:
            self.ob = entry.load_oracle()  # plain-C port, single thread
            self.handles = [self.ob.OracleRolled(r) for r in rolled]

    def score(self, latent, threads: int, repeats: int = 1, subset=None):
        """-> (matches/s, scores, threads used)"""
        hs = self.handles if subset is None else [self.handles[i] for i in subset]
        if self.kind == "reference":
            lp = os.path.join(self.tmp, "latent.dat")
            self.T.write_template(lp, latent)
            lh, _ = self.R.load_latent(lp)
            best = float("inf")
            for _ in range(repeats):
                t0 = time.perf_counter()
                _, fin, _ = self.R.score_gallery(lh, hs, threads)
                best = min(best, time.perf_counter() - t0)
            return len(hs) / best, fin, threads
        OL = self.ob.OracleLatent(latent, self.cb)
        t0 = time.perf_counter()
        fin = np.array([self.ob.score_pair(OL, r)[2] for r in hs], np.float32)
        return len(hs) / (time.perf_counter() - t0), fin, 1

    def close(self):
        if self.kind == "reference":
            self.R.close()
        shutil.rmtree(self.tmp, ignore_errors=True)


def host_sample(T, cb, n, n_latents=1, encoder=None, profile="iid"):
    """n host-side synthetic rolled prints (seeds 1000+g) and latents mated to the first n_latents of them.
    `encoder` (descriptors -> PQ codes) replaces the numpy nearest-centroid search when given."""
    raws = [T.synth_rolled_raw(g) for g in range(n)]
    if profile != "iid":
        from msu_latentafis_b200.synth import harden_raw
        raws = [harden_raw(r, g, profile) for g, r in enumerate(raws)]
    if encoder is None:
        rolled = [T.rolled_from_raw(r, cb) for r in raws]
    else:
        codes = encoder(np.concatenate([r.tex_des for r in raws]))
        rolled, at = [], 0
        for r in raws:
            k = r.tex_x.shape[0]
            rolled.append(T.FPTemplate(h=T.IMG_H, w=T.IMG_W, blkH=50, blkW=48, minu=[r.minu],
                                       tex=[T.TextureTemplate(r.tex_x, r.tex_y, r.tex_ori, codes[at:at + k])]))
            at += k
    latents = [T.synth_latent(q, raws[q]) for q in range(n_latents)]
    return raws, rolled, latents


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package()
    T = pkg.templates
    cb, cb_kind = load_codebook(T)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(64, min(2000, 48 * cores))
    _, rolled, latents = host_sample(T, cb, sample, profile=args.profile)
    chk = CpuChecker(T, cb, rolled)
    vals = []
    for i in range(args.warmup + args.steps):
        v, _, thr = chk.score(latents[0], cores)
        if i >= args.warmup:
            vals.append(v)
    kind = chk.kind
    chk.close()
    value = sample * len(vals) / sum(sample / v for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"1 synthetic latent (80 minutiae x 28 templates, 400 texture points) vs synthetic rolled "
                               f"gallery (~120 minutiae, ~800 texture points each); each step scores a {sample}-template "
                               f"sample with the reference's OpenMP loop", "codebook": cb_kind, "profile": args.profile},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": kind,
                         "sample": f"{sample} gallery templates per step, templates preloaded in RAM"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# file-to-file: the reference CLI as shipped next to bin/match on the same .dat directory
# --------------------------------------------------------------------------------------------------
def cli_legs(T, cb, rolled, latent, n_files: int):
    """BASELINE.md §3 timing (1): the reference CLI `match -ldir` with its hard-coded 8 OpenMP threads and per-pair file
    loads (matcher.cpp:168,173) on a directory of `n_files` rolled .dat files, and bin/match on the same directory
    (ingest, match, score-file write).  -> (as_shipped dict | None, e2e_cli dict | None)"""
    rb = _refbind()
    exe = os.path.join(entry.PKG_DIR, "bin", "match")
    tmp = tempfile.mkdtemp(prefix="lafis_cli_")
    try:
        gdir, ldir = os.path.join(tmp, "gallery"), os.path.join(tmp, "latent")
        os.makedirs(gdir)
        os.makedirs(ldir)
        work = os.path.join(tmp, "cwd")
        os.makedirs(work)
        with open(os.path.join(tmp, "afis.config"), "w") as f:  # the reference CLI aborts without ../afis.config (main.cpp:41-44)
            f.write("{}")
        cbp = os.path.join(tmp, "codebook.dat")
        T.write_codebook(cbp, cb)
        n_files = min(n_files, len(rolled))
        for g in range(n_files):
            T.write_template(os.path.join(gdir, f"r{g:06d}.dat"), rolled[g])
        T.write_template(os.path.join(ldir, "q0.dat"), latent)
        nbytes = sum(os.path.getsize(os.path.join(gdir, f)) for f in os.listdir(gdir))
        shipped = cli = None
        outs = {}
        if os.path.isfile(rb.CLI_PATH):
            sd = os.path.join(tmp, "ref_scores") + os.sep
            os.makedirs(sd)
            env = dict(os.environ, OMP_STACKSIZE="32M")
            t0 = time.perf_counter()
            r = subprocess.run([rb.CLI_PATH, "-c", cbp, "-s", sd, "-g", gdir, "-ldir", ldir], cwd=work, env=env,
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            if r.returncode == 0 and os.path.isfile(os.path.join(sd, "q0.csv")):
                outs["ref"] = open(os.path.join(sd, "q0.csv"), "rb").read()
                shipped = {"value": n_files / dt, "unit": UNIT, "cores": 8, "kind": "reference",
                           "seconds": dt, "sample": f"oracle/_ref/match -ldir (1 latent) -g ({n_files} .dat files, "
                                                    f"{nbytes / 1e6:.0f} MB): 8 OpenMP threads hard-coded "
                                                    "(matcher.cpp:168), every rolled file re-read per pair (:173)"}
        if os.path.isfile(exe):
            sd = os.path.join(tmp, "b200_scores") + os.sep
            os.makedirs(sd)
            env = dict(os.environ, LAFIS_INGEST_TIMING="1")
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-c", cbp, "-s", sd, "-g", gdir, "-ldir", ldir], cwd=work, env=env,
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
            dt = time.perf_counter() - t0
            if r.returncode == 0 and os.path.isfile(os.path.join(sd, "q0.csv")):
                outs["b200"] = open(os.path.join(sd, "q0.csv"), "rb").read()
                ingest = [l for l in r.stderr.split("\n") if l.startswith("lafis ingest:")]
                cli = {"value": n_files / dt, "unit": UNIT, "seconds": dt, "files": n_files, "bytes": nbytes,
                       "what": "bin/match -ldir: process start, CUDA context, codebook, parallel ingest of the directory "
                               "into HBM, match, score file; wall clock of the whole process",
                       "ingest": ingest[-1] if ingest else None,
                       "stamps": [l[len("lafis cli: "):] for l in r.stderr.split("\n") if l.startswith("lafis cli: ")],
                       "score_file_identical_to_reference_cli": (outs.get("ref") == outs["b200"]) if "ref" in outs else None}
        return shipped, cli
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------------------------------
# the CUDA arm
# --------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the matcher has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's own log (communicator, ranks, transport) must not reach stdout, which carries the one JSON line: it
        # is written to a per-process file and replayed on stderr once the communicators exist
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        nccl_log = None
        if "NCCL_DEBUG_FILE" not in os.environ:
            nccl_log = os.path.join(tempfile.gettempdir(), f"lafis_nccl_{os.getpid()}.log")
            os.environ["NCCL_DEBUG_FILE"] = nccl_log
        dist.init_process_group("nccl", device_id=dev)

    pkg = entry.load_package()
    T = pkg.templates
    from msu_latentafis_b200.synth import synth_gallery_device  # noqa: E402

    cb, cb_kind = load_codebook(T)
    m = pkg.Matcher(codebook=cb, device=local)
    comm = {"world": world, "backend": None}
    if world > 1:
        # the data path's communicator lives in the library (ncclCommInitRank); torch only carries the id
        box = [m.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        m.comm_init(box[0], rank, world)
        v = m.L.lafis_comm_nccl_version()
        comm = {"world": m.comm_world, "rank0": m.comm_rank, "backend": "NCCL (liblatentafis_b200.so: ncclAllGather of rank lists)",
                "nccl_version": v}
        print(f"[lafis] rank {rank}: NCCL communicator ready, nranks {m.comm_world}, version {v}", file=sys.stderr)
        if nccl_log and os.path.isfile(nccl_log):
            with open(nccl_log, errors="replace") as f:
                for ln in f:
                    if "nranks" in ln or "Init COMPLETE" in ln or "NCCL version" in ln:
                        sys.stderr.write(ln)
            sys.stderr.flush()
    ext = torch.cuda.ExternalStream(m.stream, device=dev)
    G = args.gallery_per_gpu or (100000 if world == 1 else 125000)
    K = args.topk
    Qmax = max([args.latents] + ([27, 256] if not args.no_sub else []))

    with torch.cuda.stream(ext):
        n_host = max(Qmax, args.parity_sample, args.cli_files if not args.no_sub else 0) if rank == 0 else Qmax
        raws, head, latents_all = host_sample(T, cb, n_host, Qmax, encoder=m.pq_encode, profile=args.profile)
        gal = synth_gallery_device(m, G, seed=1234 + rank, head=head if rank == 0 else (), device=dev, profile=args.profile)
        m.set_gallery(gal, index_base=rank * G)
        del gal
        torch.cuda.empty_cache()
        gallery_bytes = m.gallery_bytes
        host_hits = torch.empty((Qmax, K, 2), dtype=torch.int32).pin_memory()
        stage_acc = np.zeros(8)

        def make_batch(Q):
            packed = pkg.pack_latents(latents_all[:Q])
            return packed, m.latents_from_packed(packed).make_resident(), m.latents_from_packed(packed)

        def step(lat, Q, to_host: bool):
            d_hits = m.match_sharded_device(lat, K)  # local shard, ncclAllGather + merge when world > 1
            stage_acc[:] += np.array(m.stats()["last_stage_ms"])
            if to_host:
                host_hits[:Q].copy_(_as_tensor(torch, d_hits, (Q, K, 2), dev), non_blocking=True)
                ext.synchronize()

        def barrier():
            ext.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(lat, Q, steps, to_host):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            e0.record(ext)
            for _ in range(steps):
                step(lat, Q, to_host)
            e1.record(ext)
            barrier()
            wall = (time.perf_counter() - t0) * 1e3
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms, wall = (float(x) for x in t.cpu())
            return ms, wall

        # ---- headline: device-resident timing ----
        Q = args.latents
        packed, L_res, L_host = make_batch(Q)
        for _ in range(args.warmup):
            step(L_res, Q, False)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        stage_acc[:] = 0
        l0 = m.stats()["kernel_launches"]
        ms, _ = timed(L_res, Q, args.steps, False)
        launches = m.stats()["kernel_launches"] - l0
        overlapped_ms = stage_acc / max(args.steps, 1)
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end: pinned host latents in, merged rank lists out ----
        for _ in range(min(args.warmup, 3)):
            step(L_host, Q, True)
        ms_e2e, wall_e2e = timed(L_host, Q, args.steps, True)
        final_hits = host_hits[:Q].numpy().copy().view(np.dtype([("score", "<f4"), ("index", "<u4")])).reshape(Q, K)

        # ---- per-kernel durations: two extra steps with every kernel serialised on one stream (in the timed
        #      region the texture chain may overlap the minutiae chain, so its intervals are not exclusive) ----
        m.set_streams(1)
        step(L_res, Q, False)
        stage_acc[:] = 0
        for _ in range(2):
            step(L_res, Q, False)
        stage_ms = stage_acc / 2
        m.set_streams(2)
        barrier()

        # ---- the exchange, checked: the merged device list must equal the host merge of every rank's own list ----
        exchange = None
        local_out = m.match(L_res, K)  # this rank's shard only: scores + local rank list
        if world > 1:
            mine = torch.from_numpy(local_out["hits"].view(np.int32).reshape(Q, K, 2).copy()).to(dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            gathered = np.stack([p.cpu().numpy() for p in parts]).view(pkg.matcher.HIT_DTYPE).reshape(world, Q, K)
            want = m.merge_hits(np.ascontiguousarray(np.transpose(gathered, (1, 0, 2))))
            ok = bool(np.array_equal(want["index"], final_hits["index"]) and
                      np.array_equal(want["score"].view(np.uint32), final_hits["score"].view(np.uint32)))
            flags = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            exchange = {"merged_list_equals_host_merge_on_every_rank": bool(int(flags.item())),
                        "lists": world, "entries_per_list": Q * K, "bytes_per_rank": Q * K * 8}

        # ---- sub-results: BASELINE configs[2] and configs[4] batch sizes on the same shard ----
        sub = {}
        sub_local = {}
        if not args.no_sub:
            for name, Qs, steps_s, warm_s in (("config3_27_latents", 27, 2, 1), ("config5_256_latents", 256, 1, 0)):
                pk, Lr, _ = make_batch(Qs)
                for _ in range(warm_s):
                    step(Lr, Qs, False)
                ms_s, _ = timed(Lr, Qs, steps_s, True)
                hits_s = host_hits[:Qs].numpy().copy().view(pkg.matcher.HIT_DTYPE).reshape(Qs, K)
                sub[name] = {"latents": Qs, "steps": steps_s, "warmup": warm_s, "ms_per_step": ms_s / steps_s,
                             "value": Qs * G * world * steps_s / (ms_s / 1e3), "unit": UNIT,
                             "gallery_total": G * world, "mates_rank1": int(np.sum(hits_s["index"][:, 0] == np.arange(Qs))),
                             "timed": "pinned host latents in, merged rank lists out (e2e)"}
                sub_local[name] = (hits_s, Lr)
        if world > 1:
            dist.barrier()

    if rank != 0:
        _finish(torch, dist, world, ext)  # waits for rank 0's CPU legs
        return

    # ---- rank 0: parity samples + CPU baseline + report ----
    total_pairs = Q * G * world
    value = total_pairs * args.steps / (ms / 1e3)
    e2e_value = total_pairs * args.steps / (ms_e2e / 1e3)
    hbm_peak, peak_kind, sm_max = measured_peaks()
    bytes_per_match = gallery_bytes / G  # this shard's 392*nRm + 24*nRt average
    nLt = int(packed.tex_off[1] - packed.tex_off[0])
    dom = int(np.argmax(stage_ms[:6]))  # dominant kernel: the one with the largest share of the step
    kernel_bytes = kernel_bytes_per_pair(m, G, nLt)
    dom_ms = float(stage_ms[dom])
    achieved = kernel_bytes[dom] * Q * G / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    gathers = Q * G * nLt * kernel_bytes["nRt_mean"] * 16
    lat_minu = float(np.mean([sum(l.minu[i].n for i in (26, 2, 11) if i < len(l.minu)) for l in latents_all[:Q]]))
    smem_bytes_peak = 148 * 128 * sm_max * 1e6
    t_rowmax, t_sim = float(stage_ms[0]) / 1e3, float(stage_ms[1]) / 1e3
    roofline = {
        "bound": "hbm", "kernel": KERNEL_NAMES[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": ncu_traffic(KERNEL_NAMES[dom], Q * G), "peak_source": peak_kind,
        "algorithmic_bytes_per_match_kernel": kernel_bytes[dom], "kernel_ms": dom_ms,
        "step": {"algorithmic_bytes_per_match": bytes_per_match, "achieved": bytes_per_match * value / world / 1e9,
                 "frac": bytes_per_match * value / world / 1e9 / hbm_peak, "note": "per GPU"},
        # the resources that actually bind the two largest kernels (DESIGN.md "Which roofline binds"), from algorithmic
        # operation counts and the same CUDA-event durations: fp32 instruction issue for the dot products (the
        # reference's unfused multiply + add = 2 instructions per MAC on 128 fp32 lanes per SM), the shared-memory
        # data pipe (128 B/clk/SM) for the 1-byte LUT gathers
        "binding": [
            {"kernel": "minu_sim_kernel", "resource": "fp32 instruction issue (2 instr/MAC, 148 SM x 128 lanes x sm_max_mhz)",
             "achieved": (2.0 * lat_minu * 96 * kernel_bytes["nRm_mean"] * Q * G / t_sim / 1e12) if t_sim > 0 else 0,
             "peak": 148 * 128 * sm_max * 1e6 / 1e12, "unit": "T lane-instr/s",
             "frac": (2.0 * lat_minu * 96 * kernel_bytes["nRm_mean"] * Q * G / t_sim / (148 * 128 * sm_max * 1e6)) if t_sim > 0 else 0,
             "note": "latent minutiae of the three selected templates x nRm gallery minutiae x 96-d, padding columns not counted"},
            {"kernel": "tex_rowmax_kernel", "resource": "shared-memory data pipe (148 SM x 128 B/clk x sm_max_mhz)",
             "achieved": (gathers / t_rowmax / 1e12) if t_rowmax > 0 else 0, "peak": smem_bytes_peak / 1e12, "unit": "TB/s",
             "frac": (gathers / t_rowmax / smem_bytes_peak) if t_rowmax > 0 else 0,
             "lookups_per_s": (gathers / t_rowmax) if t_rowmax > 0 else 0,
             "note": "nLt x nRt x 16 one-byte look-ups (16 rows per LDS.128); the exact re-evaluations' uncoalesced loads "
                     "use the same L1 data pipe"}],
        "kernel_ms_per_step": {n: float(v) for n, v in zip(KERNEL_NAMES, stage_ms[:8])},
        "kernel_ms_note": "per-kernel CUDA-event durations from two extra steps with all kernels serialised on one stream "
                          "(lafis_set_streams(1)); in the timed region the same intervals measure (ms): "
                          + ", ".join(f"{n}={float(v):.1f}" for n, v in zip(KERNEL_NAMES, overlapped_ms[:8])),
    }

    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(64, min(2000, 48 * cores))
    sample = min(sample, len(head))
    chk = CpuChecker(T, cb, head)
    v, fin, thr = chk.score(latents_all[0], cores, repeats=3, subset=range(sample))
    cpu = {"value": v, "unit": UNIT, "cores": thr, "kind": chk.kind,
           "sample": f"latent 0 vs the first {sample} gallery templates, preloaded in RAM, OpenMP static,16"}
    sc = local_out["scores"][0][:sample]
    parity = {"checked_pairs": int(sample), "bit_identical": bool(np.array_equal(sc, fin)),
              "max_rel_diff": float(np.max(np.abs(sc - fin) / np.maximum(np.abs(fin), 1e-6))),
              "top_list_mate_rank1": bool(int(final_hits[0]["index"][0]) == 0), "exchange": exchange}

    # SURVEY §8d parity protocol on the batched sub-results: probe latents, the GPU's best `top` of this rank's shard and a
    # random sample of the rest re-scored by the reference; scores bit-identical, top-100 order identical, no sampled
    # non-candidate above the 100th score
    for name, (hits_s, Lr) in sub_local.items():
        Qs = sub[name]["latents"]
        probes = sorted({0, Qs // 2, Qs - 1})
        pk = pkg.pack_latents([latents_all[q] for q in probes])
        with torch.cuda.stream(ext):
            loc = m.match(m.latents_from_packed(pk), 0)["scores"]  # rank 0's shard, [probes][G]
        rng = np.random.default_rng(99)
        top_n, rand_n = args.protocol_top, args.protocol_random
        rep = {"probe_latents": probes, "top_rescored": top_n, "random_rescored": rand_n, "shard": "rank 0",
               "pairs_rescored": 0, "bit_identical": True, "max_rel_diff": 0.0, "top100_identical": True,
               "sampled_noncandidates_above_100th": 0, "checker": chk.kind}
        cache = {}

        def handle_set(idx):
            out = []
            for i in idx:
                i = int(i)
                if i not in cache:
                    cache[i] = head[i] if i < len(head) else m.gallery_template(i)
                out.append(cache[i])
            return out

        for pi, q in enumerate(probes):
            row = loc[pi]
            order = np.lexsort((np.arange(G), -row.astype(np.float64)))
            top = order[:top_n]
            rest = rng.choice(order[top_n:], min(rand_n, G - top_n), replace=False)
            idx = np.concatenate([top, rest])
            c2 = CpuChecker(T, cb, handle_set(idx))
            _, want, _ = c2.score(latents_all[q], cores)
            c2.close()
            got = row[idx]
            rep["pairs_rescored"] += int(len(idx))
            rep["bit_identical"] &= bool(np.array_equal(got, want))
            rep["max_rel_diff"] = max(rep["max_rel_diff"], float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6))))
            ref_order = idx[np.lexsort((idx, -want.astype(np.float64)))][:min(K, top_n)]
            rep["top100_identical"] &= bool(np.array_equal(ref_order, order[:len(ref_order)]))
            rep["sampled_noncandidates_above_100th"] += int(np.sum(want[top_n:] > row[order[min(K, G) - 1]]))
            if world == 1:  # the global list is this shard's list
                rep["top100_identical"] &= bool(np.array_equal(hits_s[q]["index"][:len(ref_order)].astype(np.int64), ref_order))
        sub[name]["parity"] = rep

    shipped = e2e_cli = None
    if not args.no_sub:
        shipped, e2e_cli = cli_legs(T, cb, head, latents_all[0], args.cli_files)
        if shipped:
            cpu["as_shipped"] = shipped
    chk.close()

    st = m.stats()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{Q} synthetic latent(s) (3 x 80 minutiae, {nLt} texture points) vs {G} synthetic rolled "
                               f"prints per GPU (~120 minutiae, ~800 PQ-coded texture points), top-{K} rank lists"
                               + (", ncclAllGather + merge of per-shard lists inside liblatentafis_b200.so" if world > 1 else ""),
                   "latents": Q, "gallery_per_gpu": G, "gallery_total": G * world, "topk": K, "codebook": cb_kind,
                   "profile": args.profile,
                   "l2": f"gallery shard {gallery_bytes / 1e9:.2f} GB is streamed every step (>> 126 MB L2), no flush needed",
                   "mate_rank1": bool(int(final_hits[0]["index"][0]) == 0)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(L_host.nbytes),
                "d2h_bytes_per_step": int(Q * K * 8), "ms_per_step": ms_e2e / args.steps,
                "wall_ms_per_step": wall_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e_cli": e2e_cli,
        "parity": parity,
        "comm": comm,
        "sub": sub or None,
        "exactness": {k: v for k, v in st.items() if k.startswith(("tex_", "minu_", "graph_"))},
    }
    print(json.dumps(line))
    sys.stdout.flush()
    _finish(torch, dist, world, ext)


def _finish(torch, dist, world, ext):
    """Orderly end of a rank: torch's pinned-host allocator records events on the matcher's stream when its blocks are
    released, so the process must not run destructors in arbitrary order (the context's stream may be gone first)."""
    ext.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def ncu_traffic(kernel: str, pairs: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu --set full capture
    (profiles/ncu_traffic.json, bytes per (latent, gallery) pair at the benchmark's sizes), scaled to this launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return float(d["bytes_per_pair"][kernel]) * pairs
    except Exception:
        return None


def _as_tensor(torch, ptr: int, shape, dev):
    """View of library-owned device memory as an int32 tensor (no copy)."""
    n = int(np.prod(shape))

    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 3}

    return torch.as_tensor(_Arr(), device=dev).view(*shape)


def kernel_bytes_per_pair(m, G, nLt):
    """Algorithmic bytes each kernel must move per (latent, gallery) pair, from the resident shard's mean
    template sizes (DESIGN.md "Kernels"): what it reads of the gallery + what it must write."""
    n_probe = min(G, 2000)
    nm = nt = 0
    import ctypes as C
    a, b = C.c_int(0), C.c_int(0)
    for i in range(0, n_probe):
        m.L.lafis_gallery_get_template(m.ctx, i * (G // n_probe), C.addressof(a), None, None, None, None, C.addressof(b),
                                       None, None, None, None)
        nm += a.value
        nt += b.value
    nRm, nRt = nm / n_probe, nt / n_probe
    return {
        0: 16 * nRt + 4 + 6 * nLt,                 # tex_rowmax: PQ codes in, (f32 max, u16 argmax) per latent row out
        1: 384 * nRm + 6 + 3 * 80 * 4 * nRm,       # minu_sim: descriptors in, 3 similarity matrices out
        2: 3 * 80 * 4 * nRm + 3 * (120 * 8 + 4),   # minu_select: similarity matrices in, 3 x top-120 (value, ij) out
        3: 3 * (120 * 8 + 4) + 8 * nRm + 12,       # graph_minu: candidates + minutiae coordinates in, 3 scores out
        4: 6 * nLt + 8 * nRt + 4,                  # graph_tex: row maxima + texture coordinates in, 1 score out
        5: 16 + 4,                                 # fuse: 4 components in, 1 score out
        "nRm_mean": nRm, "nRt_mean": nRt,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--latents", type=int, default=1)
    ap.add_argument("--gallery-per-gpu", type=int, default=0, help="default: 100000 on one GPU, 125000 per GPU on several")
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--profile", default="iid", choices=["iid", "hard", "harder"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--parity-sample", type=int, default=1024)
    ap.add_argument("--no-sub", action="store_true", help="skip the 27- / 256-latent sub-results and the command-line legs")
    ap.add_argument("--cli-files", type=int, default=2000)
    ap.add_argument("--protocol-top", type=int, default=300)
    ap.add_argument("--protocol-random", type=int, default=1500)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
