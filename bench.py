#!/usr/bin/env python
"""Benchmark of the 1-vs-N gallery matcher hot path (BASELINE.json metric: gallery matches/sec per latent).

  python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA path (one process per GPU)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference CPU matcher on the host cores

A step is one pass of the hot path: `--latents` latent prints (default 1, BASELINE.json configs[1]) scored
against the gallery shard resident on every rank (default 100,000 synthetic rolled prints per GPU, SURVEY.md
§8d distributions), rank lists of `--topk` entries, and for N > 1 one NCCL all-gather of the per-shard rank lists
followed by the merge.  `value` is timed with the latent batch already in HBM and the results left in HBM;
`e2e` times the same step through the public call with the latent batch in pinned host memory (host-to-device
copy inside) and the merged rank lists copied back to the host.
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
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "gallery matches/sec per latent"
UNIT = "matches/s"
BYTES_PER_MATCH_NOMINAL = 66240  # 392*nRm + 24*nRt at nRm=120, nRt=800


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
def cpu_reference(T, cb, latent, rolled, threads: int, repeats: int = 1):
    """Times the reference's preloaded OpenMP loop (matcher.cpp:168-190 schedule) on `rolled`.
    -> (matches/s, kind, scores, threads)"""
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)
    os.environ.setdefault("OMP_STACKSIZE", "32M")
    import refbind
    if refbind.available():
        tmp = tempfile.mkdtemp(prefix="lafis_cpu_")
        try:
            cbp = os.path.join(tmp, "codebook.dat")
            T.write_codebook(cbp, cb)
            R = refbind.RefMatcher(cbp)
            lp = os.path.join(tmp, "latent.dat")
            T.write_template(lp, latent)
            lh, _ = R.load_latent(lp)
            hs = []
            for i, r in enumerate(rolled):
                p = os.path.join(tmp, f"r{i}.dat")
                T.write_template(p, r)
                hs.append(R.load_rolled(p)[0])
                os.unlink(p)
            best = float("inf")
            for _ in range(repeats):
                t0 = time.perf_counter()
                rc, fin, _ = R.score_gallery(lh, hs, threads)
                best = min(best, time.perf_counter() - t0)
            R.close()
            return len(rolled) / best, "reference", fin, threads
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    ob = entry.load_oracle()  # plain-C port, single thread
    OL = ob.OracleLatent(latent, cb)
    OR = [ob.OracleRolled(r) for r in rolled]
    t0 = time.perf_counter()
    fin = np.array([ob.score_pair(OL, r)[2] for r in OR], np.float32)
    return len(rolled) / (time.perf_counter() - t0), "port", fin, 1


def host_sample(T, cb, n, n_latents=1, encoder=None):
    """n host-side synthetic rolled prints (seeds 1000+g) and latents mated to the first n_latents of them.
    `encoder` (descriptors -> PQ codes) replaces the numpy nearest-centroid search when given."""
    raws = [T.synth_rolled_raw(g) for g in range(n)]
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
    _, rolled, latents = host_sample(T, cb, sample)
    vals = []
    for i in range(args.warmup + args.steps):
        v, kind, _, thr = cpu_reference(T, cb, latents[0], rolled, cores)
        if i >= args.warmup:
            vals.append(v)
    value = sample * len(vals) / sum(sample / v for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"1 synthetic latent (80 minutiae x 28 templates, 400 texture points) vs synthetic rolled "
                               f"gallery (~120 minutiae, ~800 texture points each); each step scores a {sample}-template "
                               f"sample with the reference's OpenMP loop", "codebook": cb_kind},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": kind,
                         "sample": f"{sample} gallery templates per step, templates preloaded in RAM"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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
        os.environ["NCCL_DEBUG"] = "NONE"  # VERSION/WARN/INFO put "NCCL version ..." and more on stdout before the JSON line
        dist.init_process_group("nccl", device_id=dev)

    pkg = entry.load_package()
    T = pkg.templates
    from msu_latentafis_b200.synth import synth_gallery_device  # noqa: E402

    cb, cb_kind = load_codebook(T)
    m = pkg.Matcher(codebook=cb, device=local)
    ext = torch.cuda.ExternalStream(m.stream, device=dev)
    Q, G, K = args.latents, args.gallery_per_gpu, args.topk

    with torch.cuda.stream(ext):
        n_host = max(Q, args.parity_sample if rank == 0 else 0)
        raws, head, latents_all = host_sample(T, cb, n_host if rank == 0 else Q, Q, encoder=m.pq_encode)
        latents = latents_all[:Q]
        gal = synth_gallery_device(m, G, seed=1234 + rank, head=head if rank == 0 else (), device=dev)
        m.set_gallery(gal, index_base=rank * G)
        del gal
        torch.cuda.empty_cache()
        packed = pkg.pack_latents(latents)
        L_res = m.latents_from_packed(packed).make_resident()
        L_host = m.latents_from_packed(packed)
        gallery_bytes = m.gallery_bytes

        gathered = torch.empty((world, Q, K, 2), dtype=torch.int32, device=dev)
        merged = torch.empty((Q, K, 2), dtype=torch.int32, device=dev)
        host_hits = torch.empty((Q, K, 2), dtype=torch.int32).pin_memory()
        stage_acc = np.zeros(8)
        launches = [0]

        def step(lat, to_host: bool):
            d_hits, _ = m.match_device(lat, K)
            stage_acc[:] += np.array(m.stats()["last_stage_ms"])
            if world > 1:
                local_hits = _as_tensor(torch, d_hits, (Q, K, 2), dev)
                dist.all_gather_into_tensor(gathered.view(-1), local_hits.reshape(-1))
                m.merge_hits_device(gathered.data_ptr(), Q, world, K, merged.data_ptr())
                src = merged
            else:
                src = _as_tensor(torch, d_hits, (Q, K, 2), dev)
            if to_host:
                host_hits.copy_(src, non_blocking=True)
                ext.synchronize()
            return src

        def barrier():
            ext.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-resident timing ----
        for _ in range(args.warmup):
            step(L_res, False)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        stage_acc[:] = 0
        l0 = m.stats()["kernel_launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(args.steps):
            last = step(L_res, False)
        e1.record(ext)
        barrier()
        ms = e0.elapsed_time(e1)
        launches[0] = m.stats()["kernel_launches"] - l0
        stage_ms = stage_acc / max(args.steps, 1)
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end: pinned host latents in, merged rank lists out ----
        for _ in range(min(args.warmup, 3)):
            step(L_host, True)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record(ext)
        for _ in range(args.steps):
            step(L_host, True)
        f1.record(ext)
        barrier()
        wall_e2e = (time.perf_counter() - t0) * 1e3
        ms_e2e = max(f0.elapsed_time(f1), 0.0)

        # ---- per-kernel durations: two extra steps with every kernel serialised on one stream (in the timed
        #      region the texture chain overlaps the minutiae chain, so its intervals are not exclusive) ----
        overlapped_ms = stage_ms.copy()
        m.set_streams(1)
        step(L_res, False)
        stage_acc[:] = 0
        for _ in range(2):
            step(L_res, False)
        stage_ms = stage_acc / 2
        m.set_streams(2)
        barrier()

        if world > 1:
            t = torch.tensor([ms, ms_e2e, wall_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_e2e, wall_e2e = (float(x) for x in t.cpu())
        final_hits = host_hits.numpy().copy().view(np.dtype([("score", "<f4"), ("index", "<u4")])).reshape(Q, K)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0: parity sample + CPU baseline + report ----
    total_pairs = Q * G * world
    value = total_pairs * args.steps / (ms / 1e3)
    e2e_value = total_pairs * args.steps / (ms_e2e / 1e3)
    hbm_peak, peak_kind, sm_max = measured_peaks()
    bytes_per_match = gallery_bytes / G  # this shard's 392*nRm + 24*nRt average
    nLt = int(packed.tex_off[1] - packed.tex_off[0])
    # dominant kernel: the one with the largest share of the step
    names = ["tex_rowmax_kernel", "minu_sim_kernel", "minu_select_kernel", "graph_minu_sparse_kernel",
             "graph_tex_sparse_kernel(+dense)", "fuse+topk kernels", "minu_select_slow_kernel", "graph_minu_dense_kernel"]
    dom = int(np.argmax(stage_ms[:6]))
    kernel_bytes = kernel_bytes_per_pair(m, G, nLt)
    dom_ms = float(stage_ms[dom])
    achieved = kernel_bytes[dom] * Q * G / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    tex_points_mean = kernel_bytes["nRt_mean"]
    gathers = Q * G * nLt * tex_points_mean * 16
    lat_minu = float(np.mean([sum(l.minu[i].n for i in (26, 2, 11) if i < len(l.minu)) for l in latents]))
    smem_peak = 148 * 32 * sm_max * 1e6
    roofline = {
        "bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": ncu_traffic(names[dom], Q * G), "peak_source": peak_kind,
        "algorithmic_bytes_per_match_kernel": kernel_bytes[dom], "kernel_ms": dom_ms,
        "step": {"algorithmic_bytes_per_match": bytes_per_match, "achieved": bytes_per_match * value / 1e9,
                 "frac": bytes_per_match * value / 1e9 / hbm_peak},
        "smem_gather": {"kernel": "tex_rowmax_kernel",
                        "row_gathers_per_s": gathers / (float(stage_ms[0]) / 1e3) if stage_ms[0] > 0 else 0,
                        "fp32_gather_roofline_per_s": smem_peak,
                        "frac_of_fp32_gather_roofline": (gathers / (float(stage_ms[0]) / 1e3) / smem_peak) if stage_ms[0] > 0 else 0,
                        "smem_bandwidth_frac": (gathers * 1 / (float(stage_ms[0]) / 1e3) / (smem_peak * 4)) if stage_ms[0] > 0 else 0,
                        "note": "nLt*nRt*16 (row, column, sub-quantizer) look-ups per pair; the fp32 formulation of "
                                "SURVEY.md 8d is bounded by 148 SM x 32 banks x sm_max_mhz 4-byte gathers/s; this kernel "
                                "gathers 1-byte quantised entries, 16 rows per LDS.128 (smem_bandwidth_frac = bytes moved / "
                                "128 B/clk/SM); it is bound by instruction issue (ncu: 78 % issue-active), not by the "
                                "shared-memory pipe"},
        # the resources that actually bind the two largest kernels (DESIGN.md "Which roofline binds"), from algorithmic
        # operation counts and the same CUDA-event durations: fp32 instruction issue for the dot products (the
        # reference's unfused multiply + add = 2 instructions per MAC on 128 fp32 lanes per SM), the shared-memory
        # data pipe (128 B/clk/SM) for the 1-byte LUT gathers
        "binding": [
            {"kernel": "minu_sim_kernel", "resource": "fp32 instruction issue (2 instr/MAC, 148 SM x 128 lanes x sm_max_mhz)",
             "achieved": (2.0 * lat_minu * 96 * kernel_bytes["nRm_mean"] * Q * G / (float(stage_ms[1]) / 1e3) / 1e12) if stage_ms[1] > 0 else 0,
             "peak": 148 * 128 * sm_max * 1e6 / 1e12, "unit": "T lane-instr/s",
             "frac": (2.0 * lat_minu * 96 * kernel_bytes["nRm_mean"] * Q * G / (float(stage_ms[1]) / 1e3) / (148 * 128 * sm_max * 1e6)) if stage_ms[1] > 0 else 0,
             "note": "latent minutiae of the three selected templates x nRm gallery minutiae x 96-d, padding columns not counted"},
            {"kernel": "tex_rowmax_kernel", "resource": "shared-memory data pipe (148 SM x 128 B/clk x sm_max_mhz)",
             "achieved": (gathers / (float(stage_ms[0]) / 1e3) / 1e12) if stage_ms[0] > 0 else 0,
             "peak": smem_peak * 4 / 1e12, "unit": "TB/s",
             "frac": (gathers / (float(stage_ms[0]) / 1e3) / (smem_peak * 4)) if stage_ms[0] > 0 else 0,
             "note": "nLt x nRt x 16 one-byte look-ups; the exact re-evaluations' uncoalesced loads use the same L1 data "
                     "pipe (ncu: 77 % busy in total)"}],
        "kernel_ms_per_step": {n: float(v) for n, v in zip(names, stage_ms[:8])},
        "kernel_ms_note": "per-kernel CUDA-event durations from two extra steps with all kernels serialised on one stream "
                          "(lafis_set_streams(1)); with several pipeline chunks the timed region overlaps the texture chain with the minutiae chain on "
                          "two streams, where the same intervals measure (ms): "
                          + ", ".join(f"{n}={float(v):.1f}" for n, v in zip(names, overlapped_ms[:8])),
    }

    cpu = None
    parity = None
    if world == 1:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or max(64, min(2000, 48 * cores))
        sample = min(sample, len(head)) if len(head) >= 64 else sample
        rolled_s = head[:sample] if len(head) >= sample else host_sample(T, cb, sample)[1]
        v, kind, fin, thr = cpu_reference(T, cb, latents[0], rolled_s, cores, repeats=3)
        cpu = {"value": v, "unit": UNIT, "cores": thr, "kind": kind,
               "sample": f"latent 0 vs the first {len(rolled_s)} gallery templates, preloaded in RAM, OpenMP static,16"}
        if len(head) >= len(rolled_s):
            sc = m.match(L_res, 0)["scores"][0][:len(rolled_s)]
            parity = {"checked_pairs": int(len(rolled_s)), "bit_identical": bool(np.array_equal(sc, fin)),
                      "max_rel_diff": float(np.max(np.abs(sc - fin) / np.maximum(np.abs(fin), 1e-6)))}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{Q} synthetic latent(s) (3 x 80 minutiae, {nLt} texture points) vs {G} synthetic rolled "
                               f"prints per GPU (~120 minutiae, ~800 PQ-coded texture points), top-{K} rank lists"
                               + (", NCCL all-gather + merge of per-shard lists" if world > 1 else ""),
                   "latents": Q, "gallery_per_gpu": G, "gallery_total": G * world, "topk": K, "codebook": cb_kind,
                   "l2": f"gallery shard {gallery_bytes / 1e9:.2f} GB is streamed every step (>> 126 MB L2), no flush needed",
                   "mate_rank1": bool(int(final_hits[0]["index"][0]) == 0)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(L_host.nbytes),
                "d2h_bytes_per_step": int(Q * K * 8), "ms_per_step": ms_e2e / args.steps,
                "wall_ms_per_step": wall_e2e / args.steps},
        "gpu_launches": int(launches[0]),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "exactness": {k: v for k, v in m.stats().items() if k.startswith(("tex_", "minu_"))},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    # mean sizes from the algorithmic byte count: bytes = 392*nRm + 24*nRt; probe a sample of templates instead
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
    ap.add_argument("--gallery-per-gpu", type=int, default=100000)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--parity-sample", type=int, default=1024)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
