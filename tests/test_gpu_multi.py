"""The multi-GPU data path of the library: device-side merge of per-shard rank lists, the sharded
match (NCCL all-gather + merge, score gather to the root) through one process per GPU and through one process
driving several GPUs, and the drivers / command line over a sharded gallery.  Everything is compared bit for bit
with the single-GPU result of the same library and, for the driver files, with the reference CLI's golden output.

Tests that need two devices skip on a single-GPU box (run them with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from helpers import UB_GALLERY, rank_list, write_golden_files

pytestmark = pytest.mark.gpu


def n_gpus() -> int:
    import torch
    return torch.cuda.device_count()


need2 = pytest.mark.skipif("n_gpus() < 2", reason="needs two GPUs")


@pytest.fixture(scope="module")
def matcher(pkg, built, golden):
    m = pkg.Matcher(codebook=golden["codebook"], device=0)
    yield m
    m.close()


def _random_lists(rng, n_lists, Q, k, tie_levels=6):
    """Per-shard rank lists in the all-gather layout [n_lists][Q][k]: sorted (score desc, index asc) inside a list,
    scores drawn from a few levels so that ties across shards are the norm, some lists short (empty slots)."""
    import __graft_entry__ as entry
    dt = entry.load_package().matcher.HIT_DTYPE
    lists = np.zeros((n_lists, Q, k), dt)
    for s in range(n_lists):
        for q in range(Q):
            sc = rng.integers(0, tie_levels, k).astype(np.float32) * 0.5
            idx = np.sort(rng.choice(1000, k, replace=False)).astype(np.uint32) + 1000 * s
            order = sorted(range(k), key=lambda i: (-sc[i], idx[i]))
            lists[s, q]["score"] = sc[order]
            lists[s, q]["index"] = idx[order]
            if (s + q) % 5 == 0:  # a shard smaller than k
                cut = int(rng.integers(0, k))
                lists[s, q]["score"][cut:] = -np.inf
                lists[s, q]["index"][cut:] = 0xFFFFFFFF
    return lists


@pytest.mark.parametrize("n_lists,Q,k", [(8, 5, 100), (2, 1, 24), (3, 7, 33), (8, 2, 512), (1, 3, 10)])
def test_merge_hits_device_equals_host_merge(pkg, matcher, n_lists, Q, k):
    """lafis_merge_hits_device (merge_hits_kernel) against lafis_merge_hits and the (score desc, index asc) order,
    ties across shards and short shards included."""
    import torch
    rng = np.random.default_rng(100 * n_lists + k)
    lists = _random_lists(rng, n_lists, Q, k)
    want = matcher.merge_hits(np.ascontiguousarray(np.transpose(lists, (1, 0, 2))))
    for q in range(Q):  # the host merge itself against a plain sort
        flat = [(float(h["score"]), int(h["index"])) for h in lists[:, q].reshape(-1) if h["index"] != 0xFFFFFFFF]
        ref = sorted(flat, key=lambda t: (-t[0], t[1]))[:k]
        got = [(float(h["score"]), int(h["index"])) for h in want[q] if h["index"] != 0xFFFFFFFF]
        assert got == ref
    d_in = torch.from_numpy(lists.view(np.int32).reshape(n_lists, Q, k, 2).copy()).cuda()
    d_out = torch.zeros((Q, k, 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    matcher.merge_hits_device(d_in.data_ptr(), Q, n_lists, k, d_out.data_ptr())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(pkg.matcher.HIT_DTYPE).reshape(Q, k)
    assert np.array_equal(got["index"], want["index"])
    assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32))


def _synthetic_set(pkg, cb, n=37, seed=900):
    T = pkg.templates
    raws = [T.synth_rolled_raw(seed + g, n_minu=40 + g % 30, n_tex=200 + 7 * g) for g in range(n)]
    rolled = [T.rolled_from_raw(r, cb) for r in raws]
    latents = [T.synth_latent(60 + i, raws[7 * i], n_minu=40, n_tex_pts=110) for i in range(4)]
    return rolled, latents


def test_sharded_match_without_communicator_equals_match(pkg, matcher, golden):
    """world size 1: lafis_match_sharded is lafis_match."""
    rolled, latents = _synthetic_set(pkg, golden["codebook"])
    matcher.set_gallery(pkg.pack_rolled(rolled))
    L = matcher.latents_from_packed(pkg.pack_latents(latents))
    a = matcher.match(L, topk=9)
    b = matcher.match_sharded(L, topk=9, gather_scores=True)
    assert np.array_equal(a["hits"], b["hits"]) and np.array_equal(a["scores"], b["scores"])


def test_empty_shard_returns_empty_lists(pkg, golden):
    """A rank whose slice of the file list is empty (more ranks than files) takes part with empty rank lists."""
    T = pkg.templates
    cb = golden["codebook"]
    m = pkg.Matcher(codebook=cb, device=0)
    empty = pkg.pack_rolled([])
    m.set_gallery(empty, index_base=5)
    L = m.latents_from_packed(pkg.pack_latents([T.synth_latent(1, T.synth_rolled_raw(2, n_minu=30, n_tex=60), n_minu=20, n_tex_pts=30)]))
    out = m.match(L, topk=4, want_scores=False)
    assert (out["hits"]["index"] == 0xFFFFFFFF).all() and np.isneginf(out["hits"]["score"]).all()
    m.close()


def _write_files(pkg, rolled, latents, root):
    T = pkg.templates
    gdir, ldir = os.path.join(root, "g"), os.path.join(root, "l")
    os.makedirs(gdir)
    os.makedirs(ldir)
    gp, lp = [], []
    for i, r in enumerate(rolled):
        gp.append(os.path.join(gdir, f"r{i:03d}.dat"))
        T.write_template(gp[-1], r)
    for i, l in enumerate(latents):
        lp.append(os.path.join(ldir, f"q{i}.dat"))
        T.write_template(lp[-1], l)
    return gdir, ldir, gp, lp


def _dir_bytes(d):
    return {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))}


@need2
def test_group_match_and_drivers_equal_single_gpu(pkg, matcher, golden, tmp_path):
    """MatcherGroup over two devices (one process, a host thread per device, NCCL all-gather + merge, score gather)
    against the single-GPU matcher: rank lists, score rows and every file the two drivers write."""
    cb = golden["codebook"]
    T = pkg.templates
    cbp = os.path.join(str(tmp_path), "cb.dat")
    T.write_codebook(cbp, cb)
    rolled, latents = _synthetic_set(pkg, cb, n=37)  # odd size: unequal shards
    gdir, ldir, gp, lp = _write_files(pkg, rolled, latents, str(tmp_path))
    matcher.load_gallery_files(gp)
    L1 = matcher.load_latents(lp)
    one = matcher.match(L1, topk=12)
    grp = pkg.MatcherGroup(cbp, [0, 1])
    assert len(grp) == 2 and grp.load_gallery_files(gp) == 37
    L2 = grp.load_latents(lp)
    two = grp.match(L2, topk=12)
    assert np.array_equal(one["scores"], two["scores"])
    assert np.array_equal(one["hits"], two["hits"])
    for i in range(len(latents)):
        assert list(two["hits"][i]["index"]) == rank_list(one["scores"][i], 12)
    # fewer files than devices: one shard is empty
    assert grp.load_gallery_files(gp[:1]) == 1
    tiny = grp.match(L2, topk=3)
    assert np.array_equal(tiny["scores"][:, 0], one["scores"][:, 0]) and (tiny["hits"]["index"][:, 1:] == 0xFFFFFFFF).all()
    # drivers: same files, byte for byte
    out = {}
    for name, eng in (("single", matcher), ("group", grp)):
        sd = os.path.join(str(tmp_path), name) + os.sep
        os.makedirs(sd)
        assert eng.List2List_matching(ldir, gdir, sd) == 0
        assert eng.One2List_matching(lp[1], gdir, sd) == 0
        out[name] = _dir_bytes(sd)
    assert out["single"].keys() == out["group"].keys() and len(out["single"]) > 4
    for f in out["single"]:
        assert out["single"][f] == out["group"][f], f
    grp.close()


@need2
def test_cli_two_gpus_writes_the_reference_files(pkg, golden, tmp_path):
    """`match -gpus 2` on the golden set: N-vs-N CSVs equal the reference CLI's (golden fixture), and every file equals
    the single-GPU run's."""
    import __graft_entry__ as entry
    exe = os.path.join(entry.PKG_DIR, "bin", "match")
    gdir, ldir = write_golden_files(golden, str(tmp_path))
    cbp = os.path.join(str(tmp_path), "cb.dat")
    pkg.templates.write_codebook(cbp, golden["codebook"])
    work = tmp_path / "cwd"
    work.mkdir()
    outs = {}
    for name, extra in (("one", []), ("two", ["-gpus", "2"])):
        sd = os.path.join(str(tmp_path), name) + os.sep
        os.makedirs(sd)
        for mode in (["-ldir", ldir], ["-l", os.path.join(ldir, "lA.dat")]):
            r = subprocess.run([exe, "-c", cbp, "-s", sd, "-g", gdir] + mode + extra, cwd=str(work), capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
        outs[name] = _dir_bytes(sd)
    assert outs["one"].keys() == outs["two"].keys()
    for f in outs["one"]:
        assert outs["one"][f] == outs["two"][f], f
    for fname, text in zip(golden["n2n_files"], golden["n2n_text"]):
        if str(fname) == "lA.csv":
            continue  # overwritten by the 1-vs-N run above
        got = outs["two"][str(fname)].decode().replace(gdir, "@G@")
        rows = lambda s: sorted(r for r in s.strip().split("\n") if not any(u in r for u in UB_GALLERY))
        assert rows(got) == rows(str(text)), fname


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as entry
pkg = entry.load_package()
rank, world, root_dir = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
g = np.load(os.path.join(sys.argv[1], "tests", "golden", "golden_small.npz"))
m = pkg.Matcher(codebook=g["codebook"], device=rank)
idp = os.path.join(root_dir, "nccl_id.bin")
if rank == 0:
    with open(idp + ".tmp", "wb") as f:
        f.write(m.comm_unique_id())
    os.rename(idp + ".tmp", idp)
import time
while not os.path.exists(idp):
    time.sleep(0.05)
m.comm_init(open(idp, "rb").read(), rank, world)
gp = sorted(os.path.join(root_dir, "g", f) for f in os.listdir(os.path.join(root_dir, "g")))
lp = sorted(os.path.join(root_dir, "l", f) for f in os.listdir(os.path.join(root_dir, "l")))
m.load_gallery_files(gp, rank, world)
L = m.load_latents(lp)
tot, base, n = m.gallery_total()
out = m.match_sharded(L, topk=12, gather_scores=True, root=1)
np.save(os.path.join(root_dir, f"hits{rank}.npy"), out["hits"])
if out["scores"] is not None:
    np.save(os.path.join(root_dir, f"scores{rank}.npy"), out["scores"])
print(json.dumps({"rank": rank, "total": int(tot), "base": base.tolist(), "n": n.tolist(), "world": m.comm_world}))
m.close()
'''


@need2
def test_one_process_per_gpu_comm_init(pkg, matcher, golden, tmp_path):
    """Two processes, one GPU each (the torchrun arrangement): lafis_comm_unique_id / lafis_comm_init, sharded ingest,
    lafis_match_sharded with the score gather to a non-zero root; every rank holds the global rank lists."""
    cb = golden["codebook"]
    rolled, latents = _synthetic_set(pkg, cb, n=29, seed=1500)
    gdir, ldir, gp, lp = _write_files(pkg, rolled, latents, str(tmp_path))
    matcher.load_gallery_files(sorted(gp))
    one = matcher.match(matcher.load_latents(sorted(lp)), topk=12)
    script = os.path.join(str(tmp_path), "worker.py")
    open(script, "w").write(WORKER)
    env = dict(os.environ, NCCL_DEBUG="WARN")
    procs = [subprocess.Popen([sys.executable, script, ROOT, str(r), "2", str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    for r in range(2):
        hits = np.load(os.path.join(str(tmp_path), f"hits{r}.npy"))
        assert np.array_equal(hits, one["hits"]), r
    assert not os.path.exists(os.path.join(str(tmp_path), "scores0.npy"))
    assert np.array_equal(np.load(os.path.join(str(tmp_path), "scores1.npy")), one["scores"])
    import json
    info = json.loads(outs[0][0].strip().split("\n")[-1])
    assert info["total"] == 29 and info["base"] == [0, 14] and info["n"] == [14, 15] and info["world"] == 2


@need2
def test_group_over_all_visible_gpus(pkg, matcher, golden, tmp_path):
    """MatcherGroup over every visible device (8 on the scaling box): rank lists and score rows equal the single-GPU
    matcher's, with shards of unequal size and, when there are more devices than a multiple allows, tiny shards."""
    cb = golden["codebook"]
    T = pkg.templates
    cbp = os.path.join(str(tmp_path), "cb.dat")
    T.write_codebook(cbp, cb)
    n_dev = n_gpus()
    rolled, latents = _synthetic_set(pkg, cb, n=max(3 * n_dev + 5, 23), seed=2100)  # the latents mate with prints 0, 7, 14, 21
    gdir, ldir, gp, lp = _write_files(pkg, rolled, latents[:2], str(tmp_path))
    matcher.load_gallery_files(gp)
    one = matcher.match(matcher.load_latents(lp), topk=10)
    grp = pkg.MatcherGroup(cbp, list(range(n_dev)))
    assert grp.load_gallery_files(gp) == len(gp)
    two = grp.match(grp.load_latents(lp), topk=10)
    assert np.array_equal(one["scores"], two["scores"]) and np.array_equal(one["hits"], two["hits"])
    grp.close()
