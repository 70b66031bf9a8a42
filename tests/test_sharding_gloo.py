"""World-size-2 run of the exchange step on CPU (gloo): shard split, all-gather of per-shard rank lists,
merge — checked against the rank list of the unsharded score matrix."""
import os
import socket

import numpy as np
import pytest

from helpers import rank_list


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scores, k, outdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    import __graft_entry__ as entry
    pkg = entry.load_package()
    from msu_latentafis_b200.sharding import allgather_merge_host, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Q, G = scores.shape
    lo, hi = shard_range(G, rank, world)
    local = np.zeros((Q, k), pkg.matcher.HIT_DTYPE)
    for q in range(Q):
        idx = rank_list(scores[q, lo:hi], k, base=lo)
        local[q]["index"][:len(idx)] = idx
        local[q]["score"][:len(idx)] = scores[q, idx]
        local[q]["index"][len(idx):] = 0xFFFFFFFF
        local[q]["score"][len(idx):] = -np.inf
    merged = allgather_merge_host(local)
    np.save(os.path.join(outdir, f"merged_{rank}.npy"), merged)
    dist.destroy_process_group()


@pytest.mark.parametrize("G", [7, 101])
def test_two_rank_gather_and_merge(built, tmp_path, G):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(G)
    Q, k, world = 3, 5, 2
    # impostor-like scores: mostly exact zeros, a few small sums, one mate
    scores = np.where(rng.random((Q, G)) < 0.6, 0.0, rng.uniform(1, 3, (Q, G))).astype(np.float32)
    scores[:, 3] = 250.0
    scores[0, G - 1] = -1.0
    mp.spawn(_worker, args=(world, _free_port(), scores, k, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        merged = np.load(os.path.join(str(tmp_path), f"merged_{r}.npy"))
        for q in range(Q):
            want = rank_list(scores[q], k)
            assert list(merged[q]["index"]) == want
            assert np.array_equal(merged[q]["score"], scores[q, want])


def test_shard_ranges_tile_the_gallery(pkg):
    from msu_latentafis_b200.sharding import shard_range
    for n in (0, 1, 7, 100000, 1000003):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
