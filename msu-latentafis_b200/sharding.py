"""Gallery sharding across ranks and the one exchange step of the path: an all-gather of every shard's
top-k rank list followed by a merge with the (score desc, gallery index asc) rule.
`torch.distributed` is the plumbing (NCCL between GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .matcher import HIT_DTYPE, LAFIS_OK, LafisError, load_library


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of an n-template gallery owned by `rank` — the same split
    lafis_gallery_load_files applies (floor(n*r/w))."""
    return n * rank // world, n * (rank + 1) // world


def merge_hits_host(gathered: np.ndarray) -> np.ndarray:
    """[n_lists, Q, k] (all-gather layout) -> [Q, k] with lafis_merge_hits."""
    g = np.ascontiguousarray(np.transpose(np.ascontiguousarray(gathered, HIT_DTYPE), (1, 0, 2)))
    Q, n_lists, k = g.shape
    out = np.zeros((Q, k), HIT_DTYPE)
    rc = load_library().lafis_merge_hits(g.ctypes.data, Q, n_lists, k, out.ctypes.data)
    if rc != LAFIS_OK:
        raise LafisError(rc, "lafis_merge_hits")
    return out


def allgather_merge_host(local_hits: np.ndarray, group=None) -> np.ndarray:
    """Every rank contributes its shard's [Q, k] rank list (host memory); every rank gets the merged
    global [Q, k] list.  Used with the gloo backend; the GPU path keeps the lists in HBM
    (bench.py: all_gather_into_tensor + Matcher.merge_hits_device)."""
    import torch
    import torch.distributed as dist

    a = np.ascontiguousarray(local_hits, HIT_DTYPE)
    Q, k = a.shape
    mine = torch.from_numpy(a.view(np.int32).reshape(Q, k, 2).copy())
    world = dist.get_world_size(group)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    gathered = np.stack([p.numpy() for p in parts]).view(HIT_DTYPE).reshape(world, Q, k)
    return merge_hits_host(gathered)
