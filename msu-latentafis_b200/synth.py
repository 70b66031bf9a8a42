"""Synthetic rolled galleries generated directly in HBM (SURVEY.md §8d distributions), for the
benchmark and the large-scale property tests: a 100K gallery is ~6 GB and is never materialised as
.dat files.  torch is used only as the device-memory and RNG plumbing; PQ codes come from the
library's own encoder kernel."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from . import templates as T
from .matcher import Matcher, PackedGallery, pack_rolled


def synth_gallery_device(m: Matcher, n: int, seed: int, head: Sequence[T.FPTemplate] = (), device=None,
                         n_minu=(90, 150), n_tex=(600, 1000)) -> PackedGallery:
    """`head` templates (host objects, e.g. the mates of the benchmark latents) followed by
    n - len(head) templates drawn on the device.  Must be called with the matcher's stream current
    (`torch.cuda.stream(torch.cuda.ExternalStream(m.stream))`)."""
    import torch

    dev = device if device is not None else torch.device("cuda", m.device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_gen = n - len(head)
    assert n_gen >= 0
    hp = pack_rolled(list(head))

    nm = torch.randint(n_minu[0], n_minu[1] + 1, (n_gen,), generator=g, device=dev)
    tot_m = int(nm.sum().item())
    h_m = int(hp.minu_off[-1])
    mx = torch.empty(h_m + tot_m, dtype=torch.int16, device=dev)
    my = torch.empty_like(mx)
    mori = torch.empty(h_m + tot_m, dtype=torch.float32, device=dev)
    mdes = torch.empty((h_m + tot_m, T.DES_LEN), dtype=torch.float32, device=dev)
    if h_m:
        mx[:h_m] = torch.from_numpy(hp.minu_x).to(dev)
        my[:h_m] = torch.from_numpy(hp.minu_y).to(dev)
        mori[:h_m] = torch.from_numpy(hp.minu_ori).to(dev)
        mdes[:h_m] = torch.from_numpy(hp.minu_des).to(dev)
    mx[h_m:] = torch.randint(40, 760, (tot_m,), generator=g, device=dev).to(torch.int16)
    my[h_m:] = torch.randint(40, 728, (tot_m,), generator=g, device=dev).to(torch.int16)
    mori[h_m:] = (torch.rand(tot_m, generator=g, device=dev) * 2 - 1) * math.pi
    step = 1 << 21
    for a in range(0, tot_m, step):
        b = min(tot_m, a + step)
        d = torch.randn((b - a, T.DES_LEN), generator=g, device=dev)
        mdes[h_m + a:h_m + b] = d * (T.DES_NORM / d.norm(dim=1, keepdim=True))
    minu_off = np.concatenate([hp.minu_off.astype(np.int64), h_m + np.cumsum(nm.cpu().numpy().astype(np.int64))])

    # texture: distinct cells of the 47x45 block grid in row-major order
    cells = T.GRID_W * T.GRID_H
    want = torch.randint(n_tex[0], n_tex[1] + 1, (n_gen,), generator=g, device=dev).float()
    tcounts, tcell = [], []
    tstep = 8192
    for a in range(0, n_gen, tstep):
        b = min(n_gen, a + tstep)
        mask = torch.rand((b - a, cells), generator=g, device=dev) < (want[a:b, None] / cells)
        tcounts.append(mask.sum(dim=1))
        tcell.append(mask.nonzero()[:, 1])
    tcounts = torch.cat(tcounts) if tcounts else torch.zeros(0, dtype=torch.int64, device=dev)
    tcell = torch.cat(tcell) if tcell else torch.zeros(0, dtype=torch.int64, device=dev)
    tot_t = int(tcell.numel())
    h_t = int(hp.tex_off[-1])
    tx = torch.empty(h_t + tot_t, dtype=torch.int16, device=dev)
    ty = torch.empty_like(tx)
    tori = torch.empty(h_t + tot_t, dtype=torch.float32, device=dev)
    codes = torch.empty((h_t + tot_t, T.PQ_SUBS), dtype=torch.uint8, device=dev)
    if h_t:
        tx[:h_t] = torch.from_numpy(hp.tex_x).to(dev)
        ty[:h_t] = torch.from_numpy(hp.tex_y).to(dev)
        tori[:h_t] = torch.from_numpy(hp.tex_ori).to(dev)
        codes[:h_t] = torch.from_numpy(hp.tex_codes).to(dev)
    tx[h_t:] = (tcell % T.GRID_W).to(torch.int16)
    ty[h_t:] = (tcell // T.GRID_W).to(torch.int16)
    tori[h_t:] = (torch.rand(tot_t, generator=g, device=dev) - 0.5) * math.pi
    for a in range(0, tot_t, step):
        b = min(tot_t, a + step)
        d = torch.randn((b - a, T.DES_LEN), generator=g, device=dev)
        d = (d * (T.DES_NORM / d.norm(dim=1, keepdim=True))).contiguous()
        m.pq_encode(d.data_ptr(), b - a, codes[h_t + a:h_t + b].data_ptr())
    tex_off = np.concatenate([hp.tex_off.astype(np.int64), h_t + np.cumsum(tcounts.cpu().numpy().astype(np.int64))])
    assert minu_off[-1] < 2 ** 32 and tex_off[-1] < 2 ** 32
    keep = (mx, my, mori, mdes, tx, ty, tori, codes)
    return PackedGallery(minu_off.astype(np.uint32), mx.data_ptr(), my.data_ptr(), mori.data_ptr(), mdes.data_ptr(),
                         tex_off.astype(np.uint32), tx.data_ptr(), ty.data_ptr(), tori.data_ptr(), codes.data_ptr(),
                         None, True, keep)
