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


# The "hard" profile (bench.py --profile hard): what real prints do to the fast paths that i.i.d. Gaussian data does
# not - minutiae concentrated in the central ridge area (denser distance-consistency graphs), descriptors drawn from
# a few directions per print plus noise (correlated similarity rows; many PQ columns near each row minimum), and
# near-duplicates of the probes' mates in the gallery (mated-like dense graphs).
HARD = {  # profile name -> (std of the minutiae cloud [px], descriptor directions per print, noise, near-duplicate share)
    "hard": (110.0, 12, 0.45, 0.01),
    "harder": (60.0, 4, 0.25, 0.05),
}
HARD_SIGMA_PX, HARD_CENTRES, HARD_NOISE, HARD_DUP_FRACTION = HARD["hard"]


def _use_profile(profile: str):
    global HARD_SIGMA_PX, HARD_CENTRES, HARD_NOISE, HARD_DUP_FRACTION
    HARD_SIGMA_PX, HARD_CENTRES, HARD_NOISE, HARD_DUP_FRACTION = HARD.get(profile, HARD["hard"])


def harden_raw(raw: T.RolledRaw, seed: int, profile: str = "hard") -> T.RolledRaw:
    """Host-side version of the hard profile for the head templates (the probes' mates)."""
    _use_profile(profile)
    rng = np.random.default_rng(77000 + seed)
    m = raw.minu
    n = m.n
    x = np.clip(np.rint(T.IMG_W / 2 + rng.normal(0, HARD_SIGMA_PX, n)), 40, 759).astype(np.int16)
    y = np.clip(np.rint(T.IMG_H / 2 + rng.normal(0, HARD_SIGMA_PX, n)), 40, 727).astype(np.int16)

    def mix(k):
        c = T._unit_rows(rng.standard_normal((HARD_CENTRES, T.DES_LEN)))
        d = c[rng.integers(0, HARD_CENTRES, k)] + HARD_NOISE * T._unit_rows(rng.standard_normal((k, T.DES_LEN)))
        return (T.DES_NORM * T._unit_rows(d)).astype(np.float32)

    return T.RolledRaw(T.MinutiaeTemplate(x, y, m.ori, mix(n)), raw.tex_x, raw.tex_y, raw.tex_ori, mix(raw.tex_x.shape[0]))


def synth_gallery_device(m: Matcher, n: int, seed: int, head: Sequence[T.FPTemplate] = (), device=None,
                         n_minu=(90, 150), n_tex=(600, 1000), profile: str = "iid") -> PackedGallery:
    """`head` templates (host objects, e.g. the mates of the benchmark latents) followed by
    n - len(head) templates drawn on the device.  Must be called with the matcher's stream current
    (`torch.cuda.stream(torch.cuda.ExternalStream(m.stream))`).  profile "hard": see HARD_* above."""
    import torch
    hard = profile in HARD
    _use_profile(profile)

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
    if hard:
        mx[h_m:] = (T.IMG_W / 2 + HARD_SIGMA_PX * torch.randn(tot_m, generator=g, device=dev)).round().clamp(40, 759).to(torch.int16)
        my[h_m:] = (T.IMG_H / 2 + HARD_SIGMA_PX * torch.randn(tot_m, generator=g, device=dev)).round().clamp(40, 727).to(torch.int16)
    else:
        mx[h_m:] = torch.randint(40, 760, (tot_m,), generator=g, device=dev).to(torch.int16)
        my[h_m:] = torch.randint(40, 728, (tot_m,), generator=g, device=dev).to(torch.int16)
    mori[h_m:] = (torch.rand(tot_m, generator=g, device=dev) * 2 - 1) * math.pi
    step = 1 << 21

    def unit(t):
        return t / t.norm(dim=1, keepdim=True)

    def draw_des(count, owner):
        """`count` descriptors; `owner` [count] = template index of every point (hard profile: the print's centres)."""
        d = torch.randn((count, T.DES_LEN), generator=g, device=dev)
        if hard:
            # centre k of print t = a hash-seeded direction: regenerate from (t, k) instead of storing n x 12 x 96
            k = torch.randint(0, HARD_CENTRES, (count,), generator=g, device=dev)
            key = (owner * HARD_CENTRES + k)
            uniq, inv = torch.unique(key, return_inverse=True)
            cg = torch.Generator(device=dev)
            cg.manual_seed(seed * 1000003 + int(uniq.numel()) + count)
            centres = unit(torch.randn((uniq.numel(), T.DES_LEN), generator=cg, device=dev))
            d = centres[inv] + HARD_NOISE * unit(d)
        return d * (T.DES_NORM / d.norm(dim=1, keepdim=True))

    owner_m = torch.repeat_interleave(torch.arange(n_gen, device=dev), nm) if hard else None
    for a in range(0, tot_m, step):
        b = min(tot_m, a + step)
        mdes[h_m + a:h_m + b] = draw_des(b - a, owner_m[a:b] if hard else None)
    minu_off = np.concatenate([hp.minu_off.astype(np.int64), h_m + np.cumsum(nm.cpu().numpy().astype(np.int64))])
    if hard and len(head) and n_gen:
        # near-duplicates: a share of the generated prints becomes a jittered copy of a head template's minutiae
        # (same count is not required: the first min(n_head, n_gen) minutiae are overwritten)
        n_dup = int(n_gen * HARD_DUP_FRACTION)
        dup_t = torch.randperm(n_gen, generator=g, device=dev)[:n_dup].cpu().numpy()
        gen_off = minu_off[len(head):]
        for t in dup_t:
            src = int(t) % len(head)
            s0, s1 = int(hp.minu_off[src]), int(hp.minu_off[src + 1])
            d0, d1 = int(gen_off[t]), int(gen_off[t + 1])
            k = min(s1 - s0, d1 - d0)
            mx[d0:d0 + k] = (mx[s0:s0 + k].float() + 2.0 * torch.randn(k, generator=g, device=dev)).round().clamp(0, 767).to(torch.int16)
            my[d0:d0 + k] = (my[s0:s0 + k].float() + 2.0 * torch.randn(k, generator=g, device=dev)).round().clamp(0, 799).to(torch.int16)
            mori[d0:d0 + k] = mori[s0:s0 + k] + 0.05 * torch.randn(k, generator=g, device=dev)
            nz = mdes[s0:s0 + k] + 0.35 * T.DES_NORM * unit(torch.randn((k, T.DES_LEN), generator=g, device=dev))
            mdes[d0:d0 + k] = nz * (T.DES_NORM / nz.norm(dim=1, keepdim=True))

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
    owner_t = torch.repeat_interleave(torch.arange(n_gen, device=dev), tcounts) if hard else None
    for a in range(0, tot_t, step):
        b = min(tot_t, a + step)
        d = draw_des(b - a, owner_t[a:b] if hard else None).contiguous()
        m.pq_encode(d.data_ptr(), b - a, codes[h_t + a:h_t + b].data_ptr())
    tex_off = np.concatenate([hp.tex_off.astype(np.int64), h_t + np.cumsum(tcounts.cpu().numpy().astype(np.int64))])
    assert minu_off[-1] < 2 ** 32 and tex_off[-1] < 2 ** 32
    keep = (mx, my, mori, mdes, tx, ty, tori, codes)
    return PackedGallery(minu_off.astype(np.uint32), mx.data_ptr(), my.data_ptr(), mori.data_ptr(), mdes.data_ptr(),
                         tex_off.astype(np.uint32), tx.data_ptr(), ty.data_ptr(), tori.data_ptr(), codes.data_ptr(),
                         None, True, keep)
