"""Template / codebook file formats of the MSU-LatentAFIS matcher and synthetic template generators.

Byte layouts follow the reference writers `extraction/descriptor_PQ.py:80-175` (latent) and
`:178-272` (rolled) and the reference readers `matching/matcher.cpp:785-884`, `:886-983`;
the codebook layout follows `matching/matcher.cpp:74-93` / `extraction/descriptor_PQ.py:320-327`.
The synthetic generators implement SURVEY.md §8(d).  Pure numpy — used by tests, the bench and the
CLI tooling; the GPU path never goes through this module for its arithmetic.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

DES_LEN = 96
PQ_SUBS = 16
PQ_CLUSTERS = 256
PQ_SUBDIM = 6
DES_NORM = 1.73  # extraction/descriptor_DR.py:150-152

CODEBOOK_NAME = "codebook_EmbeddingSize_96_stride_16_subdim_6.dat"


@dataclass
class MinutiaeTemplate:
    x: np.ndarray  # int16 [n] pixels
    y: np.ndarray  # int16 [n]
    ori: np.ndarray  # float32 [n] radians
    des: np.ndarray  # float32 [n, des_len]

    @property
    def n(self) -> int:
        return int(self.x.shape[0])


@dataclass
class TextureTemplate:
    x: np.ndarray  # int16 [n] BLOCK coordinates, (px-24)/16
    y: np.ndarray
    ori: np.ndarray  # float32 [n]
    des: np.ndarray  # latent: float32 [n, 96]; rolled: uint8 [n, 16] PQ codes

    @property
    def n(self) -> int:
        return int(self.x.shape[0])


@dataclass
class FPTemplate:
    h: int = 800
    w: int = 768
    blkH: int = 50
    blkW: int = 48
    minu: List[MinutiaeTemplate] = field(default_factory=list)
    tex: List[TextureTemplate] = field(default_factory=list)


# ----------------------------------------------------------------------------------------------
# codebook
# ----------------------------------------------------------------------------------------------
def load_codebook(path: str) -> np.ndarray:
    """u16 nrof_subs, u16 nrof_clusters, u16 sub_dim, then f32[subs][clusters][sub_dim]."""
    with open(path, "rb") as f:
        raw = f.read()
    subs, clusters, sub_dim = struct.unpack_from("<HHH", raw, 0)
    cw = np.frombuffer(raw, dtype="<f4", count=subs * clusters * sub_dim, offset=6)
    return cw.reshape(subs, clusters, sub_dim).copy()


def write_codebook(path: str, codewords: np.ndarray) -> None:
    subs, clusters, sub_dim = codewords.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<HHH", subs, clusters, sub_dim))
        f.write(np.ascontiguousarray(codewords, dtype="<f4").tobytes())


def synthetic_codebook(seed: int = 7) -> np.ndarray:
    """A stand-in codebook with the shipped file's shape and value scale (only used when the real
    98,310-byte codebook is not available, e.g. on the GPU box)."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((PQ_SUBS, PQ_CLUSTERS, PQ_SUBDIM)) * (DES_NORM / np.sqrt(DES_LEN))).astype(np.float32)


def pq_encode(des: np.ndarray, codewords: np.ndarray) -> np.ndarray:
    """TrainedPQEncoder.encode_multi (extraction/descriptor_PQ.py:19-27): per sub-quantizer the index of the nearest
    centroid, first minimum wins.  Same arithmetic as the device encoder (pq_encode_kernel): fp32 squared differences
    accumulated in dimension order, so that the Python tooling and the GPU produce identical codes.  The reference's
    scipy.cluster.vq.vq evaluates |o|^2 + |c|^2 - 2 o.c through BLAS for sub-vectors of 5+ dimensions; the two can
    only disagree where two centroids are equidistant to within fp32 rounding (tests/test_pq_encode.py)."""
    des = np.ascontiguousarray(des, dtype=np.float32)
    n = des.shape[0]
    subs, _, sd = codewords.shape
    codes = np.empty((n, subs), dtype=np.uint8)
    cw = np.ascontiguousarray(codewords, dtype=np.float32)
    for m in range(subs):
        sub = des[:, m * sd:(m + 1) * sd]
        d = np.zeros((n, cw.shape[1]), np.float32)
        for k in range(sd):  # dist = fl(dist + fl(t * t)), t = fl(x - c), k ascending
            t = sub[:, k:k + 1] - cw[m][None, :, k]
            d = d + t * t
        codes[:, m] = np.argmin(d, axis=1).astype(np.uint8)
    return codes


# ----------------------------------------------------------------------------------------------
# .dat writers / readers
# ----------------------------------------------------------------------------------------------
def _u16(a) -> bytes:
    return np.ascontiguousarray(np.asarray(a).astype(np.int64) & 0xFFFF, dtype="<u2").tobytes()


def write_template(path: str, T: Optional[FPTemplate], version: int = 1) -> None:
    """Serialise in the matcher's layout.  A rolled template is recognised by uint8 texture
    descriptors (PQ codes, des_len 16); a latent one carries float32 texture descriptors."""
    with open(path, "wb") as f:
        hdr = [0] * 12
        hdr[0] = version
        f.write(struct.pack("<12H", *hdr))
        if T is None or len(T.minu) == 0:
            f.write(struct.pack("<4H", 0, 0, 0, 0))  # descriptor_PQ.py:92-95 "empty" file
            return
        f.write(struct.pack("<4H", T.h, T.w, min(T.blkH, 50), min(T.blkW, 50)))
        f.write(struct.pack("<B", len(T.minu)))
        for mt in T.minu:
            n = mt.n
            f.write(struct.pack("<H", n))
            if n <= 0:
                continue
            f.write(_u16(mt.x))
            f.write(_u16(mt.y))
            f.write(np.ascontiguousarray(mt.ori, dtype="<f4").tobytes())
            f.write(struct.pack("<H", mt.des.shape[1]))
            f.write(np.ascontiguousarray(mt.des, dtype="<f4").tobytes())
        f.write(struct.pack("<B", len(T.tex)))
        for tt in T.tex:
            n = tt.n
            f.write(struct.pack("<H", n))
            if n <= 0:
                continue
            f.write(_u16(tt.x))
            f.write(_u16(tt.y))
            f.write(np.ascontiguousarray(tt.ori, dtype="<f4").tobytes())
            f.write(struct.pack("<H", tt.des.shape[1]))
            if tt.des.dtype == np.uint8:
                f.write(np.ascontiguousarray(tt.des).tobytes())
            else:
                f.write(np.ascontiguousarray(tt.des, dtype="<f4").tobytes())


def read_template(path: str, latent: bool) -> FPTemplate:
    """Python mirror of the two `load_FP_template` readers, for tooling and tests (the C++ host
    library has its own parser).  Empty minutiae records are skipped like the reference does
    (matcher.cpp:834-836), which shifts later template indices."""
    with open(path, "rb") as f:
        raw = f.read()
    T = FPTemplate(minu=[], tex=[])
    if len(raw) <= (0 if latent else 10):
        return T
    off = 24
    T.h, T.w, T.blkH, T.blkW = struct.unpack_from("<4h", raw, off)
    off += 8
    if off >= len(raw):
        return T
    n_minu_t = raw[off]
    off += 1
    for _ in range(n_minu_t):
        (n,) = struct.unpack_from("<h", raw, off)
        off += 2
        if n <= 0:
            continue
        x = np.frombuffer(raw, "<i2", n, off).copy(); off += 2 * n
        y = np.frombuffer(raw, "<i2", n, off).copy(); off += 2 * n
        ori = np.frombuffer(raw, "<f4", n, off).copy(); off += 4 * n
        (dl,) = struct.unpack_from("<h", raw, off); off += 2
        des = np.frombuffer(raw, "<f4", n * dl, off).reshape(n, dl).copy(); off += 4 * n * dl
        T.minu.append(MinutiaeTemplate(x, y, ori, des))
    if off >= len(raw):
        return T
    n_tex_t = raw[off]
    off += 1
    for _ in range(n_tex_t):
        (n,) = struct.unpack_from("<h", raw, off)
        off += 2
        if n <= 0:
            continue
        x = np.frombuffer(raw, "<i2", n, off).copy(); off += 2 * n
        y = np.frombuffer(raw, "<i2", n, off).copy(); off += 2 * n
        ori = np.frombuffer(raw, "<f4", n, off).copy(); off += 4 * n
        (dl,) = struct.unpack_from("<h", raw, off); off += 2
        if latent:
            des = np.frombuffer(raw, "<f4", n * dl, off).reshape(n, dl).copy(); off += 4 * n * dl
        else:
            des = np.frombuffer(raw, "u1", n * dl, off).reshape(n, dl).copy(); off += n * dl
        T.tex.append(TextureTemplate(x, y, ori, des))
    return T


# ----------------------------------------------------------------------------------------------
# synthetic templates, SURVEY.md §8(d)
# ----------------------------------------------------------------------------------------------
IMG_H, IMG_W = 800, 768
GRID_W, GRID_H = 47, 45  # block grid of the rolled texture template (extraction_rolled.py:113-132)


def _unit_rows(a: np.ndarray) -> np.ndarray:
    return a / np.linalg.norm(a, axis=1, keepdims=True)


@dataclass
class RolledRaw:
    """A synthetic rolled print before PQ encoding (keeps the float texture descriptors so that a
    mated latent can be derived from it)."""
    minu: MinutiaeTemplate
    tex_x: np.ndarray
    tex_y: np.ndarray
    tex_ori: np.ndarray
    tex_des: np.ndarray  # float32 [n,96]


def synth_rolled_raw(g: int, n_minu=None, n_tex=None) -> RolledRaw:
    rng = np.random.default_rng(1000 + g)
    n = int(rng.integers(90, 151)) if n_minu is None else int(n_minu)
    x = rng.integers(40, 760, n).astype(np.int16)
    y = rng.integers(40, 728, n).astype(np.int16)
    ori = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    des = (DES_NORM * _unit_rows(rng.standard_normal((n, DES_LEN)))).astype(np.float32)
    nt = int(rng.integers(600, 1001)) if n_tex is None else int(n_tex)
    cells = np.sort(rng.choice(GRID_W * GRID_H, nt, replace=False))  # row-major order
    ty = (cells // GRID_W).astype(np.int16)
    tx = (cells % GRID_W).astype(np.int16)
    tori = rng.uniform(-np.pi / 2, np.pi / 2, nt).astype(np.float32)
    tdes = (DES_NORM * _unit_rows(rng.standard_normal((nt, DES_LEN)))).astype(np.float32)
    return RolledRaw(MinutiaeTemplate(x, y, ori, des), tx, ty, tori, tdes)


def rolled_from_raw(raw: RolledRaw, codewords: np.ndarray) -> FPTemplate:
    codes = pq_encode(raw.tex_des, codewords)
    return FPTemplate(h=IMG_H, w=IMG_W, blkH=50, blkW=48, minu=[raw.minu],
                      tex=[TextureTemplate(raw.tex_x, raw.tex_y, raw.tex_ori, codes)])


def synth_rolled(g: int, codewords: np.ndarray, n_minu=None, n_tex=None) -> FPTemplate:
    return rolled_from_raw(synth_rolled_raw(g, n_minu, n_tex), codewords)


def synth_latent(q: int, mate: RolledRaw, n_minu: int = 80, n_tex_pts: int = 200,
                 n_minu_templates: int = 28, noise: float = 0.35) -> FPTemplate:
    """Latent derived from gallery mate `mate`: 28 minutiae templates of `n_minu` mate minutiae
    each (rotated, jittered, descriptor-perturbed) and one texture template of `n_tex_pts` mate grid
    points x 2 orientations (extraction_latent.py:200-202)."""
    rng = np.random.default_rng(9000 + q)
    theta = float(rng.uniform(-0.3, 0.3))
    c, s = np.cos(theta), np.sin(theta)
    cx, cy = IMG_W / 2.0, IMG_H / 2.0

    def rot(px, py):
        dx, dy = px - cx, py - cy
        return cx + c * dx - s * dy, cy + s * dx + c * dy

    def perturb(d):
        nz = noise * DES_NORM * _unit_rows(rng.standard_normal(d.shape))
        return (DES_NORM * _unit_rows(d.astype(np.float64) + nz)).astype(np.float32)

    T = FPTemplate(h=IMG_H, w=IMG_W, blkH=50, blkW=48, minu=[], tex=[])
    m = mate.minu
    for _ in range(n_minu_templates):
        k = min(n_minu, m.n)
        sel = np.sort(rng.choice(m.n, k, replace=False))
        rx, ry = rot(m.x[sel].astype(np.float64), m.y[sel].astype(np.float64))
        rx = np.clip(np.rint(rx + rng.normal(0, 2, k)), 0, IMG_W - 1).astype(np.int16)
        ry = np.clip(np.rint(ry + rng.normal(0, 2, k)), 0, IMG_H - 1).astype(np.int16)
        ori = (m.ori[sel] - theta + rng.normal(0, 0.05, k)).astype(np.float32)
        T.minu.append(MinutiaeTemplate(rx, ry, ori, perturb(m.des[sel])))
    nt = mate.tex_x.shape[0]
    k = min(n_tex_pts, nt)
    sel = np.sort(rng.choice(nt, k, replace=False))
    px = mate.tex_x[sel].astype(np.float64) * 16 + 24
    py = mate.tex_y[sel].astype(np.float64) * 16 + 24
    rx, ry = rot(px, py)
    bx = np.clip(np.rint((rx - 24) / 16), 0, 49).astype(np.int16)
    by = np.clip(np.rint((ry - 24) / 16), 0, 49).astype(np.int16)
    ori = (mate.tex_ori[sel] - theta).astype(np.float32)
    # two virtual minutiae per grid point: (ori, ori+pi), each with its own descriptor
    tx = np.repeat(bx, 2)
    ty = np.repeat(by, 2)
    tori = np.stack([ori, ori + np.float32(np.pi)], axis=1).reshape(-1).astype(np.float32)
    tdes = perturb(np.repeat(mate.tex_des[sel], 2, axis=0))
    T.tex.append(TextureTemplate(tx, ty, tori, tdes))
    return T


def find_codebook() -> Optional[str]:
    """The shipped codebook if a reference checkout is reachable (never on the GPU box)."""
    for p in (os.environ.get("LAFIS_CODEBOOK", ""),
              os.path.join("/root/reference/matching", CODEBOOK_NAME)):
        if p and os.path.isfile(p):
            return p
    return None
