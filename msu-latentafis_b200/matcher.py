"""Python mirror of the reference's in-process matcher surface `PQ::Matcher`
(matching/matcher.h:34-75) on top of the C ABI of `include/latentafis_b200.h`.

`Matcher(code_file)`, `One2List_matching(latent_file, rolled_dir, score_path)` and
`List2List_matching(latent_dir, rolled_dir, score_path)` keep the reference's names, argument meaning,
return codes and score-file formats (matching/matcher.cpp:31, :96, :216).  The gallery-resident API
(`set_gallery`, `load_gallery_dir`, `match`) is what the drivers are built on.  Everything that
computes goes through `lib/liblatentafis_b200.so`; there is no Python or CPU implementation of any
scoring stage, and construction fails when the library or an sm_100 device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import templates as T

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblatentafis_b200.so")

LAFIS_OK = 0
LAFIS_LATENT_EMPTY = 1
LAFIS_ERR_NO_TEMPLATES = -1
LAFIS_ERR_ARG = -2
LAFIS_ERR_IO = -3
LAFIS_ERR_CODEBOOK = -4
LAFIS_ERR_CUDA = -5
LAFIS_ERR_UNSUPPORTED_SIZE = -6
LAFIS_ERR_LATENT_LAYOUT = -7
LAFIS_ERR_NO_GALLERY = -8

SELECTED = (26, 2, 11)  # matching/matcher.cpp:380

HIT_DTYPE = np.dtype([("score", "<f4"), ("index", "<u4")])


class LafisError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"latentafis_b200 status {status}: {message}")
        self.status = status


class _PackedGallery(C.Structure):
    _fields_ = [("n_templates", C.c_int32), ("minu_off", C.c_void_p), ("minu_x", C.c_void_p), ("minu_y", C.c_void_p),
                ("minu_ori", C.c_void_p), ("minu_des", C.c_void_p), ("tex_off", C.c_void_p), ("tex_x", C.c_void_p),
                ("tex_y", C.c_void_p), ("tex_ori", C.c_void_p), ("tex_codes", C.c_void_p), ("status", C.c_void_p),
                ("on_device", C.c_int32)]


class _PackedLatents(C.Structure):
    _fields_ = [("n_latents", C.c_int32), ("n_minu_templates", C.c_void_p), ("n_tex_templates", C.c_void_p),
                ("minu_off", C.c_void_p), ("minu_x", C.c_void_p), ("minu_y", C.c_void_p), ("minu_ori", C.c_void_p),
                ("minu_des", C.c_void_p), ("tex_off", C.c_void_p), ("tex_x", C.c_void_p), ("tex_y", C.c_void_p),
                ("tex_ori", C.c_void_p), ("tex_des", C.c_void_p)]


class _RolledFeatures(C.Structure):
    _fields_ = [("h", C.c_int), ("w", C.c_int), ("blkH", C.c_int), ("blkW", C.c_int), ("n_minu", C.c_int),
                ("minu_xyo", C.c_void_p), ("minu_des", C.c_void_p), ("n_tex", C.c_int), ("tex_xyo", C.c_void_p),
                ("tex_des", C.c_void_p), ("des_len", C.c_int)]


class _PointSet(C.Structure):
    _fields_ = [("n", C.c_int), ("xyo", C.c_void_p), ("des", C.c_void_p)]


class _LatentFeatures(C.Structure):
    _fields_ = [("h", C.c_int), ("w", C.c_int), ("blkH", C.c_int), ("blkW", C.c_int), ("n_minu_templates", C.c_int),
                ("minu", C.POINTER(_PointSet)), ("n_tex_templates", C.c_int), ("tex", C.POINTER(_PointSet)),
                ("des_len", C.c_int)]


class _CompNetWeights(C.Structure):
    _fields_ = [("weight", C.c_void_p * 4), ("bias", C.c_void_p * 4), ("bn_weight", C.c_void_p * 4),
                ("bn_bias", C.c_void_p * 4), ("bn_mean", C.c_void_p * 4), ("bn_var", C.c_void_p * 4),
                ("bn_eps", C.c_float)]


class _Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pairs_scored", C.c_uint64), ("last_match_ms", C.c_float),
                ("last_stage_ms", C.c_float * 8), ("minu_replays", C.c_uint64), ("tex_replays", C.c_uint64),
                ("tex_queued", C.c_uint64), ("tex_exact", C.c_uint64), ("tex_overflow", C.c_uint64),
                ("tex_templates", C.c_uint64), ("minu_big_jobs", C.c_uint64), ("graph_minu_dense_jobs", C.c_uint64),
                ("graph_tex_dense_jobs", C.c_uint64), ("graph_minu_mid_jobs", C.c_uint64)]


_lib = None

EXPORTS = [
    "lafis_create", "lafis_create_from_codebook", "lafis_destroy", "lafis_last_error", "lafis_version",
    "lafis_gallery_load_dir", "lafis_gallery_load_files", "lafis_gallery_set_packed", "lafis_gallery_size",
    "lafis_gallery_path", "lafis_gallery_status", "lafis_gallery_bytes", "lafis_gallery_get_template",
    "lafis_latents_load_files", "lafis_latents_from_packed", "lafis_latents_count", "lafis_latents_status",
    "lafis_latents_free", "lafis_latents_make_resident", "lafis_latents_bytes", "lafis_match", "lafis_match_device",
    "lafis_correspondences", "lafis_merge_hits", "lafis_merge_hits_device",
    "lafis_one2list_matching", "lafis_list2list_matching", "lafis_forget_gallery_dir", "lafis_pq_encode",
    "lafis_enroll_rolled", "lafis_enroll_latent", "lafis_compnet_load", "lafis_compress_descriptors",
    "lafis_get_stats", "lafis_set_streams", "lafis_stream", "lafis_latents_minu_templates",
    "lafis_comm_unique_id", "lafis_comm_init", "lafis_comm_rank", "lafis_comm_world", "lafis_comm_destroy",
    "lafis_comm_nccl_version", "lafis_gallery_total", "lafis_match_sharded", "lafis_match_sharded_device",
    "lafis_group_create", "lafis_group_destroy", "lafis_group_size", "lafis_group_ctx", "lafis_group_last_error",
    "lafis_group_gallery_load_dir", "lafis_group_gallery_load_files", "lafis_group_gallery_size", "lafis_group_match",
    "lafis_group_one2list_matching", "lafis_group_list2list_matching",
]


def load_library():
    """dlopen the native library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LafisError(LAFIS_ERR_CUDA, f"{LIB_PATH} is missing: run `python msu-latentafis_b200/build.py` "
                                         "(or __graft_entry__.build()); there is no fallback implementation")
    L = C.CDLL(LIB_PATH)
    vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p
    L.lafis_create.argtypes = [cp, ci, C.POINTER(vp)]
    L.lafis_create_from_codebook.argtypes = [vp, ci, ci, ci, ci, C.POINTER(vp)]
    L.lafis_destroy.argtypes = [vp]
    L.lafis_destroy.restype = None
    L.lafis_last_error.argtypes = [vp]
    L.lafis_last_error.restype = cp
    L.lafis_version.restype = cp
    L.lafis_gallery_load_dir.argtypes = [vp, cp, ci, ci]
    L.lafis_gallery_load_files.argtypes = [vp, C.POINTER(cp), ci, ci, ci]
    L.lafis_gallery_set_packed.argtypes = [vp, C.POINTER(_PackedGallery), C.c_uint32]
    L.lafis_gallery_size.argtypes = [vp]
    L.lafis_gallery_path.argtypes = [vp, ci]
    L.lafis_gallery_path.restype = cp
    L.lafis_gallery_status.argtypes = [vp, ci]
    L.lafis_gallery_bytes.argtypes = [vp]
    L.lafis_gallery_bytes.restype = C.c_uint64
    L.lafis_gallery_get_template.argtypes = [vp, ci] + [vp] * 10
    L.lafis_latents_load_files.argtypes = [vp, C.POINTER(cp), ci, C.POINTER(vp)]
    L.lafis_latents_from_packed.argtypes = [vp, C.POINTER(_PackedLatents), C.POINTER(vp)]
    L.lafis_latents_count.argtypes = [vp]
    L.lafis_latents_status.argtypes = [vp, ci]
    L.lafis_latents_bytes.argtypes = [vp]
    L.lafis_latents_bytes.restype = C.c_uint64
    L.lafis_latents_free.argtypes = [vp]
    L.lafis_latents_free.restype = None
    L.lafis_latents_make_resident.argtypes = [vp, vp]
    L.lafis_match.argtypes = [vp, vp, ci, vp, vp, vp]
    L.lafis_match_device.argtypes = [vp, vp, ci, C.POINTER(vp), C.POINTER(vp)]
    L.lafis_correspondences.argtypes = [vp, vp, ci, ci, vp, vp]
    L.lafis_merge_hits.argtypes = [vp, ci, ci, ci, vp]
    L.lafis_merge_hits_device.argtypes = [vp, vp, ci, ci, ci, vp]
    L.lafis_one2list_matching.argtypes = [vp, cp, cp, cp]
    L.lafis_list2list_matching.argtypes = [vp, cp, cp, cp]
    L.lafis_forget_gallery_dir.argtypes = [vp]
    L.lafis_forget_gallery_dir.restype = None
    L.lafis_pq_encode.argtypes = [vp, vp, C.c_int64, vp, ci]
    L.lafis_enroll_rolled.argtypes = [vp, C.POINTER(_RolledFeatures), cp]
    L.lafis_enroll_latent.argtypes = [vp, C.POINTER(_LatentFeatures), cp]
    L.lafis_compnet_load.argtypes = [vp, C.POINTER(_CompNetWeights)]
    L.lafis_compress_descriptors.argtypes = [vp, vp, C.c_int64, vp, ci, ci]
    L.lafis_get_stats.argtypes = [vp, C.POINTER(_Stats)]
    L.lafis_set_streams.argtypes = [vp, ci]
    L.lafis_stream.argtypes = [vp]
    L.lafis_stream.restype = vp
    L.lafis_latents_minu_templates.argtypes = [vp, ci]
    L.lafis_comm_unique_id.argtypes = [vp]
    L.lafis_comm_init.argtypes = [vp, vp, ci, ci]
    L.lafis_comm_rank.argtypes = [vp]
    L.lafis_comm_world.argtypes = [vp]
    L.lafis_comm_destroy.argtypes = [vp]
    L.lafis_comm_destroy.restype = None
    L.lafis_gallery_total.argtypes = [vp, vp, vp]
    L.lafis_match_sharded.argtypes = [vp, vp, ci, vp, ci, vp, ci]
    L.lafis_match_sharded_device.argtypes = [vp, vp, ci, C.POINTER(vp)]
    L.lafis_group_create.argtypes = [cp, vp, ci, C.POINTER(vp)]
    L.lafis_group_destroy.argtypes = [vp]
    L.lafis_group_destroy.restype = None
    L.lafis_group_size.argtypes = [vp]
    L.lafis_group_ctx.argtypes = [vp, ci]
    L.lafis_group_ctx.restype = vp
    L.lafis_group_last_error.argtypes = [vp]
    L.lafis_group_last_error.restype = cp
    L.lafis_group_gallery_load_dir.argtypes = [vp, cp]
    L.lafis_group_gallery_load_files.argtypes = [vp, C.POINTER(cp), ci]
    L.lafis_group_gallery_size.argtypes = [vp]
    L.lafis_group_match.argtypes = [vp, vp, ci, vp, vp]
    L.lafis_group_one2list_matching.argtypes = [vp, cp, cp, cp]
    L.lafis_group_list2list_matching.argtypes = [vp, cp, cp, cp]
    _lib = L
    return L


# --------------------------------------------------------------------------------------------------
# packed host-side containers
# --------------------------------------------------------------------------------------------------
@dataclass
class PackedGallery:
    """Structure-of-arrays gallery (`lafis_packed_gallery`).  Arrays are numpy (host) or raw device
    pointers (ints) when `on_device` is set; offsets/status are always numpy."""
    minu_off: np.ndarray
    minu_x: object
    minu_y: object
    minu_ori: object
    minu_des: object
    tex_off: np.ndarray
    tex_x: object
    tex_y: object
    tex_ori: object
    tex_codes: object
    status: Optional[np.ndarray] = None
    on_device: bool = False
    keepalive: object = None

    @property
    def n(self) -> int:
        return int(self.minu_off.shape[0]) - 1


@dataclass
class PackedLatents:
    n_minu_templates: np.ndarray
    n_tex_templates: np.ndarray
    minu_off: np.ndarray
    minu_x: np.ndarray
    minu_y: np.ndarray
    minu_ori: np.ndarray
    minu_des: np.ndarray
    tex_off: np.ndarray
    tex_x: np.ndarray
    tex_y: np.ndarray
    tex_ori: np.ndarray
    tex_des: np.ndarray

    @property
    def n(self) -> int:
        return int(self.n_minu_templates.shape[0])


def _cat(parts, dtype, width=None):
    if not parts:
        return np.zeros((0,) if width is None else (0, width), dtype)
    return np.ascontiguousarray(np.concatenate(parts), dtype)


def pack_rolled(templates: Sequence[T.FPTemplate]) -> PackedGallery:
    """FPTemplate objects -> packed gallery, keeping what the matcher reads: minutiae template 0 and
    texture template 0 of every print (matching/matcher.cpp:406, :413)."""
    mo, to = [0], [0]
    mx, my, mori, mdes, tx, ty, tori, tc = [], [], [], [], [], [], [], []
    for t in templates:
        if t.minu:
            m = t.minu[0]
            mx.append(m.x); my.append(m.y); mori.append(m.ori); mdes.append(m.des.reshape(-1, T.DES_LEN))
            mo.append(mo[-1] + m.n)
        else:
            mo.append(mo[-1])
        if t.tex:
            x = t.tex[0]
            tx.append(x.x); ty.append(x.y); tori.append(x.ori); tc.append(x.des.reshape(-1, T.PQ_SUBS))
            to.append(to[-1] + x.n)
        else:
            to.append(to[-1])
    return PackedGallery(np.asarray(mo, np.uint32), _cat(mx, np.int16), _cat(my, np.int16), _cat(mori, np.float32),
                         _cat(mdes, np.float32, T.DES_LEN), np.asarray(to, np.uint32), _cat(tx, np.int16),
                         _cat(ty, np.int16), _cat(tori, np.float32), _cat(tc, np.uint8, T.PQ_SUBS))


def pack_latents(templates: Sequence[T.FPTemplate]) -> PackedLatents:
    """FPTemplate objects (non-empty minutiae templates in file order) -> packed latent batch with the
    three selected minutiae templates {26, 2, 11} and texture template 0."""
    nm, nt, mo, to = [], [], [0], [0]
    mx, my, mori, mdes, tx, ty, tori, td = [], [], [], [], [], [], [], []
    for t in templates:
        minu = [m for m in t.minu if m.n > 0]
        tex = [x for x in t.tex if x.n > 0]
        nm.append(len(minu)); nt.append(len(tex))
        for s in SELECTED:
            if s < len(minu):
                m = minu[s]
                mx.append(m.x); my.append(m.y); mori.append(m.ori); mdes.append(m.des.reshape(-1, T.DES_LEN))
                mo.append(mo[-1] + m.n)
            else:
                mo.append(mo[-1])
        if tex:
            x = tex[0]
            tx.append(x.x); ty.append(x.y); tori.append(x.ori); td.append(x.des.reshape(-1, T.DES_LEN))
            to.append(to[-1] + x.n)
        else:
            to.append(to[-1])
    return PackedLatents(np.asarray(nm, np.int32), np.asarray(nt, np.int32), np.asarray(mo, np.uint32),
                         _cat(mx, np.int16), _cat(my, np.int16), _cat(mori, np.float32), _cat(mdes, np.float32, T.DES_LEN),
                         np.asarray(to, np.uint32), _cat(tx, np.int16), _cat(ty, np.int16), _cat(tori, np.float32),
                         _cat(td, np.float32, T.DES_LEN))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


class Latents:
    """Owner of a `lafis_latents` handle."""

    def __init__(self, matcher: "Matcher", handle):
        self.m = matcher
        self.h = handle

    @property
    def n(self) -> int:
        return self.m.L.lafis_latents_count(self.h)

    def status(self, q: int) -> int:
        return self.m.L.lafis_latents_status(self.h, q)

    @property
    def nbytes(self) -> int:
        return int(self.m.L.lafis_latents_bytes(self.h))

    def make_resident(self) -> "Latents":
        self.m._chk(self.m.L.lafis_latents_make_resident(self.m.ctx, self.h))
        return self

    def free(self):
        if self.h:
            self.m.L.lafis_latents_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.m.ctx:
                self.free()
        except Exception:
            pass


def compnet_layers(state):
    """CompNet state_dict (ordered mapping name -> array, or a sequence of arrays in state_dict order) ->
    4 dicts {weight, bias, bn_weight, bn_bias, bn_mean, bn_var}, one per Linear + BatchNorm1d pair:
    layer1.0/.1, layer2.layers.0/.1, layer2.layers.3/.4, layer3.0/.1 (net_compress.py:10-17, :40-50)."""
    vals = list(state.values()) if hasattr(state, "values") else list(state)
    arrs = []
    for v in vals:
        a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        if a.ndim == 0:  # BatchNorm1d.num_batches_tracked
            continue
        arrs.append(a.astype(np.float32))
    if len(arrs) != 24:
        raise ValueError(f"CompNet state has {len(arrs)} tensors, expected 24")
    names = ("weight", "bias", "bn_weight", "bn_bias", "bn_mean", "bn_var")
    layers = [dict(zip(names, arrs[6 * l:6 * l + 6])) for l in range(4)]
    for l, L in enumerate(layers):
        want = (96, 192) if l == 0 else (96, 96)
        if L["weight"].shape != want or any(L[k].shape != (96,) for k in names[1:]):
            raise ValueError(f"CompNet layer {l}: unexpected tensor shapes")
    return layers


class Matcher:
    """`PQ::Matcher` (matching/matcher.h:34-52) on one B200."""

    def __init__(self, code_file: Optional[str] = None, device: int = 0, codebook: Optional[np.ndarray] = None,
                 _borrowed_ctx=None):
        self.L = load_library()
        self.ctx = C.c_void_p()
        self._borrowed = _borrowed_ctx is not None
        if self._borrowed:  # a context owned by a MatcherGroup
            self.ctx = C.c_void_p(_borrowed_ctx)
            self.device = device
            self._keep = None
            return
        if codebook is not None:
            cb = np.ascontiguousarray(codebook, np.float32)
            rc = self.L.lafis_create_from_codebook(cb.ctypes.data, cb.shape[0], cb.shape[1], cb.shape[2], device,
                                                   C.byref(self.ctx))
        else:
            rc = self.L.lafis_create(os.fsencode(code_file), device, C.byref(self.ctx))
        if rc != LAFIS_OK:
            self.ctx = C.c_void_p()
            why = (self.L.lafis_last_error(None) or b"").decode()
            raise LafisError(rc, f"cannot create matcher context: {why} (this library has no CPU path)")
        self.device = device
        self._keep = None

    # ---- reference-named drivers ----
    def One2List_matching(self, latent_template_file: str, rolled_dir: str, score_path: str) -> int:
        return self.L.lafis_one2list_matching(self.ctx, os.fsencode(latent_template_file), os.fsencode(rolled_dir),
                                              os.fsencode(score_path))

    def List2List_matching(self, latent_dir: str, rolled_dir: str, score_path: str) -> int:
        return self.L.lafis_list2list_matching(self.ctx, os.fsencode(latent_dir), os.fsencode(rolled_dir),
                                               os.fsencode(score_path))

    # ---- gallery ----
    def load_gallery_dir(self, rolled_dir: str, shard_rank: int = 0, shard_count: int = 1) -> int:
        self._chk(self.L.lafis_gallery_load_dir(self.ctx, os.fsencode(rolled_dir), shard_rank, shard_count))
        return self.gallery_size

    def load_gallery_files(self, paths: Sequence[str], shard_rank: int = 0, shard_count: int = 1) -> int:
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        self._chk(self.L.lafis_gallery_load_files(self.ctx, arr, len(paths), shard_rank, shard_count))
        return self.gallery_size

    def set_gallery(self, g: PackedGallery, index_base: int = 0) -> int:
        s = _PackedGallery(g.n, _ptr(g.minu_off), _ptr(g.minu_x), _ptr(g.minu_y), _ptr(g.minu_ori), _ptr(g.minu_des),
                           _ptr(g.tex_off), _ptr(g.tex_x), _ptr(g.tex_y), _ptr(g.tex_ori), _ptr(g.tex_codes),
                           _ptr(g.status), 1 if g.on_device else 0)
        self._chk(self.L.lafis_gallery_set_packed(self.ctx, C.byref(s), index_base))
        return self.gallery_size

    @property
    def gallery_size(self) -> int:
        return self.L.lafis_gallery_size(self.ctx)

    @property
    def gallery_bytes(self) -> int:
        return int(self.L.lafis_gallery_bytes(self.ctx))

    def gallery_path(self, i: int) -> str:
        return os.fsdecode(self.L.lafis_gallery_path(self.ctx, i))

    def gallery_status(self, i: int) -> int:
        return self.L.lafis_gallery_status(self.ctx, i)

    def gallery_template(self, i: int) -> T.FPTemplate:
        nm, nt = C.c_int(0), C.c_int(0)
        self._chk(self.L.lafis_gallery_get_template(self.ctx, i, C.addressof(nm), None, None, None, None,
                                                    C.addressof(nt), None, None, None, None))
        mx = np.zeros(nm.value, np.int16); my = np.zeros(nm.value, np.int16); mori = np.zeros(nm.value, np.float32)
        mdes = np.zeros((nm.value, T.DES_LEN), np.float32)
        tx = np.zeros(nt.value, np.int16); ty = np.zeros(nt.value, np.int16); tori = np.zeros(nt.value, np.float32)
        tc = np.zeros((nt.value, T.PQ_SUBS), np.uint8)
        self._chk(self.L.lafis_gallery_get_template(self.ctx, i, C.addressof(nm), _ptr(mx), _ptr(my), _ptr(mori),
                                                    _ptr(mdes), C.addressof(nt), _ptr(tx), _ptr(ty), _ptr(tori), _ptr(tc)))
        out = T.FPTemplate(minu=[], tex=[])
        if nm.value:
            out.minu.append(T.MinutiaeTemplate(mx, my, mori, mdes))
        if nt.value:
            out.tex.append(T.TextureTemplate(tx, ty, tori, tc))
        return out

    # ---- latents ----
    def load_latents(self, paths: Sequence[str]) -> Latents:
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        h = C.c_void_p()
        self._chk(self.L.lafis_latents_load_files(self.ctx, arr, len(paths), C.byref(h)))
        return Latents(self, h)

    def latents_from_packed(self, p: PackedLatents) -> Latents:
        s = _PackedLatents(p.n, _ptr(p.n_minu_templates), _ptr(p.n_tex_templates), _ptr(p.minu_off), _ptr(p.minu_x),
                           _ptr(p.minu_y), _ptr(p.minu_ori), _ptr(p.minu_des), _ptr(p.tex_off), _ptr(p.tex_x),
                           _ptr(p.tex_y), _ptr(p.tex_ori), _ptr(p.tex_des))
        h = C.c_void_p()
        self._chk(self.L.lafis_latents_from_packed(self.ctx, C.byref(s), C.byref(h)))
        return Latents(self, h)

    # ---- the hot path ----
    def match(self, latents: Latents, topk: int = 0, want_scores: bool = True, want_components: bool = False,
              out_hits: Optional[np.ndarray] = None, out_scores: Optional[np.ndarray] = None):
        """-> dict(hits [Q, topk] structured (score, index), scores [Q, G], components [Q, G, 4])"""
        Q, G = latents.n, self.gallery_size
        hits = scores = comps = None
        if topk > 0:
            hits = out_hits if out_hits is not None else np.zeros((Q, topk), HIT_DTYPE)
        if want_scores:
            scores = out_scores if out_scores is not None else np.zeros((Q, G), np.float32)
        if want_components:
            comps = np.zeros((Q, G, 4), np.float32)
        self._chk(self.L.lafis_match(self.ctx, latents.h, topk, _ptr(hits), _ptr(scores), _ptr(comps)))
        return {"hits": hits, "scores": scores, "components": comps}

    def correspondences(self, latents: Latents, q: int, gallery_index: int):
        """Surviving minutiae correspondences of latent q against one gallery template, the reference's
        save_corr output (matcher.cpp:497-505): -> 3 arrays (one per selected minutiae template) of shape
        (n_i, 4) = latent x, latent y, rolled x, rolled y."""
        xy = np.zeros((3, 120, 4), np.int16)
        n = np.zeros(3, np.int32)
        self._chk(self.L.lafis_correspondences(self.ctx, latents.h, q, gallery_index, _ptr(xy), _ptr(n)))
        return [xy[s, :n[s]].copy() for s in range(3)]

    def match_device(self, latents: Latents, topk: int = 0):
        """Scores and rank lists stay in HBM: returns (d_hits, d_scores) raw device pointers."""
        dh, ds = C.c_void_p(), C.c_void_p()
        self._chk(self.L.lafis_match_device(self.ctx, latents.h, topk, C.byref(dh), C.byref(ds)))
        return dh.value, ds.value

    def merge_hits(self, shard_hits: np.ndarray) -> np.ndarray:
        """[Q, n_lists, topk] per-shard rank lists -> [Q, topk] global rank lists."""
        a = np.ascontiguousarray(shard_hits, HIT_DTYPE)
        Q, n_lists, topk = a.shape
        out = np.zeros((Q, topk), HIT_DTYPE)
        rc = self.L.lafis_merge_hits(a.ctypes.data, Q, n_lists, topk, out.ctypes.data)
        if rc != LAFIS_OK:
            raise LafisError(rc, "merge_hits")
        return out

    # ---- multi-GPU: one process per GPU ----
    def comm_unique_id(self) -> bytes:
        """Rank 0: the 128-byte id every rank passes to comm_init (ncclGetUniqueId)."""
        buf = C.create_string_buffer(128)
        rc = self.L.lafis_comm_unique_id(buf)
        if rc != LAFIS_OK:
            raise LafisError(rc, (self.L.lafis_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        """Collective: joins this context to an NCCL communicator of `world` ranks (ncclCommInitRank)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._chk(self.L.lafis_comm_init(self.ctx, buf, rank, world))

    @property
    def comm_world(self) -> int:
        return self.L.lafis_comm_world(self.ctx)

    @property
    def comm_rank(self) -> int:
        return self.L.lafis_comm_rank(self.ctx)

    def gallery_total(self):
        """Collective: (total templates, per-shard bases, per-shard sizes)."""
        w = self.comm_world
        base, n = np.zeros(w, np.uint32), np.zeros(w, np.uint32)
        tot = self.L.lafis_gallery_total(self.ctx, _ptr(base), _ptr(n))
        if tot < 0:
            raise LafisError(tot, self.last_error())
        return tot, base, n

    def match_sharded(self, latents: Latents, topk: int = 0, gather_scores: bool = False, root: int = 0):
        """Collective over the communicator: this rank's shard is scored, the per-shard rank lists are all-gathered
        (ncclAllGather) and merged on the device; with gather_scores the score rows of all shards arrive on `root`.
        -> dict(hits [Q, topk] global lists on every rank, scores [Q, G_total] on root else None)"""
        Q = latents.n
        hits = np.zeros((Q, topk), HIT_DTYPE) if topk > 0 else None
        scores = None
        if gather_scores:
            tot = self.gallery_total()[0]
            if self.comm_rank in (root, -1):
                scores = np.zeros((Q, tot), np.float32)
        self._chk(self.L.lafis_match_sharded(self.ctx, latents.h, topk, _ptr(hits), 1 if gather_scores else 0, _ptr(scores),
                                             root))
        return {"hits": hits, "scores": scores}

    def match_sharded_device(self, latents: Latents, topk: int) -> int:
        """Same exchange, the merged global rank lists stay in HBM: returns the raw device pointer."""
        dh = C.c_void_p()
        self._chk(self.L.lafis_match_sharded_device(self.ctx, latents.h, topk, C.byref(dh)))
        return dh.value

    def merge_hits_device(self, d_gathered: int, n_latents: int, n_lists: int, topk: int, d_out: int) -> None:
        """[n_lists, Q, topk] gathered rank lists in HBM -> [Q, topk]; enqueued on the matcher's stream."""
        self._chk(self.L.lafis_merge_hits_device(self.ctx, d_gathered, n_latents, n_lists, topk, d_out))

    def pq_encode(self, des, n: Optional[int] = None, codes_ptr: Optional[int] = None):
        """PQ-encode descriptors.  numpy [n,96] -> numpy [n,16]; or raw device pointers (des, n, codes_ptr)."""
        if isinstance(des, np.ndarray):
            d = np.ascontiguousarray(des, np.float32)
            codes = np.zeros((d.shape[0], T.PQ_SUBS), np.uint8)
            self._chk(self.L.lafis_pq_encode(self.ctx, d.ctypes.data, d.shape[0], codes.ctypes.data, 0))
            return codes
        self._chk(self.L.lafis_pq_encode(self.ctx, int(des), int(n), int(codes_ptr), 1))
        return None

    def load_compnet(self, state) -> None:
        """Make the descriptor-compression network resident (lafis_compnet_load).  `state` is the CompNet
        state_dict (extraction/models/net_compress.py:33-53) as an ordered mapping or sequence of arrays in
        state_dict order - mapped by POSITION like the reference's load_model (descriptor_DR.py:39-46);
        num_batches_tracked entries are skipped."""
        arrs = compnet_layers(state)
        W = _CompNetWeights()
        keep = []
        for l, layer in enumerate(arrs):
            for name in ("weight", "bias", "bn_weight", "bn_bias", "bn_mean", "bn_var"):
                a = np.ascontiguousarray(layer[name], np.float32)
                keep.append(a)
                getattr(W, name)[l] = a.ctypes.data
        W.bn_eps = 1e-5
        self._chk(self.L.lafis_compnet_load(self.ctx, C.byref(W)))

    def compress_descriptors(self, des, n: Optional[int] = None, out_ptr: Optional[int] = None, normalise: bool = True):
        """CompNet 192 -> 96 + normalisation to 1.73 (descriptor_DR.py:141-152).  numpy [n,192] -> numpy [n,96];
        or raw device pointers (des, n, out_ptr)."""
        if isinstance(des, np.ndarray):
            d = np.ascontiguousarray(des, np.float32)
            if d.ndim != 2 or d.shape[1] != 192:
                raise ValueError("descriptors must be [n, 192]")
            out = np.zeros((d.shape[0], 96), np.float32)
            self._chk(self.L.lafis_compress_descriptors(self.ctx, _ptr(d), d.shape[0], _ptr(out), int(normalise), 0))
            return out
        self._chk(self.L.lafis_compress_descriptors(self.ctx, int(des), int(n), int(out_ptr), int(normalise), 1))
        return None

    def enroll_rolled(self, out_path: str, minu_xyo, minu_des, tex_xyo, tex_des, h: int = 800, w: int = 768,
                      blkH: int = 50, blkW: int = 48) -> None:
        """Tail of the reference's enrollment (descriptor_PQ.py:19-27 + :178-272): PQ-encodes the texture
        descriptors on the GPU and writes the rolled template file.  *_xyo: [n, 3] rows of x, y (pixels), orientation.
        192-d descriptors are first compressed by the resident CompNet (descriptor_DR.py:141-165)."""
        a = [np.ascontiguousarray(v, np.float32) for v in (minu_xyo, minu_des, tex_xyo, tex_des)]
        des_len = 96
        for d in (a[1], a[3]):
            if d.size:
                des_len = int(d.shape[1])
        F = _RolledFeatures(h, w, blkH, blkW, len(a[0]), _ptr(a[0]), _ptr(a[1]), len(a[2]), _ptr(a[2]), _ptr(a[3]),
                            des_len)
        self._chk(self.L.lafis_enroll_rolled(self.ctx, C.byref(F), out_path.encode()))

    def enroll_latent(self, out_path: str, minu_sets, tex_sets, h: int = 800, w: int = 768, blkH: int = 50,
                      blkW: int = 48) -> None:
        """Template2Bin_Byte_latent (descriptor_PQ.py:80-175): minu_sets / tex_sets are sequences of (xyo [n,3], des [n,L])
        pairs in template order, L = 96, or 192 (compressed by the resident CompNet in one device call)."""
        keep = []
        des_len = 96

        def pack(sets):
            nonlocal des_len
            arr = (_PointSet * max(len(sets), 1))()
            for t, (xyo, des) in enumerate(sets):
                a = np.ascontiguousarray(xyo, np.float32).reshape(-1, 3)
                d = np.ascontiguousarray(des, np.float32)
                if d.size:
                    des_len = int(d.shape[1])
                keep.extend([a, d])
                arr[t] = _PointSet(len(a), _ptr(a), _ptr(d))
            return arr

        ma, ta = pack(list(minu_sets)), pack(list(tex_sets))
        F = _LatentFeatures(h, w, blkH, blkW, len(minu_sets), ma, len(tex_sets), ta, des_len)
        self._chk(self.L.lafis_enroll_latent(self.ctx, C.byref(F), out_path.encode()))

    def stats(self) -> dict:
        s = _Stats()
        self.L.lafis_get_stats(self.ctx, C.byref(s))
        return {"kernel_launches": int(s.kernel_launches), "pairs_scored": int(s.pairs_scored),
                "last_match_ms": float(s.last_match_ms), "last_stage_ms": [float(x) for x in s.last_stage_ms],
                "minu_replays": int(s.minu_replays), "tex_replays": int(s.tex_replays), "tex_queued": int(s.tex_queued),
                "tex_exact": int(s.tex_exact), "tex_overflow": int(s.tex_overflow), "tex_templates": int(s.tex_templates),
                "minu_big_jobs": int(s.minu_big_jobs), "graph_minu_dense_jobs": int(s.graph_minu_dense_jobs),
                "graph_tex_dense_jobs": int(s.graph_tex_dense_jobs), "graph_minu_mid_jobs": int(s.graph_minu_mid_jobs)}

    def set_streams(self, n: int) -> None:
        """2: rare-path kernels on a second, high-priority stream (default); 1: all kernels serialised on one stream."""
        self._chk(self.L.lafis_set_streams(self.ctx, n))

    @property
    def stream(self) -> int:
        return int(self.L.lafis_stream(self.ctx) or 0)

    def last_error(self) -> str:
        return (self.L.lafis_last_error(self.ctx) or b"").decode()

    def _chk(self, rc: int):
        if rc != LAFIS_OK:
            raise LafisError(rc, self.last_error())

    def close(self):
        if self.ctx and not self._borrowed:
            self.L.lafis_destroy(self.ctx)
        self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MatcherGroup:
    """`PQ::Matcher` over several B200s of one process: the gallery is sharded over the devices (contiguous index
    ranges), every match ends in one NCCL all-gather of the shards' rank lists + the device-side merge, and the N-vs-N
    drivers gather the score rows to device 0.  Same drivers, same files as `Matcher`."""

    def __init__(self, code_file: str, devices: Sequence[int]):
        self.L = load_library()
        self.g = C.c_void_p()
        dev = np.asarray(list(devices), np.int32)
        rc = self.L.lafis_group_create(os.fsencode(code_file), dev.ctypes.data, len(dev), C.byref(self.g))
        if rc != LAFIS_OK:
            self.g = C.c_void_p()
            raise LafisError(rc, (self.L.lafis_last_error(None) or b"").decode())
        self.members = [Matcher(device=int(d), _borrowed_ctx=self.L.lafis_group_ctx(self.g, i)) for i, d in enumerate(dev)]

    def __len__(self):
        return self.L.lafis_group_size(self.g)

    def _chk(self, rc: int):
        if rc != LAFIS_OK:
            raise LafisError(rc, (self.L.lafis_group_last_error(self.g) or b"").decode())

    def load_gallery_dir(self, rolled_dir: str) -> int:
        self._chk(self.L.lafis_group_gallery_load_dir(self.g, os.fsencode(rolled_dir)))
        return self.gallery_size

    def load_gallery_files(self, paths: Sequence[str]) -> int:
        arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
        self._chk(self.L.lafis_group_gallery_load_files(self.g, arr, len(paths)))
        return self.gallery_size

    @property
    def gallery_size(self) -> int:
        return self.L.lafis_group_gallery_size(self.g)

    def load_latents(self, paths: Sequence[str]) -> Latents:
        return self.members[0].load_latents(paths)

    def latents_from_packed(self, p: PackedLatents) -> Latents:
        return self.members[0].latents_from_packed(p)

    def match(self, latents: Latents, topk: int = 0, want_scores: bool = True):
        Q, G = latents.n, self.gallery_size
        hits = np.zeros((Q, topk), HIT_DTYPE) if topk > 0 else None
        scores = np.zeros((Q, G), np.float32) if want_scores else None
        self._chk(self.L.lafis_group_match(self.g, latents.h, topk, _ptr(hits), _ptr(scores)))
        return {"hits": hits, "scores": scores}

    def One2List_matching(self, latent_template_file: str, rolled_dir: str, score_path: str) -> int:
        return self.L.lafis_group_one2list_matching(self.g, os.fsencode(latent_template_file), os.fsencode(rolled_dir),
                                                    os.fsencode(score_path))

    def List2List_matching(self, latent_dir: str, rolled_dir: str, score_path: str) -> int:
        return self.L.lafis_group_list2list_matching(self.g, os.fsencode(latent_dir), os.fsencode(rolled_dir),
                                                     os.fsencode(score_path))

    def close(self):
        if self.g:
            for m in self.members:
                m.ctx = C.c_void_p()
            self.L.lafis_group_destroy(self.g)
            self.g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
