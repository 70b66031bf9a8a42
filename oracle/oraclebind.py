"""ctypes bindings for the plain-C oracle (oracle/lafis_oracle.c) — TEST INFRASTRUCTURE ONLY.
Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, nowhere else."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liblafis_oracle.so")

SUBS, CLUSTERS, SUBDIM = 16, 256, 6


class LoTemplate(C.Structure):
    _fields_ = [("n", C.c_int), ("x", C.c_void_p), ("y", C.c_void_p), ("ori", C.c_void_p),
                ("des_len", C.c_int), ("des", C.c_void_p), ("codes", C.c_void_p)]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "lafis_oracle.c")
    if force or not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "_build/liblafis_oracle.so"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.lo_minutiae_score.restype = C.c_float
        L.lo_texture_score.restype = C.c_float
        L.lo_fuse.restype = C.c_float
        L.lo_fuse.argtypes = [C.c_float] * 4
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


class Keep:
    """Owns the numpy buffers a LoTemplate points into."""

    def __init__(self, x, y, ori, des=None, codes=None):
        self.x = np.ascontiguousarray(x, np.int16)
        self.y = np.ascontiguousarray(y, np.int16)
        self.ori = np.ascontiguousarray(ori, np.float32)
        self.des = None if des is None else np.ascontiguousarray(des, np.float32)
        self.codes = None if codes is None else np.ascontiguousarray(codes, np.uint8)
        dl = self.des.shape[1] if self.des is not None else (self.codes.shape[1] if self.codes is not None else 0)
        self.c = LoTemplate(len(self.x), _p(self.x), _p(self.y), _p(self.ori), dl, _p(self.des), _p(self.codes))


def make_table() -> np.ndarray:
    t = np.zeros(2500, np.float32)
    lib().lo_make_table(C.c_void_p(t.ctypes.data))
    return t


def build_lut(des: np.ndarray, codebook: np.ndarray) -> np.ndarray:
    des = np.ascontiguousarray(des, np.float32)
    cb = np.ascontiguousarray(codebook, np.float32)
    n = des.shape[0]
    lut = np.zeros((n, cb.shape[0], cb.shape[1]), np.float32)
    lib().lo_build_lut(C.c_void_p(des.ctypes.data), n, des.shape[1], C.c_void_p(cb.ctypes.data), cb.shape[0],
                       cb.shape[1], cb.shape[2], C.c_void_p(lut.ctypes.data))
    return lut


def texture_similarity(lut: np.ndarray, codes: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, np.uint8)
    nL, nR = lut.shape[0], codes.shape[0]
    sim = np.zeros((nL, nR), np.float32)
    lib().lo_texture_similarity(C.c_void_p(lut.ctypes.data), nL, C.c_void_p(codes.ctypes.data), nR, lut.shape[1],
                                lut.shape[2], codes.shape[1], C.c_void_p(sim.ctypes.data))
    return sim


def texture_initial_corr(sim: np.ndarray, N: int = 200):
    nL, nR = sim.shape
    cap = max(1, min(nL, N))
    v = np.zeros(cap, np.float32); li = np.zeros(cap, np.int32); rj = np.zeros(cap, np.int32)
    k = lib().lo_texture_initial_corr(C.c_void_p(sim.ctypes.data), nL, nR, N, _vp(v), _vp(li), _vp(rj))
    return v[:k], li[:k], rj[:k]


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def minutiae_initial_corr(A: np.ndarray, B: np.ndarray, want_matrices: bool = False):
    A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
    nL, nR = A.shape[0], B.shape[0]
    v = np.zeros(120, np.float32); li = np.zeros(120, np.int32); rj = np.zeros(120, np.int32)
    S = np.zeros((nL, nR), np.float32) if want_matrices else None
    Nm = np.zeros((nL, nR), np.float32) if want_matrices else None
    k = lib().lo_minutiae_initial_corr(_vp(A), nL, _vp(B), nR, A.shape[1], _vp(v), _vp(li), _vp(rj),
                                       _vp(S) if want_matrices else None, _vp(Nm) if want_matrices else None)
    return (v[:k], li[:k], rj[:k]) + ((S, Nm) if want_matrices else ())


def prune(which: str, v, li, rj, L: Keep, R: Keep, table=None):
    v = np.ascontiguousarray(v, np.float32); li = np.ascontiguousarray(li, np.int32); rj = np.ascontiguousarray(rj, np.int32)
    n = len(v)
    ov = np.zeros(max(n, 1), np.float32); oli = np.zeros(max(n, 1), np.int32); orj = np.zeros(max(n, 1), np.int32)
    Lb, Rb = C.byref(L.c), C.byref(R.c)
    if which == "dist_lookup":
        k = lib().lo_prune_dist_lookup(_vp(v), _vp(li), _vp(rj), n, Lb, Rb, _vp(table), _vp(ov), _vp(oli), _vp(orj))
    elif which == "dist_euclid":
        k = lib().lo_prune_dist_euclid(_vp(v), _vp(li), _vp(rj), n, Lb, Rb, _vp(ov), _vp(oli), _vp(orj))
    elif which == "angle":
        k = lib().lo_prune_angle(_vp(v), _vp(li), _vp(rj), n, Lb, Rb, _vp(ov), _vp(oli), _vp(orj))
    else:
        raise ValueError(which)
    return ov[:k], oli[:k], orj[:k]


def std_sort_desc(key: np.ndarray) -> np.ndarray:
    key = np.ascontiguousarray(key, np.float32)
    idx = np.zeros(len(key), np.int32)
    lib().lo_std_sort_desc(_vp(key), _vp(idx), len(key))
    return idx


class OracleLatent:
    """A latent print prepared for the oracle: non-empty minutiae templates, texture template, LUT."""

    def __init__(self, T, codebook: np.ndarray):
        self.minu = [Keep(m.x, m.y, m.ori, des=m.des) for m in T.minu if m.n > 0]
        self.tex = [Keep(t.x, t.y, t.ori, des=t.des) for t in T.tex if t.n > 0]
        self.lut = build_lut(self.tex[0].des, codebook) if self.tex else None
        self.minu_arr = (LoTemplate * max(1, len(self.minu)))(*[k.c for k in self.minu])
        self.tex_arr = (LoTemplate * max(1, len(self.tex)))(*[k.c for k in self.tex])


class OracleRolled:
    def __init__(self, T):
        self.minu = [Keep(m.x, m.y, m.ori, des=m.des) for m in T.minu if m.n > 0]
        self.tex = [Keep(t.x, t.y, t.ori, codes=t.des) for t in T.tex if t.n > 0]
        self.minu_arr = (LoTemplate * max(1, len(self.minu)))(*[k.c for k in self.minu])
        self.tex_arr = (LoTemplate * max(1, len(self.tex)))(*[k.c for k in self.tex])


_table = None


def score_pair(L: OracleLatent, R: OracleRolled):
    """-> (rc, comp[4] = score[0], score[1], score[2], score[28], fused score)"""
    global _table
    if _table is None:
        _table = make_table()
    comp = np.zeros(4, np.float32); fin = np.zeros(1, np.float32)
    rc = lib().lo_score_pair(L.minu_arr, len(L.minu), L.tex_arr, len(L.tex), _vp(L.lut) if L.lut is not None else None,
                             R.minu_arr, len(R.minu), R.tex_arr, len(R.tex), _vp(_table), SUBS, CLUSTERS,
                             _vp(comp), _vp(fin))
    return rc, comp, float(fin[0])
