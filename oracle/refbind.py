"""ctypes bindings for oracle/_ref/libref_matcher.so (the reference matcher compiled from its own
sources by oracle/build_ref.sh) — TEST INFRASTRUCTURE ONLY.  May be imported from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, nowhere else."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_matcher.so")
CLI_PATH = os.path.join(HERE, "_ref", "match")

os.environ.setdefault("OMP_STACKSIZE", "32M")  # One2One_texture_matching keeps a 4 MB array on the stack

_lib = None


def available() -> bool:
    return os.path.isfile(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_matcher_new.restype = C.c_void_p
        L.ref_matcher_new.argtypes = [C.c_char_p]
        L.ref_matcher_free.argtypes = [C.c_void_p]
        for f in (L.ref_load_latent, L.ref_load_rolled):
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.ref_free_latent.argtypes = [C.c_void_p]
        L.ref_free_rolled.argtypes = [C.c_void_p]
        L.ref_latent_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_score_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_score_gallery.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_one2list.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.ref_list2list.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.ref_minutiae_score.restype = C.c_float
        L.ref_minutiae_score.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_texture_score.restype = C.c_float
        L.ref_texture_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_prune.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3
        L.ref_std_sort_desc.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_atan2f.restype = C.c_float
        L.ref_atan2f.argtypes = [C.c_float, C.c_float]
        _lib = L
    return _lib


class RefMatcher:
    """Thin owner of a reference `PQ::Matcher` plus loaded templates."""

    def __init__(self, codebook_path: str):
        self.L = lib()
        self.m = self.L.ref_matcher_new(codebook_path.encode())
        self._latents = []
        self._rolled = []

    def load_latent(self, path: str):
        rc = C.c_int(0)
        h = self.L.ref_load_latent(self.m, path.encode(), C.byref(rc))
        self._latents.append(h)
        return h, rc.value

    def load_rolled(self, path: str):
        rc = C.c_int(0)
        h = self.L.ref_load_rolled(self.m, path.encode(), C.byref(rc))
        self._rolled.append(h)
        return h, rc.value

    def score_pair(self, latent, rolled):
        comp = np.zeros(4, np.float32)
        fin = np.zeros(1, np.float32)
        rc = self.L.ref_score_pair(self.m, latent, rolled, comp.ctypes.data, fin.ctypes.data)
        return rc, comp, float(fin[0])

    def score_gallery(self, latent, rolled_handles, nthreads: int = 0):
        n = len(rolled_handles)
        arr = (C.c_void_p * n)(*rolled_handles)
        fin = np.full(n, -1, np.float32)
        comps = np.zeros((n, 4), np.float32)
        rc = self.L.ref_score_gallery(self.m, latent, arr, n, fin.ctypes.data, comps.ctypes.data, nthreads)
        return rc, fin, comps

    def correspondences(self, latent, rolled):
        """corr3 of the three selected minutiae templates as the reference writes them (matcher.cpp:497-505):
        -> (rc, [3 arrays of shape (n_i, 4): latent x, latent y, rolled x, rolled y])"""
        import tempfile
        tmp = tempfile.mkdtemp(prefix="lafis_corr_")
        try:
            prefix = os.path.join(tmp, "corr")
            self.L.ref_save_corr.restype = C.c_int
            self.L.ref_save_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
            rc = self.L.ref_save_corr(self.m, latent, rolled, prefix.encode())
            out = []
            for i in range(3):
                p = f"{prefix}_{i}.csv"
                rows = []
                if os.path.isfile(p):
                    rows = [[int(x) for x in line.split(",")] for line in open(p).read().split() if line]
                out.append(np.array(rows, np.int16).reshape(-1, 4))
            return rc, out
        finally:
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)

    def prune(self, which, latent, latent_tpl, rolled, v, li, rj):
        v = np.ascontiguousarray(v, np.float32)
        li = np.ascontiguousarray(li, np.int32)
        rj = np.ascontiguousarray(rj, np.int32)
        n = len(v)
        ov = np.zeros(max(n, 1), np.float32)
        oli = np.zeros(max(n, 1), np.int32)
        orj = np.zeros(max(n, 1), np.int32)
        k = self.L.ref_prune(self.m, which, latent, latent_tpl, rolled, v.ctypes.data, li.ctypes.data, rj.ctypes.data,
                             n, ov.ctypes.data, oli.ctypes.data, orj.ctypes.data)
        return ov[:k], oli[:k], orj[:k]

    def close(self):
        for h in self._latents:
            self.L.ref_free_latent(h)
        for h in self._rolled:
            self.L.ref_free_rolled(h)
        self._latents, self._rolled = [], []
        if self.m:
            self.L.ref_matcher_free(self.m)
            self.m = None


def std_sort_desc(key: np.ndarray) -> np.ndarray:
    key = np.ascontiguousarray(key, np.float32)
    idx = np.zeros(len(key), np.int32)
    lib().ref_std_sort_desc(key.ctypes.data, idx.ctypes.data, len(key))
    return idx
