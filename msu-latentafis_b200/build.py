"""Builds the native parts of the package in-tree with nvcc / g++ (no JIT cache):

  lib/liblatentafis_b200.so   CUDA kernels (sm_100a) + the C ABI of include/latentafis_b200.h
  lib/libhostcheck.so         host build of the bit-exact helper headers (exact_math.h,
                              stdsort_emul.h) and of the .dat parsers, for the CPU unit tests
  bin/match                   drop-in command-line matcher (matching/main.cpp interface)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
BIN = os.path.join(HERE, "bin")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-extended-lambda",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "550"]


# dev experiments: extra compiler flags (e.g. LAFIS_EXTRA_NVCC="-DLAFIS_TEX_ADD_MODE=1")
NVCC_FLAGS += os.environ.get("LAFIS_EXTRA_NVCC", "").split()


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.isfile(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    out = [os.path.join(ROOT, "include", "latentafis_b200.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def lib_path() -> str:
    return os.path.join(LIB, "liblatentafis_b200.so")


def build_all(force: bool = False, verbose: bool = False) -> None:
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(BIN, exist_ok=True)
    srcs = _sources()
    run = lambda cmd: subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    so = lib_path()
    if force or _newer(so, srcs):
        # every translation unit as CUDA (-x cu: they share lafis_internal.h, which carries device declarations), in
        # parallel; sharded.cu needs <nccl.h> for types only - NCCL itself is bound at run time (dlopen), so no -lnccl
        from concurrent.futures import ThreadPoolExecutor
        obj_dir = os.path.join(HERE, "build")
        os.makedirs(obj_dir, exist_ok=True)
        units = ["lafis_api.cu", "sharded.cu", "dat_format.cpp", "drivers.cpp"]
        objs = [os.path.join(obj_dir, u.rsplit(".", 1)[0] + ".o") for u in units]
        with ThreadPoolExecutor(len(units)) as ex:
            list(ex.map(lambda uo: run([_nvcc()] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, uo[0]), "-o", uo[1]]),
                        zip(units, objs)))
        run([_nvcc()] + NVCC_FLAGS + ["-shared"] + objs + ["-ldl", "-o", so])
    hc = os.path.join(LIB, "libhostcheck.so")
    if force or _newer(hc, srcs):
        run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
             os.path.join(CSRC, "hostcheck.cpp"), os.path.join(CSRC, "dat_format.cpp"), "-o", hc])
    exe = os.path.join(BIN, "match")
    if force or _newer(exe, srcs + [so]):
        run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(CSRC, "match_main.cpp"), "-I", os.path.join(ROOT, "include"),
             "-L", LIB, "-llatentafis_b200", "-Wl,-rpath,$ORIGIN/../lib", "-o", exe])


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built", lib_path())
