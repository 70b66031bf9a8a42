"""Generates tests/golden/golden_pq.npz with the REFERENCE's own encoder: the class TrainedPQEncoder of
/root/reference/extraction/descriptor_PQ.py (lines 13-27) is executed from the reference source where it lies - the
module itself cannot be imported under Python 3 (it pulls in template_2 / cStringIO) - with `vq` bound to
scipy.cluster.vq.vq exactly as line 3 of that file does.  Inputs are float32 descriptors and the float32 codebook,
the dtypes of the reference pipeline (descriptor_PQ.py:323 codebook, Bin2Template_Byte_TF_C float32 descriptors).

Run in the build container (the reference checkout does not travel to the GPU box):
    python tests/golden/make_golden_pq.py
"""
import os

import numpy as np
from scipy.cluster.vq import vq

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/extraction/descriptor_PQ.py"


def reference_encoder_class():
    src = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("class TrainedPQEncoder"))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def ") or src[i].startswith("class "))
    ns = {"np": np, "vq": vq}
    exec("\n".join(src[start:end]), ns)
    return ns["TrainedPQEncoder"]


def main():
    g = np.load(os.path.join(HERE, "golden_small.npz"))
    cb = g["codebook"].astype(np.float32)  # the shipped 16 x 256 x 6 codebook
    rng = np.random.default_rng(20261017)
    n = 4096
    des = (1.73 * rng.standard_normal((n, 96)) / np.sqrt(96)).astype(np.float32)
    des = (des / np.linalg.norm(des, axis=1, keepdims=True) * 1.73).astype(np.float32)
    # near-ties on purpose: 512 points whose sub-vectors sit (almost) on the bisector of two centroids
    for i in range(512):
        m = int(rng.integers(0, 16))
        a, b = rng.choice(256, 2, replace=False)
        mid = 0.5 * (cb[m, a] + cb[m, b])
        des[i, m * 6:(m + 1) * 6] = mid + (rng.standard_normal(6) * 1e-7).astype(np.float32)
    # exact copies of centroids (distance 0 to one code)
    for i in range(512, 600):
        for m in range(16):
            des[i, m * 6:(m + 1) * 6] = cb[m, int(rng.integers(0, 256))]
    Enc = reference_encoder_class()
    codes = Enc(cb, np.uint8).encode_multi(des)
    np.savez_compressed(os.path.join(HERE, "golden_pq.npz"), des=des, codes=codes.astype(np.uint8))
    print("golden_pq.npz:", des.shape, codes.shape, codes.dtype)


if __name__ == "__main__":
    main()
