"""Shared helpers for the test-suite: golden fixture access and oracle scoring of template sets."""
from __future__ import annotations

import os

import numpy as np

# gallery entries whose outcome in the reference is undefined behaviour (header-only file: the
# loader reads its template counts from a stream that already hit EOF, matcher.cpp:921-934)
UB_GALLERY = {"r18_empty"}


def write_golden_files(golden, root):
    gdir, ldir = os.path.join(root, "gallery"), os.path.join(root, "latent")
    os.makedirs(gdir, exist_ok=True)
    os.makedirs(ldir, exist_ok=True)
    for g in golden["gallery_names"]:
        with open(os.path.join(gdir, f"{g}.dat"), "wb") as f:
            f.write(golden["gal_" + str(g)].tobytes())
    for l in golden["all_latent_names"]:
        with open(os.path.join(ldir, f"{l}.dat"), "wb") as f:
            f.write(golden["lat_" + str(l)].tobytes())
    return gdir, ldir


def oracle_scores(ob, T, latents, rolled, codebook):
    """-> (rc [Q,G], comp [Q,G,4], final [Q,G]) from the plain-C oracle."""
    Q, G = len(latents), len(rolled)
    rc = np.zeros((Q, G), np.int32)
    comp = np.zeros((Q, G, 4), np.float32)
    fin = np.zeros((Q, G), np.float32)
    OR = [ob.OracleRolled(r) for r in rolled]
    for i, l in enumerate(latents):
        OL = ob.OracleLatent(l, codebook)
        for j, r in enumerate(OR):
            rc[i, j], comp[i, j], fin[i, j] = ob.score_pair(OL, r)
    return rc, comp, fin


def rank_list(scores, k, base=0):
    """(score desc, index asc) rank list of one score row."""
    order = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), i))[:k]
    return [base + i for i in order]
