"""Descriptor compression on the GPU: compnet_kernel through the C ABI against
(a) the golden vectors of the reference's own CompNet class (tests/golden/make_golden_compnet.py),
(b) the numpy oracle on seeded inputs of awkward sizes, (c) a plain torch fp32 restatement run on the same device,
and the 192-d enrollment path (CompNet -> normalise -> PQ encode -> Template2Bin_Byte_PQ_rolled).
Floating-point kernel: tolerance rtol 1e-5 / atol 1e-5 on outputs of norm 1.73 (stated per assert)."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden_compnet():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_compnet.npz"))


@pytest.fixture(scope="module")
def matcher(pkg, built, golden, golden_compnet):
    m = pkg.Matcher(codebook=golden["codebook"], device=0)
    m.load_compnet([golden_compnet[f"state_{i:02d}"] for i in range(len(golden_compnet["state_names"]))])
    yield m
    m.close()


def _layers(pkg, g):
    return pkg.matcher.compnet_layers([g[f"state_{i:02d}"] for i in range(len(g["state_names"]))])


def test_golden_vectors_of_the_reference_network(matcher, golden_compnet):
    x = golden_compnet["x"]
    raw = matcher.compress_descriptors(x, normalise=False)
    np.testing.assert_allclose(raw, golden_compnet["y_raw"], rtol=1e-5, atol=2e-5 * np.abs(golden_compnet["y_raw"]).max())
    # without the extreme rows the absolute tolerance is the one of ordinary descriptors
    keep = [i for i in range(len(x)) if i not in (5, 6, 7)]
    np.testing.assert_allclose(raw[keep], golden_compnet["y_raw"][keep], rtol=1e-5, atol=1e-5)
    got = matcher.compress_descriptors(x)
    np.testing.assert_allclose(got, golden_compnet["y_norm"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.73, atol=1e-5)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 47, 48, 49, 1000, 148 * 6 * 8 + 3])
def test_oracle_on_ragged_sizes(pkg, matcher, golden_compnet, n):
    import compnet_oracle as co
    rng = np.random.default_rng(1000 + n)
    x = rng.standard_normal((n, 192)).astype(np.float32)
    x *= 1.73 / np.linalg.norm(x, axis=1, keepdims=True)
    want = co.compress(_layers(pkg, golden_compnet), x)
    got = matcher.compress_descriptors(x)
    assert got.shape == (n, 96)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


def test_empty_input_and_missing_network(pkg, golden, matcher):
    assert matcher.compress_descriptors(np.zeros((0, 192), np.float32)).shape == (0, 96)
    m2 = pkg.Matcher(codebook=golden["codebook"], device=0)
    try:
        with pytest.raises(pkg.LafisError) as e:
            m2.compress_descriptors(np.zeros((4, 192), np.float32))
        assert e.value.status == pkg.matcher.LAFIS_ERR_ARG
    finally:
        m2.close()


def test_device_pointers_against_torch_fp32(pkg, matcher, golden_compnet):
    import torch
    layers = _layers(pkg, golden_compnet)
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    n = 200_003
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(n, 192, generator=g).to(dev)
    x = (x / x.norm(dim=1, keepdim=True) * 1.73).contiguous()
    out = torch.empty(n, 96, device=dev)
    torch.cuda.synchronize()
    matcher.compress_descriptors(x.data_ptr(), n, out.data_ptr())

    def block(l, v):
        L = {k: torch.from_numpy(a).to(dev) for k, a in layers[l].items()}
        v = v @ L["weight"].T + L["bias"]
        return (v - L["bn_mean"]) / torch.sqrt(L["bn_var"] + 1e-5) * L["bn_weight"] + L["bn_bias"]

    lrelu = lambda v: torch.where(v > 0, v, v * 0.2)
    h1 = lrelu(block(0, x))
    t = lrelu(block(1, h1))
    u = lrelu(block(2, t) + h1)
    y = block(3, u)
    want = y / y.norm(dim=1, keepdim=True) * 1.73
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)


def test_enroll_rolled_from_raw_descriptors(pkg, matcher, golden, golden_compnet, tmp_path):
    """192-d descriptors in -> the same file as compressing on the host side first (oracle) would give, up to the
    descriptor tolerance; PQ codes are compared where the nearest centroid is unambiguous."""
    import compnet_oracle as co
    T = pkg.templates
    cb = golden["codebook"]
    rng = np.random.default_rng(77)
    nm, nt = 57, 333
    minu_raw = rng.standard_normal((nm, 192)).astype(np.float32)
    tex_raw = rng.standard_normal((nt, 192)).astype(np.float32)
    minu_xyo = np.stack([rng.integers(40, 700, nm), rng.integers(40, 700, nm), rng.uniform(-3, 3, nm)], 1).astype(np.float32)
    tex_xyo = np.stack([24 + 16 * rng.integers(0, 40, nt), 24 + 16 * rng.integers(0, 40, nt), rng.uniform(-1.5, 1.5, nt)], 1).astype(np.float32)
    p = os.path.join(str(tmp_path), "raw.dat")
    matcher.enroll_rolled(p, minu_xyo, minu_raw, tex_xyo, tex_raw)
    got = T.read_template(p, latent=False)
    layers = _layers(pkg, golden_compnet)
    want_minu = co.compress(layers, minu_raw)
    want_tex = co.compress(layers, tex_raw)
    assert got.minu[0].des.shape == (nm, 96)
    np.testing.assert_allclose(got.minu[0].des, want_minu, rtol=1e-5, atol=1e-5)
    # the file's codes are the GPU encoder's codes of the GPU-compressed descriptors
    gpu_tex = matcher.compress_descriptors(tex_raw)
    assert np.array_equal(got.tex[0].des, matcher.pq_encode(gpu_tex))
    same = got.tex[0].des == T.pq_encode(want_tex, cb)
    assert same.mean() > 0.999  # a 1e-6 perturbation flips a nearest centroid only at a near-tie


def test_enroll_latent_from_raw_descriptors(pkg, matcher, golden_compnet, tmp_path):
    """28 minutiae templates + 1 texture template of 192-d descriptors -> one device call -> latent .dat whose
    descriptors are the oracle's compressed ones (tolerance as above), template after template."""
    import compnet_oracle as co
    T = pkg.templates
    rng = np.random.default_rng(5)
    sizes = [int(rng.integers(0, 60)) for _ in range(28)]
    sizes[3] = 0  # an empty minutiae template: written as its zero count, skipped by the readers
    minu_sets = [(np.stack([rng.integers(40, 700, n), rng.integers(40, 700, n), rng.uniform(-3, 3, n)], 1).astype(np.float32),
                  rng.standard_normal((n, 192)).astype(np.float32)) for n in sizes]
    nt = 180
    tex_sets = [(np.stack([24 + 16 * rng.integers(0, 40, nt), 24 + 16 * rng.integers(0, 40, nt), rng.uniform(-1.5, 1.5, nt)], 1).astype(np.float32),
                 rng.standard_normal((nt, 192)).astype(np.float32))]
    p = os.path.join(str(tmp_path), "lat192.dat")
    matcher.enroll_latent(p, minu_sets, tex_sets)
    got = T.read_template(p, latent=True)
    layers = _layers(pkg, golden_compnet)
    nonempty = [s for s in minu_sets if len(s[0])]
    assert len(got.minu) == len(nonempty) and len(got.tex) == 1
    for g, (xyo, des) in zip(got.minu, nonempty):
        assert np.array_equal(g.x, xyo[:, 0].astype(np.int16)) and g.des.shape == (len(xyo), 96)
        np.testing.assert_allclose(g.des, co.compress(layers, des), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got.tex[0].des, co.compress(layers, tex_sets[0][1]), rtol=1e-5, atol=1e-5)
    assert np.array_equal(got.tex[0].x, ((tex_sets[0][0][:, 0] - 24) / 16).astype(np.int16))
