"""Descriptor compression, CPU side: the numpy oracle against the golden vectors produced by the
reference's own CompNet class (tests/golden/make_golden_compnet.py), and the positional state_dict mapping of the
host mirror (descriptor_DR.py:39-46)."""
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def golden_compnet():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_compnet.npz"))


def golden_state(g):
    return [g[f"state_{i:02d}"] for i in range(len(g["state_names"]))]


def test_state_dict_is_mapped_by_position(pkg, golden_compnet):
    names = [str(s) for s in golden_compnet["state_names"]]
    assert names[0] == "layer1.0.weight" and names[6] == "layer1.1.num_batches_tracked" and len(names) == 28
    layers = pkg.matcher.compnet_layers(golden_state(golden_compnet))
    assert [L["weight"].shape for L in layers] == [(96, 192), (96, 96), (96, 96), (96, 96)]
    want = {n: golden_compnet[f"state_{i:02d}"] for i, n in enumerate(names)}
    assert np.array_equal(layers[2]["weight"], want["layer2.layers.3.weight"])
    assert np.array_equal(layers[2]["bn_var"], want["layer2.layers.4.running_var"])
    assert np.array_equal(layers[3]["bn_mean"], want["layer3.1.running_mean"])
    with pytest.raises(ValueError):
        pkg.matcher.compnet_layers(golden_state(golden_compnet)[:-3])


def test_oracle_matches_the_reference_network(pkg, golden_compnet):
    import compnet_oracle as co
    layers = pkg.matcher.compnet_layers(golden_state(golden_compnet))
    x = golden_compnet["x"]
    y = co.compnet_forward(layers, x)
    # fp32 GEMMs in a different summation order than torch's: tolerance, stated here
    np.testing.assert_allclose(y, golden_compnet["y_raw"], rtol=1e-5, atol=2e-5 * np.abs(golden_compnet["y_raw"]).max())
    yn = co.compress(layers, x)
    np.testing.assert_allclose(yn, golden_compnet["y_norm"], rtol=1e-5, atol=1e-5)
    assert np.allclose(np.linalg.norm(yn, axis=1), 1.73, atol=1e-5)
