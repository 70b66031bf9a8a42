"""TEST INFRASTRUCTURE - CPU restatement (numpy, fp32) of the reference's descriptor compression:
CompNet in eval mode (extraction/models/net_compress.py:33-53, BasicBlock :7-31) followed by the
re-normalisation of extraction/descriptor_DR.py:150-152.  Only tests/ may import this module; the product path is
the CUDA kernel compnet_kernel (msu-latentafis_b200/csrc/compnet.cuh).

Pinned: tests/test_compnet_oracle.py checks it against tests/golden/golden_compnet.npz, produced by
tests/golden/make_golden_compnet.py from the reference's own CompNet class (torch, CPU, fp32) imported from
/root/reference."""
from __future__ import annotations

import numpy as np

F = np.float32


def _linear(x, w, b):  # nn.Linear: x @ W^T + b                                   net_compress.py:11, :14, :42, :47
    return (x @ w.T.astype(F) + b.astype(F)).astype(F)


def _bn_eval(x, g, b, mean, var, eps=1e-5):  # nn.BatchNorm1d, running statistics   net_compress.py:12, :15, :43, :48
    inv = (F(1.0) / np.sqrt(var.astype(F) + F(eps))).astype(F)
    return ((x - mean.astype(F)) * inv * g.astype(F) + b.astype(F)).astype(F)


def _lrelu(x):  # LeakyReLU(0.2) / F.leaky_relu(y, 0.2)                             net_compress.py:13, :30, :44
    return np.where(x > 0, x, x * F(0.2)).astype(F)


def compnet_forward(layers, x):
    """layers: 4 dicts {weight, bias, bn_weight, bn_bias, bn_mean, bn_var} (layer1, layer2.layers[0:2],
    layer2.layers[3:5], layer3); x [n,192] -> [n,96] (CompNet.forward, net_compress.py:52-56)."""
    x = np.ascontiguousarray(x, F)

    def block(l, v):
        L = layers[l]
        return _bn_eval(_linear(v, L["weight"], L["bias"]), L["bn_weight"], L["bn_bias"], L["bn_mean"], L["bn_var"])

    h1 = _lrelu(block(0, x))                 # layer1
    t = _lrelu(block(1, h1))                 # layer2.layers[0..2]
    u = _lrelu(block(2, t) + h1)             # layer2.layers[3..4] + residual, then leaky_relu (:24-30)
    return block(3, u)                       # layer3


def normalise_173(y):
    """descriptor_DR.py:150-152: features[k] / np.linalg.norm(features[k]) * 1.73, in float32."""
    y = np.ascontiguousarray(y, F).copy()
    for k in range(y.shape[0]):
        norm = np.linalg.norm(y[k])
        y[k] = y[k] / norm * 1.73
    return y


def compress(layers, x):
    return normalise_173(compnet_forward(layers, x))
