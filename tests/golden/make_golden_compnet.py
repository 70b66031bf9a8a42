"""Generates tests/golden/golden_compnet.npz from the REFERENCE's own CompNet class
(/root/reference/extraction/models/net_compress.py, imported unmodified; torch CPU fp32, eval mode) and the
normalisation loop of /root/reference/extraction/descriptor_DR.py:150-152.  The reference ships no trained
weights, so the state is random (seeded), with non-trivial BatchNorm running statistics.

Run in the build container (the reference checkout does not travel to the GPU box):
    python tests/golden/make_golden_compnet.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/extraction/models/net_compress.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_net_compress", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(20261017)
    net = ref.CompNet(in_dims=192, out_dims=96)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(0.5 + torch.rand(96, generator=g))
                m.bias.copy_(0.2 * torch.randn(96, generator=g))
                m.running_mean.copy_(0.3 * torch.randn(96, generator=g))
                m.running_var.copy_(0.4 + torch.rand(96, generator=g))
            if isinstance(m, torch.nn.Linear):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) / m.weight.shape[1] ** 0.5)
                m.bias.copy_(0.1 * torch.randn(96, generator=g))
    net.eval()
    # 419 descriptors: not a multiple of any tile size; unit-norm-ish rows like the 192-d descriptors the
    # extraction networks emit, plus a few extreme rows
    x = torch.randn(419, 192, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * 1.73
    x[5] *= 40.0
    x[6] *= 1e-3
    x[7] = 0.0
    with torch.no_grad():
        y = net(x.clone()).numpy()
    feat = y.copy()
    for k in range(feat.shape[0]):  # descriptor_DR.py:150-152
        norm = np.linalg.norm(feat[k])
        feat[k] = feat[k] / norm * 1.73
    out = {"x": x.numpy(), "y_raw": y, "y_norm": feat}
    state = net.state_dict()
    out["state_names"] = np.array(list(state.keys()))
    for i, (k, v) in enumerate(state.items()):
        out[f"state_{i:02d}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_compnet.npz"), **out)
    print("wrote golden_compnet.npz:", {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith("state_")},
          len(state), "state tensors")


if __name__ == "__main__":
    main()
