"""Dev helper: device-resident throughput of the two CompNet kernels (descriptors/s and fp32 FMA rate), CUDA events."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
g = np.load(os.path.join(ROOT, "tests", "golden", "golden_compnet.npz"))
m = pkg.Matcher(codebook=pkg.templates.synthetic_codebook(), device=0)
m.load_compnet([g[f"state_{i:02d}"] for i in range(len(g["state_names"]))])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
x = torch.randn(n, 192, device="cuda:0")
out = torch.empty(n, 96, device="cuda:0")
torch.cuda.synchronize()
for _ in range(3):
    m.compress_descriptors(x.data_ptr(), n, out.data_ptr())
s = torch.cuda.ExternalStream(m.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(s)
for _ in range(reps):
    m.compress_descriptors(x.data_ptr(), n, out.data_ptr())
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fma = n * 46080 / (ms * 1e-3)
print(f"compnet_l1_kernel + compnet_l234_kernel: {n} descriptors in {ms:.3f} ms = {n / ms / 1e3:.1f} M descriptors/s, {fma / 1e12:.2f} T FMA/s "
      f"({fma / (148 * 128 * 1.965e9) * 100:.1f} % of the fp32 FMA issue rate), {n * (192 + 96) * 4 / ms / 1e6:.0f} GB/s HBM")
m.close()
