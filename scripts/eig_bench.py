"""Time cna_sym_eig_top and print the SM clocks of its four phases (run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cna_b200 import _lib  # noqa: E402

_lib.load()
for n, k in ((200, 16), (100, 8), (500, 40)):
    rng = np.random.default_rng(0)
    X = rng.normal(size=(20 * n, n)) * np.linspace(3.0, 1.0, n)
    G = torch.as_tensor(X.T @ X, device="cuda")
    w = torch.empty(k, dtype=torch.float64, device="cuda")
    ut = torch.empty((k, n), dtype=torch.float64, device="cuda")
    de = torch.empty(2 * n + 12, dtype=torch.float64, device="cuda")
    for _ in range(3):
        _lib.sym_eig_top(G, k, w, ut, de)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        _lib.sym_eig_top(G, k, w, ut, de)
    b.record()
    torch.cuda.synchronize()
    ph = de[2 * n:].cpu().numpy()
    print(f"n={n} k={k}: {a.elapsed_time(b) / 20:.3f} ms per call; clocks: tridiagonalisation {ph[0]:.0f}, "
          f"multisection {ph[1]:.0f}, inverse iteration {ph[2]:.0f}, back-transformation {ph[3]:.0f}; sub {ph[4:]}")
