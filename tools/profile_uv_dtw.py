"""One pass of the UnitVarianceMLPG sweeps (configs[2]) and of exact / Fast DTW (configs[3], 64 pairs) for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nnmnkwii_b200 import _uvmlpg as uv  # noqa: E402
from nnmnkwii_b200 import paramgen as G  # noqa: E402
from nnmnkwii_b200.preprocessing import alignment as A  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "uv"
dev = torch.device("cuda", 0)
if what == "uv":
    T, sd, B = 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 60, 64
    R = torch.from_numpy(G.unit_variance_mlpg_matrix(bench.WINDOWS, T)).to(dev)
    band = uv.band_of(R, dev)
    x = torch.randn(B, T, 3 * sd, device=dev)
    go = torch.randn(B, T, sd, device=dev)
    for _ in range(3):
        uv.apply_forward(band, x, False)
        uv.apply_backward(band, go, False, 3 * sd)
else:
    n = 64
    X, Y = bench.make_dtw_pairs(n)
    Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
    for _ in range(2):
        A._align_batch(Xd, Yd, 1, -1 if what == "exact" else 1)
torch.cuda.synchronize()
print("done", what)
