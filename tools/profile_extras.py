"""Run the DTW (configs[3]) and UnitVarianceMLPG (configs[2]) kernels once (for ncu launch lists)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
print(bench.bench_extras(dev, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 1))
