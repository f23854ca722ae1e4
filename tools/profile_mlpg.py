"""Run the cfg2 MLPG device step a few times (for ncu).  usage: python tools/profile_mlpg.py [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nnmnkwii_b200 import _device as dev, _lib, paramgen as G  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lens, means, variances = bench.make_batch(0)
device = torch.device("cuda", 0)
layout = G.merlin_layout()
n_rows = int(lens.sum())
off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(device)
d_m, d_v = torch.from_numpy(means).to(device), torch.from_numpy(variances).to(device)
d_out = torch.zeros((n_rows, 63), dtype=torch.float32, device=device)
order = torch.from_numpy(np.argsort(-lens, kind="stable").astype(np.int32)).to(device)
chains = dev.chains_on_device(layout.chains, device)
wc = _lib.make_windows(bench.WINDOWS)
for _ in range(reps):
    dev.run_mlpg("fwd", means=d_m, variances=d_v, rhs=None, out=d_out, offsets=off, lengths=None, order=order,
                 chains=chains, n_chain=layout.n_chain, max_T=int(lens.max()), windows_c=wc, in_ld=187, var_ld=187,
                 go_ld=0, out_ld=63, dtype_code=_lib.NNK_F32, go_f64=0, n_utt=len(lens), device=device, check=False)
torch.cuda.synchronize()
print("done")
