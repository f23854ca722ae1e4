"""Phase timers of mlpg_fwd_as_kernel on the cfg2 batch.  Needs a debug build:
   NNK_NVCC_EXTRA=-DNNK_AS_PROF python -m nnmnkwii_b200.build   (rebuild without it afterwards)"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nnmnkwii_b200 import _device as dev, _lib, paramgen as G  # noqa: E402

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


def step():
    dev.run_mlpg("fwd", means=d_m, variances=d_v, rhs=None, out=d_out, offsets=off, lengths=None, order=order,
                 chains=chains, n_chain=layout.n_chain, max_T=int(lens.max()), windows_c=wc, in_ld=187, var_ld=187,
                 go_ld=0, out_ld=63, dtype_code=_lib.NNK_F32, go_f64=0, n_utt=len(lens), device=device, check=False)


NA = int(os.environ.get("NA", "3"))
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    step()
_lib.lib.nnk_as_prof_read(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
_lib.lib.nnk_as_prof_read(buf)
ms = e0.elapsed_time(e1)
n_cta = len(lens) * 2
names = ["A wait pb_empty", "A wait input TMA", "A convert+assemble+publish", "-", "S wait pb_full", "S eliminate", "S wait scratch TMA", "S backward"]
print("step %.3f ms; per-CTA average cycles (assemblers: per warp, 2 warps):" % ms)
for i, nm in enumerate(names):
    if nm == "-":
        continue
    div = n_cta * (NA if i < 4 else 1)
    print("  %-28s %10.0f" % (nm, buf[i] / div))
