"""nnk_mlpg_grad at the north_star shape (256 utterances, T=1000, static_dim=60) a few times, for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nnmnkwii_b200 import _device as dev, _lib  # noqa: E402

device = torch.device("cuda", 0)
g = torch.Generator(device=device).manual_seed(0)
B, T, sd = 256, 1000, 60
wc = _lib.make_windows(bench.WINDOWS)
chains = dev.chains_on_device(dev.simple_chains(sd), device)
off = torch.arange(B + 1, dtype=torch.int64, device=device) * T
v = torch.rand(B * T, 3 * sd, device=device, generator=g) + 0.1
go = torch.randn(B * T, sd, device=device, generator=g)
out = torch.zeros(B * T, 3 * sd, device=device)
for _ in range(3):
    dev.run_mlpg("grad", means=None, variances=v, rhs=go, out=out, offsets=off, lengths=None, order=None, chains=chains,
                 n_chain=sd, max_T=T, windows_c=wc, in_ld=3 * sd, var_ld=3 * sd, go_ld=sd, out_ld=3 * sd,
                 dtype_code=_lib.NNK_F32, go_f64=0, n_utt=B, device=device, check=False)
torch.cuda.synchronize()
print("done")
