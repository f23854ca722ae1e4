"""One masked-melcd launch at the bench size, for `ncu -k regex:frame_metric` (tools/ncu_summary.py reads the report)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from nnmnkwii_b200 import _device as dev
from nnmnkwii_b200 import _lib

device = torch.device("cuda", 0)
g = torch.Generator(device=device).manual_seed(0)
B, T, D = 2048, 1600, 25
X = torch.randn(B, T, D, device=device, generator=g)
Y = torch.randn(B, T, D, device=device, generator=g)
lens = torch.randint(700, T + 1, (B,), generator=torch.Generator().manual_seed(5)).to(device=device, dtype=torch.int32)
need = int(_lib.lib.nnk_metric_workspace_bytes(B, T))
ws = torch.zeros(need, dtype=torch.uint8, device=device)
res = torch.zeros(2, dtype=torch.float64, device=device)
for _ in range(3):
    _lib.check(_lib.lib.nnk_frame_metric(X.data_ptr(), Y.data_ptr(), _lib.NNK_F32, B, T, D, T * D, D, lens.data_ptr(), 0,
                                         ctypes.c_void_p(res.data_ptr()), ctypes.c_void_p(res.data_ptr() + 8),
                                         ctypes.c_void_p(ws.data_ptr()), ctypes.c_int64(need),
                                         dev.current_stream_ptr(device)), "nnk_frame_metric")
torch.cuda.synchronize()
print(res[0].item())
