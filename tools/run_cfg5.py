"""BASELINE.json configs[4] on ONE GPU: 8192 utterances, mixed T in [200, 2000], D=187 (9.0e6 frames).
Checks a sample of utterances against the oracle and reports device-resident frames/s."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, oracle
from nnmnkwii_b200 import paramgen as G

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(5)
lens = rng.integers(200, 2001, size=n_utt)
n = int(lens.sum())
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
m = torch.rand((n, 187), device=dev, generator=g)
v = torch.rand((n, 187), device=dev, generator=g) + 0.1
lay = G.merlin_layout()
y = G.mlpg_batch(m, v, bench.WINDOWS, lengths=lens, layout=lay)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 3
for _ in range(reps):
    y = G.mlpg_batch(m, v, bench.WINDOWS, lengths=lens, layout=lay, check=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
off = np.concatenate([[0], np.cumsum(lens)])
worst = 0.0
for u in (0, n_utt // 2, n_utt - 1, int(np.argmax(lens)), int(np.argmin(lens))):
    a, b = int(off[u]), int(off[u + 1])
    ref = oracle.mlpg(m[a:b, :180].cpu().numpy(), v[a:b, :180].cpu().numpy(), bench.WINDOWS)
    got = y[a:b, :60].cpu().numpy()
    worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
print({"config": "cfg5", "utterances": n_utt, "frames": n, "ms": ms, "frames_per_sec": n / (ms * 1e-3),
       "hbm_gbs_algorithmic": 1744 * n / (ms * 1e-3) / 1e9, "max_rel_err_vs_oracle": worst})
