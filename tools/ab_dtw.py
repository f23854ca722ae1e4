#!/usr/bin/env python
"""A/B timing of DTW kernel variants (NNK_LIB_PATH): configs[3] exact and FastDTW ms per 512 pairs + parity of
32 pairs against the C oracle.    python tools/ab_dtw.py lib1.so lib2.so ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INNER = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
import bench, oracle
from nnmnkwii_b200.preprocessing import alignment as A
X, Y = bench.make_dtw_pairs(512)
dev = torch.device("cuda", 0)
Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
out = {}
for name, radius in (("exact", -1), ("fast", 1)):
    res = A._align_batch(Xd, Yd, 1, radius)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        res = A._align_batch(Xd, Yd, 1, radius)
    e1.record(); torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / 5
    L = res.path_len.cpu().numpy(); lx = res.len_x.cpu().numpy(); ly = res.len_y.cpu().numpy()
    bad = 0
    for n in range(0, 512, 16):
        d0, oi, oj, c0 = oracle.fastdtw(X[n, :lx[n]], Y[n, :ly[n]], radius=radius, kind="melcd")
        bad += not (np.array_equal(res.path_i[n, :L[n]].cpu().numpy(), oi) and float(res.dist[n]) == d0)
    out[name + "_bad"] = bad
print("ABRES " + json.dumps(out))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, NNK_LIB_PATH=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", INNER], env=env, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("ABRES ")]
    print(os.path.basename(lib), line[0][6:] if line else ("FAILED: " + r.stderr[-600:]))
    sys.stdout.flush()
