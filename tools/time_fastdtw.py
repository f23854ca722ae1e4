import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from nnmnkwii_b200.preprocessing import alignment as A
X, Y = bench.make_dtw_pairs(512)
dev = torch.device("cuda", 0)
Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
for _ in range(2):
    A._align_batch(Xd, Yd, 1, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    A._align_batch(Xd, Yd, 1, 1)
e1.record(); torch.cuda.synchronize()
print("ms/batch=%.3f" % (e0.elapsed_time(e1) / 5))
