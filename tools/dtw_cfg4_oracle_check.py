#!/usr/bin/env python
"""configs[3] at FULL size (512 pairs, T~U{700..900}, 25-dim, melcd): every path, cell count and total
distance of the CUDA aligner compared with the C oracle, FastDTW(radius=1) and exact DP.

    python tools/dtw_cfg4_oracle_check.py > gpurun_out/dtw_cfg4_oracle_check.log

The oracle is the repo's restatement (the fastdtw package is absent: parity unpinned vs the package)."""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _one(args):
    import oracle
    x, y, radius = args
    d, pi, pj, cells = oracle.fastdtw(x, y, radius=radius, kind="melcd")
    return d, np.asarray(pi), np.asarray(pj), cells


def main():
    import torch
    import bench
    from nnmnkwii_b200.preprocessing import alignment as A
    X, Y = bench.make_dtw_pairs(512)
    dev = torch.device("cuda", 0)
    Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
    n_bad = 0
    with mp.get_context("spawn").Pool(min(32, bench.usable_cores())) as pool:
        for radius in (1, -1):
            res = A._align_batch(Xd, Yd, 1, radius)
            torch.cuda.synchronize()
            L = res.path_len.cpu().numpy()
            pi, pj = res.path_i.cpu().numpy(), res.path_j.cpu().numpy()
            dist, cells = res.dist.cpu().numpy(), res.cells.cpu().numpy()
            lx, ly = res.len_x.cpu().numpy(), res.len_y.cpu().numpy()
            t0 = time.perf_counter()
            ref = pool.map(_one, [(X[n, :lx[n]], Y[n, :ly[n]], radius) for n in range(512)], chunksize=4)
            cpu_s = time.perf_counter() - t0
            bad = 0
            for n, (d0, oi, oj, c0) in enumerate(ref):
                ok = (L[n] == len(oi) and np.array_equal(pi[n, :L[n]], oi) and np.array_equal(pj[n, :L[n]], oj)
                      and dist[n] == d0 and cells[n] == c0)
                bad += not ok
            n_bad += bad
            print("radius=%2d: 512 pairs, %d cells, paths/cells/distance identical to the C oracle on %d of 512 pairs "
                  "(oracle wall %.2f s)" % (radius, int(cells.sum()), 512 - bad, cpu_s))
    print("RESULT:", "ALL IDENTICAL" if n_bad == 0 else "%d MISMATCHES" % n_bad)
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
