#!/usr/bin/env python
"""Small MLPG workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
forward and gradient through the warp-specialised TMA kernel (ragged batch, odd T, Merlin layout) and the
single-utterance paths.  Kept small: racecheck slows kernels by two orders of magnitude.

    compute-sanitizer --tool racecheck python tools/sanitize_mlpg.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import oracle
    from nnmnkwii_b200 import paramgen as G
    ws = [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5])), (1, 1, np.array([1.0, -2.0, 1.0]))]
    rng = np.random.default_rng(0)
    lens = np.array([37, 5, 64, 1, 21])
    n = int(lens.sum())
    m = rng.random((n, 187), dtype=np.float32)
    v = rng.random((n, 187), dtype=np.float32) + 0.1
    y = G.mlpg_batch(torch.from_numpy(m).cuda(), torch.from_numpy(v).cuda(), ws, lengths=lens, layout=G.merlin_layout())
    ref = oracle.mlpg(m[:37, :180], v[:37, :180], ws)
    err = np.abs(y[:37, :60].cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-6, err
    go = rng.standard_normal((37, 60)).astype(np.float32)
    g = G.mlpg_grad(m[:37, :180], v[:37, :180], ws, go)
    gref = oracle.mlpg_grad(m[:37, :180], v[:37, :180], ws, go)
    assert np.abs(g - gref).max() / np.abs(gref).max() < 2e-6
    y1 = G.mlpg(m[:64, :177], np.ones(177, np.float32), ws)  # global variance path
    assert np.isfinite(y1).all()
    torch.cuda.synchronize()
    print("sanitize_mlpg ok: max rel err %.2e" % err)


if __name__ == "__main__":
    main()
