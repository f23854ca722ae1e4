#!/usr/bin/env python
"""Small workloads of the round-2 kernels for compute-sanitizer: fused (warp-pipelined) exact DTW, FastDTW,
factored UnitVarianceMLPG sweeps (even and odd static_dim), GMM mapping kernels, segment copy.

    compute-sanitizer --tool memcheck python tools/sanitize_others.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import oracle
    from nnmnkwii_b200 import autograd as AF
    from nnmnkwii_b200 import paramgen as G
    from nnmnkwii_b200.baseline.gmm import MLPG, MLPGBase
    from nnmnkwii_b200.preprocessing import alignment as A
    rng = np.random.default_rng(0)
    dev = torch.device("cuda", 0)
    # DTW: 3 pairs, 70..100 frames (3 row groups -> the inter-warp pipeline is exercised), 25 dims
    N, D, Tm = 3, 25, 100
    X = np.zeros((N, Tm, D), np.float32)
    Y = np.zeros((N, Tm, D), np.float32)
    lens = [(70, 100), (100, 83), (33, 96)]
    for n, (tx, ty) in enumerate(lens):
        X[n, :tx] = np.cumsum(rng.standard_normal((tx, D)), 0) * 0.1
        Y[n, :ty] = np.cumsum(rng.standard_normal((ty, D)), 0) * 0.1
    Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
    for radius in (-1, 1):
        res = A._align_batch(Xd, Yd, 1, radius)
        torch.cuda.synchronize()
        L = res.path_len.cpu().numpy()
        for n, (tx, ty) in enumerate(lens):
            d0, oi, oj, c0 = oracle.fastdtw(X[n, :tx], Y[n, :ty], radius=radius, kind="melcd")
            assert np.array_equal(res.path_i[n, :L[n]].cpu().numpy(), oi) and float(res.dist[n]) == d0, (radius, n)
    # UnitVarianceMLPG: factored path, even and odd static_dim
    ws = [(0, 0, np.array([1.0])), (1, 1, np.array([-0.5, 0.0, 0.5])), (1, 1, np.array([1.0, -2.0, 1.0]))]
    T = 200
    R = torch.from_numpy(G.unit_variance_mlpg_matrix(ws, T)).to(dev)
    for sd in (6, 5):
        mu = torch.randn(2, T, 3 * sd, device=dev).requires_grad_(True)
        y = AF.unit_variance_mlpg(R, mu)
        y.sum().backward()
        ref = R.double() @ mu[0].detach().double().view(T, 3, sd).transpose(0, 1).reshape(3 * T, sd)
        assert float((y[0].double() - ref).abs().max()) < 1e-5
    # GMM kernels
    M, dim = 4, 12
    Am = rng.standard_normal((M, 2 * dim, 2 * dim)) / np.sqrt(2 * dim)
    gmm = types.SimpleNamespace(means_=rng.standard_normal((M, 2 * dim)), covariances_=Am @ Am.transpose(0, 2, 1) + 0.5 * np.eye(2 * dim),
                                weights_=np.full(M, 1.0 / M), covariance_type="full")
    src = rng.standard_normal((37, dim))
    assert np.isfinite(MLPGBase(gmm).transform(src)).all()
    assert np.isfinite(MLPG(gmm, windows=ws[:2]).transform(src)).all()
    torch.cuda.synchronize()
    print("sanitize_others ok")


if __name__ == "__main__":
    main()
