"""GPU parity tests for DTW: the CUDA wavefront kernel (through the C ABI / DTWAligner) must return
back-track indices IDENTICAL to the oracle (bit-exact integer work) and the same float64 distance.
The oracle is a restatement of fastdtw's published algorithm (parity unpinned w.r.t. the package
itself -- see oracle/nnk_oracle.c); golden paths come from the literal pure-Python restatement."""
import numpy as np
import pytest

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu


def _series(T, D, seed, dtype=np.float32):
    r = np.random.default_rng(seed)
    return (np.cumsum(r.standard_normal((T, D)), 0) * 0.3).astype(dtype)


def _pad(seqs, T=None, dtype=np.float32):
    T = T or max(len(s) for s in seqs)
    out = np.zeros((len(seqs), T, seqs[0].shape[1]), dtype=dtype)
    for i, s in enumerate(seqs):
        out[i, :len(s)] = s
    return out


def _align(X, Y, kind, radius):
    import torch
    from nnmnkwii_b200.preprocessing import alignment as A
    res = A._align_batch(torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda(), kind, radius)
    torch.cuda.synchronize()
    return (res.path_i.cpu().numpy(), res.path_j.cpu().numpy(), res.path_len.cpu().numpy(), res.dist.cpu().numpy(),
            res.cells.cpu().numpy(), res.len_x.cpu().numpy(), res.len_y.cpu().numpy())


def test_paths_match_golden(dtw_golden):
    g = dtw_golden
    for case in range(5):
        x, y, radius = g["c%d_x" % case], g["c%d_y" % case], int(g["c%d_radius" % case])
        for r, name in ((radius, "fast"), (-1, "exact")):
            pi, pj, L, d, cells, lx, ly = _align(x[None], y[None], 1, r)
            path = g["c%d_%s_path" % (case, name)]
            assert L[0] == len(path)
            assert np.array_equal(pi[0, :L[0]], path[:, 0]) and np.array_equal(pj[0, :L[0]], path[:, 1])
            assert d[0] == float(g["c%d_%s_dist" % (case, name)])
            if name == "fast":
                assert cells[0] == int(g["c%d_fast_cells" % case])


def test_batch_vs_oracle_bit_exact_indices():
    rng = np.random.default_rng(4)
    for D, dtype in ((25, np.float32), (3, np.float64), (130, np.float32)):
        xs = [_series(int(rng.integers(1, 120)), D, 10 + i, dtype) for i in range(24)]
        ys = [_series(int(rng.integers(1, 120)), D, 500 + i, dtype) for i in range(24)]
        X, Y = _pad(xs, 125, dtype), _pad(ys, 121, dtype)
        for kind, kname in ((0, "euclid"), (1, "melcd")):
            for radius in (1, 2, -1):
                pi, pj, L, d, cells, lx, ly = _align(X, Y, kind, radius)
                for n in range(len(xs)):
                    assert lx[n] == len(xs[n]) and ly[n] == len(ys[n])
                    d0, oi, oj, c0 = oracle.fastdtw(xs[n], ys[n], radius=radius, kind=kname)
                    assert L[n] == len(oi), (D, kind, radius, n)
                    assert np.array_equal(pi[n, :L[n]], oi) and np.array_equal(pj[n, :L[n]], oj), (D, kind, radius, n)
                    assert d[n] == d0 and cells[n] == c0


def test_aligner_matches_reference_semantics():
    """reference tests/test_preprocessing.py:441-457 (shapes after forced growth) + docstring case."""
    from nnmnkwii_b200.metrics import melcd
    from nnmnkwii_b200.preprocessing.alignment import DTWAligner
    xs = [_series(t, 5, i) for i, t in enumerate((35, 40, 39))]
    X = _pad(xs, 40)
    Y = X.copy()
    Xa, Ya = DTWAligner().transform((X, Y))  # alignment.py:20-32
    assert Xa.shape == (3, 40, 5) and Ya.shape == (3, 40, 5) and Xa.dtype == X.dtype
    assert np.array_equal(Xa, X) and np.array_equal(Ya, Y)
    # shift Y by 5 frames -> paths longer than T -> the padded length grows (test_preprocessing.py:444-451)
    Y2 = np.zeros_like(X)
    for i, x in enumerate(xs):
        Y2[i, 5:len(x)] = x[:len(x) - 5]
        Y2[i, :5] = x[0] + 1.0
    for dist in (None, melcd):
        al = DTWAligner(verbose=0) if dist is None else DTWAligner(dist=dist)
        Xa, Ya = al.transform((X, Y2))
        assert Xa.shape == Ya.shape and Xa.shape[1] >= 40
        kind = "euclid" if dist is None else "melcd"
        for n, x in enumerate(xs):
            _, oi, oj, _ = oracle.fastdtw(x, Y2[n, :oracle.trim_zeros_frames_len(Y2[n])], 1, kind)
            assert np.array_equal(Xa[n, :len(oi)], x[oi]) and np.array_equal(Ya[n, :len(oj)], Y2[n][oj])
            assert not Xa[n, len(oi):].any() and not Ya[n, len(oj):].any()
        # aligned distance < unaligned distance (test_preprocessing.py:464-501)
        assert np.linalg.norm(Xa - Ya) < np.linalg.norm(X - Y2)
    # a user-written callable that computes the default cost (the reference's own default is such a
    # lambda, alignment.py:35) takes the same kernel path as the sentinel default
    Xd, Yd = DTWAligner(dist=lambda a, b: np.linalg.norm(a - b)).transform((X, Y2))
    X0, Y0 = DTWAligner().transform((X, Y2))
    assert np.array_equal(Xd, X0) and np.array_equal(Yd, Y0)
    with pytest.raises(NotImplementedError):
        DTWAligner(dist=lambda a, b: float(np.abs(a - b).sum())).transform((X, Y))
    with pytest.raises(AssertionError):  # alignment.py:42
        DTWAligner().transform((X[0], Y[0]))


def test_iterative_aligner_shapes():
    from nnmnkwii_b200.preprocessing.alignment import IterativeDTWAligner
    xs = [_series(t, 4, i) for i, t in enumerate((35, 40, 39))]
    X = _pad(xs, 40)
    Y = np.zeros_like(X)
    for i, x in enumerate(xs):
        Y[i, 5:len(x)] = x[:len(x) - 5]
        Y[i, :5] = x[0] + 1.0
    Xa, Ya = IterativeDTWAligner(n_iter=1, max_iter_gmm=1, n_components_gmm=1, random_state=0).transform((X, Y))
    assert Xa.shape == Ya.shape and Xa.shape[0] == 3 and Xa.shape[2] == 4
    assert np.linalg.norm(Xa - Ya) < np.linalg.norm(X - Y)


def test_cfg4_size_properties():
    """BASELINE.json configs[3] at reduced pair count for the oracle part: T~800, 25-dim, melcd.
    Oracle-exact on a sample; monotone / boundary / cost-consistency properties on all pairs."""
    rng = np.random.default_rng(8)
    N, D = 32, 25
    xs, ys = [], []
    for n in range(N):
        Tx = int(rng.integers(700, 901))
        x = _series(Tx, D, 1000 + n)
        Ty = int(rng.integers(700, 901))
        src = np.sort(rng.random(Ty)) * (Tx - 1)  # random monotone time warp of x + noise
        y = x[np.round(src).astype(int)] + 0.05 * rng.standard_normal((Ty, D)).astype(np.float32)
        xs.append(x); ys.append(y.astype(np.float32))
    X, Y = _pad(xs, 900), _pad(ys, 900)
    for radius in (1, -1):
        pi, pj, L, d, cells, lx, ly = _align(X, Y, 1, radius)
        for n in range(N):
            a, b = pi[n, :L[n]], pj[n, :L[n]]
            assert a[0] == 0 and b[0] == 0 and a[-1] == lx[n] - 1 and b[-1] == ly[n] - 1
            da, db = np.diff(a), np.diff(b)
            assert ((da >= 0) & (db >= 0) & (da + db >= 1) & (da <= 1) & (db <= 1)).all()
            # accumulated cost along the returned path == reported distance (float64, same order)
            z = xs[n][a].astype(np.float64) - ys[n][b].astype(np.float64)
            c = oracle.LOGDB_CONST * np.sqrt((z * z).sum(-1))
            acc = 0.0
            for v in c:
                acc += v
            assert abs(acc - d[n]) <= 1e-9 * abs(d[n])
        for n in range(N):  # every pair against the C oracle
            d0, oi, oj, c0 = oracle.fastdtw(xs[n], ys[n], radius=radius, kind="melcd")
            assert np.array_equal(pi[n, :L[n]], oi) and np.array_equal(pj[n, :L[n]], oj)
            assert d[n] == d0 and cells[n] == c0
    assert rel_err([1.0], [1.0]) == 0.0
