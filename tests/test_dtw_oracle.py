"""CPU: the DTW oracle.  PARITY UNPINNED w.r.t. the reference (fastdtw is not installable): the C
restatement is checked against the literal pure-Python restatement of fastdtw.py and against the
committed vectors that restatement produced with the REFERENCE's melcd as the cost."""
import numpy as np

import oracle
from oracle import fastdtw_py


def _series(T, D, seed):
    r = np.random.default_rng(seed)
    return (np.cumsum(r.standard_normal((T, D)), 0) * 0.3).astype(np.float32)


def test_dtw_golden(dtw_golden):
    g = dtw_golden
    for case in range(5):
        x, y, radius = g["c%d_x" % case], g["c%d_y" % case], int(g["c%d_radius" % case])
        d, pi, pj, cells = oracle.fastdtw(x, y, radius=radius, kind="melcd")
        path = g["c%d_fast_path" % case]
        assert d == float(g["c%d_fast_dist" % case])
        assert np.array_equal(pi, path[:, 0]) and np.array_equal(pj, path[:, 1])
        assert cells == int(g["c%d_fast_cells" % case])
        d, pi, pj, _ = oracle.dtw(x, y, kind="melcd")
        path = g["c%d_exact_path" % case]
        assert d == float(g["c%d_exact_dist" % case])
        assert np.array_equal(pi, path[:, 0]) and np.array_equal(pj, path[:, 1])


def test_c_vs_literal_python():
    rng = np.random.default_rng(3)
    for it in range(12):
        Tx, Ty, D = int(rng.integers(1, 50)), int(rng.integers(1, 50)), int(rng.integers(1, 6))
        x, y = _series(Tx, D, it), _series(Ty, D, it + 100)
        for radius in (1, 2):
            d0, p0, n0 = fastdtw_py.fastdtw(x, y, radius=radius, dist=oracle.melcd, return_cells=True)
            d1, pi, pj, n1 = oracle.fastdtw(x, y, radius=radius, kind="melcd")
            assert d0 == d1 and n0 == n1 and [tuple(p) for p in p0] == list(zip(pi.tolist(), pj.tolist()))


def test_expand_window_closed_form_vs_literal():
    rng = np.random.default_rng(5)
    for it in range(10):
        cx, cy = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        # random monotone connected coarse path
        i = j = 0
        path = [(0, 0)]
        while (i, j) != (cx - 1, cy - 1):
            moves = [(a, b) for a, b in ((1, 0), (0, 1), (1, 1)) if i + a < cx and j + b < cy]
            a, b = moves[int(rng.integers(len(moves)))]
            i, j = i + a, j + b
            path.append((i, j))
        for radius in (1, 2):
            for len_x, len_y in ((2 * cx, 2 * cy), (2 * cx + 1, 2 * cy + 1)):
                win = fastdtw_py.expand_window(path, len_x, len_y, radius)
                lo, hi = oracle.expand_window([p[0] for p in path], [p[1] for p in path], len_x, len_y, radius)
                rows = {}
                for a, b in win:
                    rows.setdefault(a, []).append(b)
                for r in range(len_x):
                    assert rows[r] == list(range(lo[r], hi[r]))


def test_path_properties_and_ties():
    x = _series(40, 5, 1)
    d, pi, pj, _ = oracle.fastdtw(x, x.copy(), 1, "melcd")  # the docstring case Y = X.copy()
    assert d == 0.0 and np.array_equal(pi, np.arange(40)) and np.array_equal(pj, np.arange(40))
    y = _series(55, 5, 2)
    for radius in (-1, 1, 3):
        d, pi, pj, _ = oracle.fastdtw(x, y, radius, "euclid")
        assert pi[0] == 0 and pj[0] == 0 and pi[-1] == 39 and pj[-1] == 54
        di, dj = np.diff(pi), np.diff(pj)
        assert ((di >= 0) & (dj >= 0) & (di + dj >= 1) & (di <= 1) & (dj <= 1)).all()
    # exact DTW is never worse than the windowed approximation
    assert oracle.dtw(x, y, "melcd")[0] <= oracle.fastdtw(x, y, 1, "melcd")[0] + 1e-12
