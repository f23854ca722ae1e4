"""Literal pure-Python restatement of slaypni/fastdtw's ``fastdtw.py`` (0.3.x).

TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/nnk_oracle.c header).  PARITY UNPINNED: the package
itself is an unpinned, un-vendored dependency of the reference (setup.py:139; call sites
nnmnkwii/preprocessing/alignment.py:50,138) and is not installable here.  This file restates the
package's published pure-Python algorithm (functions ``fastdtw``, ``__fastdtw``, ``dtw``, ``__dtw``,
``__reduce_by_half``, ``__expand_window``) from its documented behaviour so that the C restatement
(nnk_oracle.c: orc_fastdtw) and the CUDA kernels can be cross-checked on small cases:

* inputs coerced to float64;
* recursion: if either series is shorter than ``radius + 2`` run the full DTW, otherwise halve
  both series (mean of adjacent pairs, odd tail dropped), recurse, expand the coarse path by
  ``radius`` and project it to the fine grid, then run DTW restricted to that window;
* DP cell: ``min`` over ((D[i-1,j]+dt), (D[i,j-1]+dt), (D[i-1,j-1]+dt)) with ``key=lambda a: a[0]``,
  i.e. the FIRST minimum wins in the order up, left, diagonal;
* the returned path runs from (0, 0) to (len_x-1, len_y-1).

Only for small inputs (pure-Python loops, dict-of-cells).
"""
from collections import defaultdict

import numpy as np


def _reduce_by_half(x):
    return [(x[i] + x[1 + i]) / 2 for i in range(0, len(x) - len(x) % 2, 2)]


def expand_window(path, len_x, len_y, radius):
    path_ = set(path)
    for i, j in path:
        for a, b in ((i + a, j + b) for a in range(-radius, radius + 1) for b in range(-radius, radius + 1)):
            path_.add((a, b))

    window_ = set()
    for i, j in path_:
        for a, b in ((i * 2, j * 2), (i * 2, j * 2 + 1), (i * 2 + 1, j * 2), (i * 2 + 1, j * 2 + 1)):
            window_.add((a, b))

    window = []
    start_j = 0
    for i in range(0, len_x):
        new_start_j = None
        for j in range(start_j, len_y):
            if (i, j) in window_:
                window.append((i, j))
                if new_start_j is None:
                    new_start_j = j
            elif new_start_j is not None:
                break
        start_j = new_start_j

    return window


def _dtw(x, y, window, dist):
    len_x, len_y = len(x), len(y)
    if window is None:
        window = [(i, j) for i in range(len_x) for j in range(len_y)]
    window = ((i + 1, j + 1) for i, j in window)
    D = defaultdict(lambda: (float("inf"),))
    D[0, 0] = (0, 0, 0)
    ncells = 0
    for i, j in window:
        dt = dist(x[i - 1], y[j - 1])
        D[i, j] = min(
            (D[i - 1, j][0] + dt, i - 1, j),
            (D[i, j - 1][0] + dt, i, j - 1),
            (D[i - 1, j - 1][0] + dt, i - 1, j - 1),
            key=lambda a: a[0],
        )
        ncells += 1
    path = []
    i, j = len_x, len_y
    while not (i == j == 0):
        path.append((i - 1, j - 1))
        i, j = D[i, j][1], D[i, j][2]
    path.reverse()
    return (D[len_x, len_y][0], path, ncells)


def _fastdtw(x, y, radius, dist):
    min_time_size = radius + 2
    if len(x) < min_time_size or len(y) < min_time_size:
        return _dtw(x, y, None, dist)
    x_shrinked = _reduce_by_half(x)
    y_shrinked = _reduce_by_half(y)
    _, path, n0 = _fastdtw(x_shrinked, y_shrinked, radius=radius, dist=dist)
    window = expand_window(path, len(x), len(y), radius)
    d, p, n1 = _dtw(x, y, window, dist)
    return d, p, n0 + n1


def fastdtw(x, y, radius=1, dist=None, return_cells=False):
    x = np.asanyarray(x, dtype="float")
    y = np.asanyarray(y, dtype="float")
    d, p, n = _fastdtw(x, y, radius, dist)
    return (d, p, n) if return_cells else (d, p)


def dtw(x, y, dist=None, return_cells=False):
    x = np.asanyarray(x, dtype="float")
    y = np.asanyarray(y, dtype="float")
    d, p, n = _dtw(x, y, None, dist)
    return (d, p, n) if return_cells else (d, p)
