"""GPU: delta_features vs the reference golden and the oracle."""
import numpy as np
import pytest

import oracle
from conftest import rel_err, windows_set

pytestmark = pytest.mark.gpu


def test_delta_features_matches_reference(golden):
    from nnmnkwii_b200.preprocessing import delta_features
    for wi, ws in enumerate(windows_set()):
        for dt, tol in (("float32", 1e-6), ("float64", 1e-14)):
            x = golden["delta_w%d_%s_x" % (wi, dt)]
            y = delta_features(x, ws)
            assert y.dtype == x.dtype and y.shape == (x.shape[0], x.shape[1] * len(ws))
            assert rel_err(y, golden["delta_w%d_%s_y" % (wi, dt)]) <= tol
    rng = np.random.default_rng(3)
    x = rng.standard_normal((300, 70)).astype(np.float32)
    ws = windows_set()[3]
    assert rel_err(delta_features(x, ws), oracle.delta_features(x, ws)) < 1e-6
    # asymmetric / even-length plain windows follow np.correlate's centring (len // 2)
    wa = [np.array([1.0]), np.array([-1.0, 1.0]), np.array([0.25, -1.0, 0.5, 0.25])]
    assert rel_err(delta_features(x[:50], wa), oracle.delta_features(x[:50], wa)) < 1e-6
    # batched: deltas do not cross utterance boundaries; round trip with MLPG: mlpg(delta(x), 1) == x
    lens = [40, 100, 160]
    yb = delta_features(x, windows_set()[2], lengths=lens)
    off = np.concatenate([[0], np.cumsum(lens)])
    for u in range(3):
        a, b = off[u], off[u + 1]
        assert np.array_equal(yb[a:b], delta_features(x[a:b], windows_set()[2]))
    with pytest.raises(ValueError):
        delta_features(x[:2], windows_set()[2])
