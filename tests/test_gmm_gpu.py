"""GMM-based conversion in front of MLPG: nnmnkwii_b200.baseline.gmm vs
outputs of the reference's nnmnkwii.baseline.gmm on the same fitted GMM (tests/golden/make_golden.py).
float64 throughout, like the reference; tolerance 1e-9 relative (different but equivalent algebra:
A = Sigma_yx Sigma_xx^-1 is formed once instead of one solve per frame)."""
import types

import numpy as np
import pytest

from conftest import rel_err, windows_set

pytestmark = pytest.mark.gpu


def _gmm(golden):
    return types.SimpleNamespace(means_=golden["gmm_means"], covariances_=golden["gmm_covars"],
                                 weights_=golden["gmm_weights"], covariance_type="full")


def test_gmm_mlpg_matches_reference_golden(golden):
    from nnmnkwii_b200.baseline.gmm import MLPG, MLPGBase
    gmm, src = _gmm(golden), golden["gmm_src"]
    w3 = windows_set()[2]
    static = [(0, 0, np.array([1.0]))]
    cases = [("gmm_default", MLPG(gmm)), ("gmm_w3", MLPG(gmm, windows=w3)), ("gmm_w3_diff", MLPG(gmm, windows=w3, diff=True)),
             ("gmm_w3_swap", MLPG(gmm, windows=w3, swap=True)), ("gmm_static", MLPG(gmm, windows=static))]
    for key, conv in cases:
        y = conv.transform(src)
        assert y.shape == golden[key].shape and y.dtype == golden[key].dtype
        assert rel_err(y, golden[key]) < 1e-9, key
    y32 = MLPG(gmm, windows=static).transform(src.astype(np.float32))
    assert y32.dtype == np.float32 and rel_err(y32, golden["gmm_static_f32"]) < 1e-6
    assert rel_err(MLPGBase(gmm, diff=True).transform(src), golden["gmm_base_2d"]) < 1e-9
    y1 = MLPGBase(gmm).transform(src[3])
    assert y1.shape == (12,) and rel_err(y1, golden["gmm_base_1d"]) < 1e-9
    # batched form == per-utterance form
    conv = MLPG(gmm, windows=w3)
    parts = [src[:7], src[7:8], src[8:40], src[40:]]
    for got, part in zip(conv.transform_batch(parts), parts):
        assert rel_err(got, conv.transform(part)) < 1e-12
    assert conv.transform_batch([]) == []
