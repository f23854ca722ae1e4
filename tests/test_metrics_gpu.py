"""Length-masked metrics: CUDA reductions through the C ABI vs values
produced by the reference's nnmnkwii.metrics (tests/golden/make_golden.py).

Tolerance: the reference sums in the input dtype (float32 pairwise sums), the kernels accumulate in
float64 -> 2e-6 relative on float32 inputs, 1e-12 on float64 inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _vals(M, X, Y, lens, f0a, f0b, va, vb):
    return np.array([
        M.melcd(X, Y), M.melcd(X, Y, lens), M.melcd(X[0], Y[0]), M.melcd(X[0, 0], Y[0, 0]),
        M.mean_squared_error(X, Y), M.mean_squared_error(X, Y, lens), M.mean_squared_error(X[0, 0], Y[0, 0]),
        M.lf0_mean_squared_error(f0a, va, f0b, vb), M.lf0_mean_squared_error(f0a, va, f0b, vb, lens),
        M.lf0_mean_squared_error(f0a, va, f0b, vb, lens, linear_domain=True),
        M.lf0_mean_squared_error(f0a[0], va[0], f0b[0], vb[0], linear_domain=True),
        M.lf0_mean_squared_error(f0a[:, :, None], va[:, :, None], f0b[:, :, None], vb[:, :, None], lens),
        M.vuv_error(va, vb), M.vuv_error(va, vb, lens), M.vuv_error(va[:, :, None], vb[:, :, None], lens),
        M.melcd(f0a, f0b, lens), M.mean_squared_error(f0a, f0b, lens),
    ], dtype=np.float64)


@pytest.mark.parametrize("name,tol", [("float32", 2e-6), ("float64", 1e-12)])
def test_metrics_match_reference_golden(golden, name, tol):
    import torch
    from nnmnkwii_b200 import metrics as M
    arrs = [golden["met_%s_%s" % (name, k)] for k in ("X", "Y")]
    f = [golden["met_%s_%s" % (name, k)] for k in ("f0a", "f0b", "va", "vb")]
    lens = [int(v) for v in golden["met_lens"]]
    ref = golden["met_%s_vals" % name]
    got = _vals(M, arrs[0], arrs[1], lens, *f)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < tol, (got, ref)
    # CUDA tensors (the reference accepts torch tensors too), lengths as a tensor
    dX, dY = (torch.from_numpy(a).cuda() for a in arrs)
    df = [torch.from_numpy(a).cuda() for a in f]
    got_t = _vals(M, dX, dY, torch.tensor(lens), *df)
    assert np.array_equal(got_t, got)  # deterministic reduction order
    # identical inputs -> exactly 0 (tests/test_metrics.py:8-32)
    assert M.melcd(arrs[0], arrs[0]) == 0.0 and M.mean_squared_error(arrs[0], arrs[0], lens) == 0.0


def test_masked_melcd_large_batch_properties():
    """Aligned-output sized batch: masked value == the mean over the valid frames computed by torch,
    padding content is ignored, non-contiguous inputs are accepted."""
    import torch
    from nnmnkwii_b200 import metrics as M
    g = torch.Generator(device="cuda").manual_seed(3)
    B, T, D = 300, 1100, 25
    X = torch.randn(B, T, D, device="cuda", generator=g)
    Y = X + 0.1 * torch.randn(B, T, D, device="cuda", generator=g)
    lens = torch.randint(0, T + 1, (B,), generator=torch.Generator().manual_seed(1))
    mask = (torch.arange(T)[None, :] < lens[:, None]).cuda()
    frame = (X.double() - Y.double()).pow(2).sum(-1).sqrt()
    want = float(M._logdb_const * (frame * mask).sum() / lens.sum())
    got = M.melcd(X, Y, lens.tolist())
    assert abs(got - want) / want < 1e-9
    Y2 = Y.clone()
    Y2[~mask] = 1e30  # garbage in the padding must not matter
    assert M.melcd(X, Y2, lens.tolist()) == got
    Xt = X.transpose(0, 1).contiguous().transpose(0, 1)  # same values, non-contiguous
    assert M.melcd(Xt, Y, lens.tolist()) == got
    mse_want = float(((X.double() - Y.double()).pow(2).sum(-1) * mask).sum() / (lens.sum() * D)) ** 0.5
    assert abs(M.mean_squared_error(X, Y, lens.tolist()) - mse_want) / mse_want < 1e-9
    # wide feature rows (D > 32: several lanes-strided passes per frame) and D == 1
    Xw = torch.randn(7, 50, 513, device="cuda", generator=g)
    Yw = torch.randn(7, 50, 513, device="cuda", generator=g)
    want = float(M._logdb_const * (Xw.double() - Yw.double()).pow(2).sum(-1).sqrt().mean())
    assert abs(M.melcd(Xw, Yw) - want) / want < 1e-9


def test_masked_melcd_alignment_phases():
    """Tiles start at every 16-byte phase (odd D, odd T) and X / Y may sit at different phases."""
    import torch
    from nnmnkwii_b200 import metrics as M
    g = torch.Generator(device="cuda").manual_seed(11)
    for dt, tol in ((torch.float32, 1e-6), (torch.float64, 1e-12)):
        for D in (1, 3, 25, 26, 59, 127):
            B, T = 9, 301
            X = torch.randn(B, T, D, device="cuda", generator=g, dtype=dt)
            buf = torch.randn(B * T * D + 3, device="cuda", generator=g, dtype=dt)
            lens = [301, 0, 5, 129, 128, 127, 300, 1, 77]
            mask = (torch.arange(T)[None, :] < torch.tensor(lens)[:, None]).cuda()
            for shift in (0, 1, 3):
                Y = buf[shift:shift + B * T * D].view(B, T, D)
                frame = (X.double() - Y.double()).pow(2).sum(-1).sqrt()
                want = float(M._logdb_const * (frame * mask).sum() / sum(lens))
                got = M.melcd(X, Y, lens)
                assert abs(got - want) / want < tol, (dt, D, shift)
