"""GPU: autograd.MLPG / UnitVarianceMLPG, mirroring the reference's tests/test_autograd.py
(gradcheck eps=1e-3 atol=1e-3 on float32, minibatch == per-item, variance expansion) plus parity
with the golden fwd/bwd vectors produced by the reference's own autograd Functions."""
import numpy as np
import pytest
import torch
from torch.autograd import gradcheck

import oracle
from conftest import rel_err, windows_set

pytestmark = pytest.mark.gpu


def _mods():
    from nnmnkwii_b200 import autograd as AF
    from nnmnkwii_b200 import paramgen as G
    return AF, G


def test_golden_unit_variance_fwd_bwd(golden):
    AF, G = _mods()
    ws = windows_set()[2]
    mu = torch.from_numpy(golden["uv_means"]).requires_grad_(True)
    R = torch.from_numpy(G.unit_variance_mlpg_matrix(ws, mu.shape[1]))
    y = AF.unit_variance_mlpg(R, mu)
    assert rel_err(y.detach().numpy(), golden["uv_y"]) < 1e-5
    (y * torch.from_numpy(golden["uv_wgt"])).sum().backward()
    assert rel_err(mu.grad.numpy(), golden["uv_grad"]) < 1e-5
    # CUDA tensors stay on the device
    muc = torch.from_numpy(golden["uv_means"]).cuda().requires_grad_(True)
    yc = AF.unit_variance_mlpg(R.cuda(), muc)
    assert yc.is_cuda and rel_err(yc.detach().cpu().numpy(), golden["uv_y"]) < 1e-5


def test_golden_mlpg_fwd_bwd(golden):
    AF, G = _mods()
    ws = windows_set()[2]
    mu = torch.from_numpy(golden["ag_means"]).requires_grad_(True)
    y = AF.mlpg(mu, torch.from_numpy(golden["ag_vars"]), ws)
    assert y.dtype == torch.float32 and rel_err(y.detach().numpy(), golden["ag_y"]) < 1e-6
    (y * torch.from_numpy(golden["ag_wgt"])).sum().backward()
    assert rel_err(mu.grad.numpy(), golden["ag_grad"]) < 2e-6


def test_functional_mlpg_equivalences():
    """reference tests/test_autograd.py:39-72"""
    AF, G = _mods()
    static_dim, T = 2, 10
    for windows in windows_set():
        torch.manual_seed(1234)
        means = torch.rand(T, static_dim * len(windows))
        variances = torch.ones(static_dim * len(windows))
        y = G.mlpg(means.numpy(), variances.numpy(), windows)
        y = torch.from_numpy(y).clone()
        means = means.clone().requires_grad_(True)
        y_hat = AF.mlpg(means, variances, windows)
        assert np.allclose(y.numpy(), y_hat.detach().numpy(), atol=1e-6)
        torch.nn.MSELoss()(y_hat, y).backward()
        R = torch.from_numpy(G.unit_variance_mlpg_matrix(windows, T))
        y_hat = AF.unit_variance_mlpg(R, means)
        assert np.allclose(y.numpy(), y_hat.detach().numpy(), atol=1e-5)
        torch.nn.MSELoss()(y_hat, y).backward()
        y_hat = AF.unit_variance_mlpg(R, means.view(1, -1, means.size(-1)))
        assert np.allclose(y.numpy(), y_hat.detach().numpy()[0], atol=1e-5)


def test_unit_variance_mlpg_gradcheck():
    """reference tests/test_autograd.py:75-113"""
    AF, G = _mods()
    static_dim, T = 2, 10
    for windows in windows_set():
        torch.manual_seed(1234)
        means = torch.rand(T, static_dim * len(windows), requires_grad=True)
        variances = torch.ones(static_dim * len(windows)).expand(T, static_dim * len(windows))
        y1 = AF.MLPG.apply(means, variances, windows)
        R = torch.from_numpy(G.unit_variance_mlpg_matrix(windows, T))
        y2 = AF.UnitVarianceMLPG.apply(means, R)
        assert np.allclose(y1.detach().numpy(), y2.detach().numpy(), atol=1e-5)
        assert gradcheck(AF.UnitVarianceMLPG.apply, (means, R), eps=1e-3, atol=1e-3)
        reshaped = torch.from_numpy(G.reshape_means(means.detach().numpy(), static_dim)).requires_grad_(True)
        assert gradcheck(AF.UnitVarianceMLPG.apply, (reshaped, R), eps=1e-3, atol=1e-3)


def test_minibatch_unit_variance_mlpg():
    """reference tests/test_autograd.py:116-177: 3-D batch == per item; stride-0 expanded batch."""
    AF, G = _mods()
    static_dim, T, batch_size = 2, 5, 3
    for windows in windows_set():
        torch.manual_seed(0)
        means = torch.rand(T, static_dim * len(windows), requires_grad=True)
        means_expanded = means.expand(batch_size, means.shape[0], means.shape[1])
        reshaped = torch.from_numpy(G.reshape_means(means.detach().numpy(), static_dim)).requires_grad_(True)
        reshaped_expanded = reshaped.expand(batch_size, reshaped.shape[0], reshaped.shape[1])
        R = torch.from_numpy(G.unit_variance_mlpg_matrix(windows, T))
        y = AF.unit_variance_mlpg(R, means)
        y_b = AF.unit_variance_mlpg(R, means_expanded)
        y_rb = AF.unit_variance_mlpg(R, reshaped_expanded)
        assert y_b.shape == (batch_size, T, static_dim)
        for i in range(batch_size):
            assert np.allclose(y.detach().numpy(), y_b[i].detach().numpy(), atol=1e-6)
            assert np.allclose(y.detach().numpy(), y_rb[i].detach().numpy(), atol=1e-6)
        # gradients
        nlf = torch.nn.MSELoss()
        tgt = torch.rand(T, static_dim)
        means.grad = None
        nlf(AF.unit_variance_mlpg(R, means), tgt).backward()
        g1 = means.grad.clone()
        mb = means.detach().clone().expand(batch_size, T, means.shape[1]).contiguous().requires_grad_(True)
        nlf(AF.unit_variance_mlpg(R, mb), tgt.expand(batch_size, T, static_dim)).backward()
        for i in range(batch_size):
            assert np.allclose(g1.numpy(), mb.grad[i].numpy() * batch_size, atol=1e-6)


def test_mlpg_gradcheck():
    """reference tests/test_autograd.py:180-204 (stride-0 expanded / random variances)"""
    AF, G = _mods()
    static_dim, T = 2, 10
    for windows in windows_set():
        torch.manual_seed(1234)
        means = torch.rand(T, static_dim * len(windows), requires_grad=True)
        variances = torch.ones(static_dim * len(windows)).expand(T, static_dim * len(windows))
        assert gradcheck(AF.MLPG.apply, (means, variances, windows), eps=1e-3, atol=1e-3)
        variances = torch.rand(static_dim * len(windows)).expand(T, static_dim * len(windows))
        assert gradcheck(AF.MLPG.apply, (means, variances, windows), eps=1e-3, atol=1e-3)


def test_mlpg_variance_expand():
    """reference tests/test_autograd.py:207-218"""
    AF, G = _mods()
    static_dim, T = 2, 10
    for windows in windows_set():
        torch.manual_seed(1234)
        means = torch.rand(T, static_dim * len(windows), requires_grad=True)
        variances = torch.rand(static_dim * len(windows))
        y1 = AF.mlpg(means, variances, windows)
        y2 = AF.mlpg(means, variances.expand(T, static_dim * len(windows)), windows)
        assert np.allclose(y1.detach().numpy(), y2.detach().numpy())


def test_cfg3_full_size_unit_variance():
    """BASELINE.json configs[2]: batch=64, T=1000, static_dim=60, fwd + loss.backward() on the GPU.
    Checked against the float64 dense product with the SAME R on a slice, and by the adjoint
    identity <R x, o> == <x, R^T o> on everything."""
    AF, G = _mods()
    ws = windows_set()[2]
    T, sd, B = 1000, 60, 64
    R = torch.from_numpy(G.unit_variance_mlpg_matrix(ws, T)).cuda()
    assert np.abs(R[:, :].cpu().numpy()[::97] - oracle.unit_variance_mlpg_matrix(ws, T)[::97]).max() < 2e-7
    g = torch.Generator(device="cuda").manual_seed(1)
    mu = torch.randn(B, T, 3 * sd, device="cuda", generator=g).requires_grad_(True)
    y = AF.unit_variance_mlpg(R, mu)
    assert y.shape == (B, T, sd)
    loss = y.pow(2).mean()
    loss.backward()
    # dense float64 check on 2 batch items
    Rd = R.double()
    for b in (0, 63):
        xr = mu[b].detach().double().view(T, 3, sd).transpose(0, 1).reshape(3 * T, sd)
        yref = Rd @ xr
        assert rel_err(y[b].detach().cpu().numpy(), yref.cpu().numpy()) < 1e-5
        gref = (Rd.t() @ (2.0 * yref / (B * T * sd))).view(3, T, sd).transpose(0, 1).reshape(T, 3 * sd)
        assert rel_err(mu.grad[b].cpu().numpy(), gref.cpu().numpy()) < 1e-5
    o = torch.randn(B, T, sd, device="cuda", generator=g)
    lhs = (y.detach().double() * o.double()).sum()
    (gx,) = torch.autograd.grad(AF.unit_variance_mlpg(R, mu), mu, o)
    rhs = (mu.detach().double() * gx.double()).sum()
    assert abs(float(lhs - rhs)) / abs(float(lhs)) < 1e-5


def test_batched_mlpg_autograd_matches_per_utterance():
    """Additive MLPGBatch: one launch per direction for a ragged mini-batch == MLPG applied per utterance."""
    import torch
    from nnmnkwii_b200 import autograd as AF
    ws = windows_set()[2]
    g = torch.Generator().manual_seed(21)
    sd, lens = 37, [50, 1, 33, 128, 7]
    B, Tmax, D = len(lens), max(lens), 3 * sd
    means = torch.randn(B, Tmax, D, generator=g)
    var = torch.rand(B, Tmax, D, generator=g) + 0.1
    w = torch.randn(B, Tmax, sd, generator=g)
    mask = (torch.arange(Tmax)[None, :] < torch.tensor(lens)[:, None]).float()[:, :, None]
    for device in ("cuda", "cpu"):
        for gvar in (False, True):
            v = (var[0, 0] if gvar else var).to(device)
            m = (means * mask).to(device).requires_grad_(True)
            y = AF.mlpg_batch(m, v, ws, lens)
            assert y.shape == (B, Tmax, sd) and y.dtype == torch.float32 and y.device.type == device
            (y * w.to(device) * mask.to(device)).sum().backward()
            for b, n in enumerate(lens):
                mb = means[b, :n].to(device).requires_grad_(True)
                vb = v if gvar else v[b, :n]
                yb = AF.mlpg(mb, vb, ws)
                (yb * w[b, :n].to(device)).sum().backward()
                assert torch.allclose(y[b, :n], yb, rtol=1e-6, atol=1e-6)
                assert torch.allclose(m.grad[b, :n], mb.grad, rtol=1e-5, atol=1e-6)
                assert not y[b, n:].any() and not m.grad[b, n:].any()
    # flat (sum_T, D) form
    flat = torch.cat([means[b, :n] for b, n in enumerate(lens)]).cuda().requires_grad_(True)
    fv = torch.cat([var[b, :n] for b, n in enumerate(lens)]).cuda()
    yf = AF.mlpg_batch(flat, fv, ws, lens)
    yp = AF.mlpg_batch(means.cuda(), var.cuda(), ws, lens)
    off = np.concatenate([[0], np.cumsum(lens)])
    for b, n in enumerate(lens):
        assert torch.equal(yf[off[b]:off[b + 1]], yp[b, :n])


@pytest.mark.parametrize("sd,T,reshaped,nwin", [(59, 500, False, 3), (24, 500, True, 3), (60, 333, False, 3), (5, 260, False, 2),
                                                 (1, 200, True, 3)])
def test_unit_variance_factored_sweeps_match_dense(sd, T, reshaped, nwin):
    """The factored Toeplitz path (one long filter + short window stencils, packed dim pairs) on the
    reference's own perf grid (static_dim 24 / 59, T 500: perf/autograd_mlpg_perf.py:110-120) and odd /
    tiny static dims, both layouts, forward and backward, against the float64 dense product with the
    SAME R."""
    AF, G = _mods()
    from nnmnkwii_b200 import _uvmlpg as uv
    ws = windows_set()[2][:nwin]
    R = torch.from_numpy(G.unit_variance_mlpg_matrix(ws, T)).cuda()
    band = uv.band_of(R, R.device)
    assert band.fact is not None and band.factT is not None, "factored path not selected"
    assert band.fact[0] == 1 and np.allclose(band.fact[2][0], [0, 1, 0], atol=1e-6)
    g = torch.Generator(device="cuda").manual_seed(3)
    B = 3
    mu = torch.randn(B, T, nwin * sd, device="cuda", generator=g)
    x = mu.view(B, T, nwin, sd).transpose(1, 2).reshape(B, nwin * T, sd).contiguous() if reshaped else mu
    x = x.requires_grad_(True)
    y = AF.unit_variance_mlpg(R, x)
    o = torch.randn(B, T, sd, device="cuda", generator=g)
    (gx,) = torch.autograd.grad(y, x, o)
    Rd = R.double()
    for b in range(B):
        xr = mu[b].detach().double().view(T, nwin, sd).transpose(0, 1).reshape(nwin * T, sd)
        yref = Rd @ xr
        assert rel_err(y[b].detach().cpu().numpy(), yref.cpu().numpy()) < 2e-6
        gref = Rd.t() @ o[b].double()  # (nw*T, sd)
        if not reshaped:
            gref = gref.view(nwin, T, sd).transpose(0, 1).reshape(T, nwin * sd)
        assert rel_err(gx[b].cpu().numpy(), gref.cpu().numpy()) < 2e-6


def test_unit_variance_arbitrary_dense_R_is_applied_exactly():
    """an R that is NOT an MLPG matrix (dense random) is never silently banded: K = T - 1, the
    per-row table kernels, result == torch.mm to float32 rounding; float64 R keeps float64 accuracy."""
    AF, G = _mods()
    g = torch.Generator(device="cuda").manual_seed(9)
    T, sd = 70, 5
    for dt, tol in ((torch.float32, 2e-6), (torch.float64, 1e-13)):
        R = torch.randn(T, 3 * T, device="cuda", generator=g, dtype=dt)
        x = torch.randn(2, 3 * T, sd, device="cuda", generator=g, dtype=dt).requires_grad_(True)
        y = AF.unit_variance_mlpg(R, x)
        ref = torch.matmul(R.double(), x.detach().double())
        assert rel_err(y.detach().cpu().numpy(), ref.cpu().numpy()) < tol
        o = torch.randn_like(y)
        (gx,) = torch.autograd.grad(y, x, o)
        gref = torch.matmul(R.double().t(), o.double())
        assert rel_err(gx.cpu().numpy(), gref.cpu().numpy()) < tol


def test_unit_variance_step_is_cuda_graph_capturable():
    """after the one-off band extraction the forward + backward sweeps issue no host
    synchronisation (no .item(), no upload): a whole step can be captured and replayed as a CUDA graph."""
    AF, G = _mods()
    ws = windows_set()[2]
    T, sd, B = 300, 59, 4
    R = torch.from_numpy(G.unit_variance_mlpg_matrix(ws, T)).cuda()
    mu = torch.randn(B, T, 3 * sd, device="cuda").requires_grad_(True)
    o = torch.randn(B, T, sd, device="cuda")
    (g_eager,) = torch.autograd.grad(AF.unit_variance_mlpg(R, mu), mu, o)  # warm-up: builds + caches the band
    y_eager = AF.unit_variance_mlpg(R, mu).detach().clone()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            y = AF.unit_variance_mlpg(R, mu)
            (gx,) = torch.autograd.grad(y, mu, o)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y.detach(), y_eager) and torch.equal(gx, g_eager)
