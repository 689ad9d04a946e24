"""GPU parity of the backward building blocks against the oracle's autograd (torch-CPU fp32 double-checked
in fp64 where noted): data gradient on the tensor-core kernel (LIN mode), padding adjoints, leaky-ReLU mask,
bias gradient, pool adjoint, sub-pixel upsample form."""
import numpy as np
import pytest
import torch

from oracle import networks_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rt():
    from texturemixer_b200.runtime import Runtime
    return Runtime.get(0)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _nmax(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


def _nhwc(a):
    return np.ascontiguousarray(np.transpose(a, (0, 2, 3, 1)))


def _nchw(a):
    return np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2)))


def _oracle_layer_grads(x, w, b, dy, gain, lrelu, up2=False):
    """dL/dx, dL/dw, dL/db of y = act(bias(conv(up2?(x)))) for L = sum(y * dy), in fp64."""
    xt = torch.from_numpy(x).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    bt = torch.from_numpy(b).double().requires_grad_(True)
    xin = R.upscale2d(xt) if up2 else xt
    y = R.apply_bias(R.conv2d(xin, wt, gain), bt)
    if lrelu:
        y = R.leaky_relu(y)
    (y * torch.from_numpy(dy).double()).sum().backward()
    return xt.grad.numpy(), wt.grad.numpy(), bt.grad.numpy(), y.detach().numpy()


DGRAD_CASES = [
    # n, cin, cout, h, w, k, lrelu
    (2, 256, 256, 32, 32, 3, True),
    (3, 64, 128, 8, 8, 3, True),
    (2, 32, 16, 16, 12, 3, False),
    (2, 16, 16, 6, 10, 3, True),
    (5, 512, 512, 4, 4, 3, True),
    (9, 512, 512, 2, 2, 3, True),
    (2, 64, 256, 8, 8, 1, False),
    (1, 256, 64, 96, 96, 3, True),
]


@pytest.mark.parametrize('n,cin,cout,h,w,k,lrelu', DGRAD_CASES)
def test_conv_dgrad_vs_autograd(rt, n, cin, cout, h, w, k, lrelu):
    rng = np.random.RandomState(n + cin + cout + h)
    x = rng.randn(n, cin, h, w).astype(np.float32)
    wt = rng.randn(k, k, cin, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    dy = rng.randn(n, cout, h, w).astype(np.float32)
    dx_want, _, db_want, y = _oracle_layer_grads(x, wt, b, dy, R.SQRT2, lrelu)
    ws = float(R.wscale_of(wt.shape))
    # dz = dy * lrelu'(y) as planes on the zero-ringed grid (+ bias gradient), mask from the fp32 forward output
    dbias = torch.zeros(cout, device='cuda')
    dz, dz_f32 = rt.grad_prepare(_dev(_nhwc(dy)), n, h, w, cout, src_kind=1, y_f32=_dev(_nhwc(y.astype(np.float32))) if lrelu
                                 else None, want_planes=True, want_f32=True, dbias=dbias)
    assert _nmax(dbias.cpu().numpy(), db_want) <= 1e-5
    hp = (dz[0].float() + dz[1].float()).cpu().numpy()
    assert hp.shape == (n, h + 4, w + 4, cout)
    assert np.abs(hp[:, :2]).max() == 0 and np.abs(hp[:, -2:]).max() == 0 and np.abs(hp[:, :, :2]).max() == 0
    assert _nmax(hp[:, 2:-2, 2:-2], dz_f32.cpu().numpy()) <= 2.0 ** -15
    fwd = rt.prepare_weights(_dev(wt), ws, k, cin, cout)
    wtp = rt.transpose_weights(fwd, cout, k * k, cin)
    g = rt.conv_dgrad(dz, n, h, w, cin, cout, k, wtp)
    _, dx = rt.grad_prepare(g, n, h, w, cin, src_kind=0, fold=0 if k == 3 else 2, want_planes=False, want_f32=True)
    torch.cuda.synchronize()
    assert _nmax(_nchw(dx.cpu().numpy()), dx_want) <= 1e-4


def test_grad_prepare_pool_adjoint_and_residual(rt):
    rng = np.random.RandomState(3)
    n, c, h, w = 2, 32, 8, 12
    y = rng.randn(n, c, h, w).astype(np.float32)
    dyp = rng.randn(n, c, h // 2, w // 2).astype(np.float32)
    add = rng.randn(n, c, h, w).astype(np.float32)
    yt = torch.from_numpy(y).double().requires_grad_(True)
    out = R.downscale2d(R.leaky_relu(yt))
    (out * torch.from_numpy(dyp).double()).sum().backward()
    want = yt.grad.numpy() + add * np.where(y > 0, 1.0, 0.2)
    # (dL/d act) = pool adjoint + add, then masked by the activation derivative of y
    _, got = rt.grad_prepare(_dev(_nhwc(dyp)), n, h, w, c, src_kind=2, add=_dev(_nhwc(add)), y_f32=_dev(_nhwc(y)),
                             want_planes=False, want_f32=True)
    assert _nmax(_nchw(got.cpu().numpy()), want) <= 1e-6


@pytest.mark.parametrize('n,cin,cout,h,w', [(2, 64, 32, 8, 8), (2, 32, 16, 16, 8), (1, 64, 64, 4, 6)])
def test_upscale_conv_dgrad_subpixel_form(rt, n, cin, cout, h, w):
    """dL/dx of conv3x3(upscale2d(x)): phase-packed dz on the low-res grid, 4*Cout contraction, REPLICATE adjoint."""
    rng = np.random.RandomState(n + cin + cout)
    x = rng.randn(n, cin, h, w).astype(np.float32)
    wt = rng.randn(3, 3, cin, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    dy = rng.randn(n, cout, 2 * h, 2 * w).astype(np.float32)
    dx_want, _, _, y = _oracle_layer_grads(x, wt, b, dy, R.SQRT2, True, up2=True)
    ws = float(R.wscale_of(wt.shape))
    dz, _ = rt.grad_prepare(_dev(_nhwc(dy)), n, 2 * h, 2 * w, cout, src_kind=1, y_f32=_dev(_nhwc(y.astype(np.float32))),
                            want_planes=True, phase_pack=True)
    assert dz[0].shape == (n, h + 4, w + 4, 4 * cout)
    fwd = rt.prepare_weights(_dev(wt), ws, 3, cin, cout, up2_phase=True)         # [4*cout][9*cin]
    wtp = rt.transpose_weights(fwd, 4 * cout, 9, cin)                              # [cin][9*4*cout]
    g = rt.conv_dgrad(dz, n, h, w, cin, 4 * cout, 3, wtp)
    _, dx = rt.grad_prepare(g, n, h, w, cin, src_kind=0, fold=1, want_planes=False, want_f32=True)
    assert _nmax(_nchw(dx.cpu().numpy()), dx_want) <= 1e-4


WGRAD_CASES = [
    # n, cin, cout, h, w, k
    (2, 256, 256, 32, 32, 3),
    (3, 64, 128, 8, 8, 3),
    (5, 512, 512, 4, 4, 3),
    (9, 128, 64, 2, 2, 3),
    (2, 64, 256, 8, 8, 1),
    (1, 256, 64, 96, 96, 3),
    (4, 64, 64, 32, 32, 3),
]


@pytest.mark.parametrize('n,cin,cout,h,w,k', WGRAD_CASES)
def test_conv_wgrad_vs_autograd(rt, n, cin, cout, h, w, k):
    from texturemixer_b200.runtime import Act
    rng = np.random.RandomState(n + cin + cout + h)
    x = rng.randn(n, cin, h, w).astype(np.float32)
    wt = rng.randn(k, k, cin, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    dy = rng.randn(n, cout, h, w).astype(np.float32)
    _, dw_want, _, _ = _oracle_layer_grads(x, wt, b, dy, R.SQRT2, False)
    ws = float(R.wscale_of(wt.shape))
    xa = rt.split_pack(Act(n, h, w, cin, f32=_dev(_nhwc(x))))
    dz, _ = rt.grad_prepare(_dev(_nhwc(dy)), n, h, w, cout, src_kind=1, want_planes=True)
    dw = torch.zeros(k, k, cin, cout, device='cuda')
    rt.conv_wgrad((xa.hi, xa.lo), dz, n, h, w, cin, cout, k, ws, dw)
    assert _nmax(dw.cpu().numpy(), dw_want) <= 1e-4
    rt.conv_wgrad((xa.hi, xa.lo), dz, n, h, w, cin, cout, k, ws, dw)       # accumulates
    assert _nmax(dw.cpu().numpy(), 2 * dw_want) <= 1e-4
