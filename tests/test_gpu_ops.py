"""GPU parity of every libtmx kernel against the CPU oracle (oracle/*.py), called
through the C ABI via texturemixer_b200.runtime.  Run on the B200 box:
    python -m pytest tests -m gpu
Tolerances: bit-exact for layout moves, the integer tile-index gather and the
matte / lerp blends; 1e-5 (normalised max) for the CUDA-core fp32 convs; 1e-4
for the tensor-core bf16x3 convs (north_star bar: 1e-3)."""
import os

import numpy as np
import pytest
import torch

from oracle import interp_ref as I
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


# ---------------------------------------------------------------------- layout
@pytest.mark.parametrize('n,c,h,w', [(2, 3, 5, 7), (3, 128, 32, 32), (1, 33, 2, 2)])
def test_layout_roundtrip(rt, n, c, h, w):
    x = np.random.RandomState(0).randn(n, c, h, w).astype(np.float32)
    y = rt.nchw_to_nhwc(_dev(x))
    assert np.array_equal(y.cpu().numpy(), _nhwc(x))
    z = rt.nhwc_to_nchw(y)
    assert np.array_equal(z.cpu().numpy(), x)


def test_layout_slices_and_broadcast(rt):
    rng = np.random.RandomState(1)
    zg = rng.randn(2, 8, 1, 1).astype(np.float32)
    zl = rng.randn(2, 8, 4, 6).astype(np.float32)
    buf = rt.empty(2, 4, 6, 16)
    rt.nchw_to_nhwc(_dev(zg), out=buf, c_off=0, c_total=16, bcast_hw=(4, 6))
    rt.nchw_to_nhwc(_dev(zl), out=buf, c_off=8, c_total=16)
    want = np.concatenate([np.tile(zg, (1, 1, 4, 6)), zl], axis=1)
    assert np.array_equal(buf.cpu().numpy(), _nhwc(want))
    back = rt.nhwc_to_nchw(buf, c_off=8, c=8)
    assert np.array_equal(back.cpu().numpy(), zl)


def test_split_halo_pack_unpack(rt):
    from texturemixer_b200.runtime import Act
    x = (np.random.RandomState(2).randn(2, 6, 5, 16) * 3).astype(np.float32)       # NHWC
    a = rt.split_pack(Act(2, 6, 5, 16, f32=_dev(x)))
    hi = a.hi.float().cpu().numpy().astype(np.float64)
    lo = a.lo.float().cpu().numpy().astype(np.float64)
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)), mode='reflect').astype(np.float64)
    assert hi.shape == (2, 8, 7, 16)
    assert np.abs(hi + lo - xp).max() <= 2.0 ** -16 * np.abs(xp).max()
    b = rt.split_unpack(Act(2, 6, 5, 16, hi=a.hi, lo=a.lo))
    assert np.abs(b.f32.cpu().numpy() - x).max() <= 2.0 ** -16 * np.abs(x).max()


@pytest.mark.parametrize('halo', ['reflect', 'replicate', 'zero'])
def test_avgpool2_pack_equals_pool_then_pack(rt, halo):
    """downscale2d whose consumer is a tensor-core conv: one kernel writes the pooled fp32 map AND its split planes with
    the consumer's halo - bit-identical to avgpool2 followed by split_halo_pack, and equal to the oracle's pooling."""
    from texturemixer_b200.runtime import Act
    x = (np.random.RandomState(5).randn(3, 12, 20, 24) * 2).astype(np.float32)       # NHWC
    a = rt.avgpool2(Act(3, 12, 20, 24, f32=_dev(x)))
    rt.split_pack(a, halo)
    b = rt.avgpool2(Act(3, 12, 20, 24, f32=_dev(x)), pack=halo)
    assert b.hi is not None and b.halo == halo and tuple(b.hi.shape) == (3, 8, 12, 24)
    assert torch.equal(a.f32, b.f32)
    assert torch.equal(a.hi.view(torch.int16), b.hi.view(torch.int16))
    assert torch.equal(a.lo.view(torch.int16), b.lo.view(torch.int16))
    want = R.downscale2d(torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))).numpy().transpose(0, 2, 3, 1)
    assert np.abs(b.f32.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max()


# ---------------------------------------------------------------------- convs
def _oracle_conv(x, w, b, gain, lrelu, residual=None, up2=False):
    xt = torch.from_numpy(x)
    if up2:
        xt = R.upscale2d(xt)
    y = R.apply_bias(R.conv2d(xt, torch.from_numpy(w), gain), torch.from_numpy(b))
    if lrelu:
        y = R.leaky_relu(y)
    if residual is not None:
        y = y + torch.from_numpy(residual)
    return y.numpy()


def _run_conv(rt, x, w, b, gain, lrelu, residual, up2, algo, want_split=False, up2_out=False):
    from texturemixer_b200.runtime import Act
    n, cin, h, wd = x.shape
    k, cout = w.shape[0], w.shape[3]
    a = Act(n, h, wd, cin, f32=_dev(_nhwc(x)))
    ws = float(R.wscale_of(w.shape, gain))
    res = None if residual is None else _dev(_nhwc(residual))
    out = rt.conv2d(a, _dev(w), _dev(b), ws, k, cout, lrelu=lrelu, residual=res, up2=up2, want_f32=True,
                    want_split=want_split, up2_out=up2_out, algo=algo)
    torch.cuda.synchronize()
    return out


FFMA_CASES = [
    # n, cin, cout, h, w, k, lrelu, residual, up2
    (2, 16, 16, 12, 10, 3, True, False, False),
    (1, 32, 16, 8, 8, 3, True, False, True),
    (3, 16, 32, 5, 7, 3, False, True, False),
    (2, 64, 256, 6, 6, 1, False, False, False),
    (2, 512, 256, 1, 1, 1, False, False, False),
    (1, 32, 32, 2, 2, 3, True, False, False),
    (2, 48, 20, 9, 4, 3, True, False, False),
]


@pytest.mark.parametrize('n,cin,cout,h,w,k,lrelu,has_res,up2', FFMA_CASES)
def test_conv_ffma_vs_oracle(rt, n, cin, cout, h, w, k, lrelu, has_res, up2):
    from texturemixer_b200 import _lib
    rng = np.random.RandomState(n * 1000 + cin + cout + h)
    x = rng.randn(n, cin, h, w).astype(np.float32)
    wt = rng.randn(k, k, cin, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    f = 2 if up2 else 1
    res = rng.randn(n, cout, h * f, w * f).astype(np.float32) if has_res else None
    want = _oracle_conv(x, wt, b, R.SQRT2, lrelu, res, up2)
    out = _run_conv(rt, x, wt, b, R.SQRT2, lrelu, res, up2, _lib.ALGO_FFMA)
    got = _nchw(out.f32.cpu().numpy())
    assert got.shape == want.shape
    assert _nmax(got, want) <= 1e-5


TC_CASES = [
    # n, cin, cout, h, w, k, lrelu, residual, algo('tc'|'k32')
    (2, 256, 256, 32, 32, 3, True, False, 'tc'),
    (1, 256, 256, 32, 32, 3, False, True, 'tc'),
    (3, 64, 128, 8, 8, 3, True, False, 'tc'),
    (5, 512, 512, 4, 4, 3, True, False, 'tc'),
    (33, 512, 512, 2, 2, 3, True, False, 'tc'),
    (2, 256, 64, 32, 32, 3, True, False, 'tc'),
    (2, 64, 32, 64, 64, 3, True, False, 'tc'),
    (2, 64, 64, 16, 16, 3, True, False, 'k32'),
    (2, 32, 32, 16, 8, 3, True, False, 'k32'),
    (2, 64, 256, 32, 32, 1, False, False, 'tc'),
    (1, 128, 128, 96, 96, 3, True, False, 'tc'),
    (2, 128, 64, 12, 20, 3, True, True, 'tc'),
    (2, 16, 16, 32, 32, 3, True, False, 'tc'),        # KC=16 (32B swizzle), BN=16
    (1, 16, 32, 128, 128, 3, True, False, 'tc'),
    (2, 32, 16, 16, 16, 3, True, True, 'tc'),
    (3, 48, 48, 8, 8, 3, False, False, 'tc'),
    (2, 64, 16, 8, 16, 3, True, False, 'tc'),
    (2, 32, 32, 64, 64, 3, True, False, 'tc'),
]


@pytest.mark.parametrize('n,cin,cout,h,w,k,lrelu,has_res,algo', TC_CASES)
def test_conv_tc_vs_oracle(rt, n, cin, cout, h, w, k, lrelu, has_res, algo):
    from texturemixer_b200 import _lib
    from texturemixer_b200.runtime import Act
    rng = np.random.RandomState(n * 1000 + cin + cout + h)
    x = rng.randn(n, cin, h, w).astype(np.float32)
    wt = rng.randn(k, k, cin, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    res = rng.randn(n, cout, h, w).astype(np.float32) if has_res else None
    want = _oracle_conv(x, wt, b, R.SQRT2, lrelu, res)
    out = _run_conv(rt, x, wt, b, R.SQRT2, lrelu, res, False, _lib.ALGO_TC if algo == 'tc' else _lib.ALGO_TC_K32,
                    want_split=True)
    got = _nchw(out.f32.cpu().numpy())
    assert _nmax(got, want) <= 1e-4
    # the split-plane output carries the same values (16 mantissa bits) and the REFLECT halo
    un = rt.split_unpack(Act(n, h, w, cout, hi=out.hi, lo=out.lo))
    assert _nmax(_nchw(un.f32.cpu().numpy()), want) <= 1e-4
    hp = out.hi.float().cpu().numpy() + out.lo.float().cpu().numpy()
    wantp = np.pad(_nhwc(want), ((0, 0), (1, 1), (1, 1), (0, 0)), mode='reflect')
    assert _nmax(hp, wantp) <= 1e-4


def test_conv_tc_matches_ffma_closely(rt):
    """Device-side cross-check: the tensor-core bf16x3 result against the exact-fp32 CUDA-core kernel."""
    from texturemixer_b200 import _lib
    rng = np.random.RandomState(5)
    x = rng.randn(4, 256, 32, 32).astype(np.float32)
    wt = rng.randn(3, 3, 256, 256).astype(np.float32)
    b = (0.1 * rng.randn(256)).astype(np.float32)
    a = _run_conv(rt, x, wt, b, 1.0, False, None, False, _lib.ALGO_FFMA).f32.cpu().numpy()
    t = _run_conv(rt, x, wt, b, 1.0, False, None, False, _lib.ALGO_TC).f32.cpu().numpy()
    assert _nmax(t, a) <= 5e-5


def test_conv_tc_up2_out(rt):
    from texturemixer_b200 import _lib
    rng = np.random.RandomState(6)
    x = rng.randn(2, 64, 16, 16).astype(np.float32)
    wt = rng.randn(3, 3, 64, 64).astype(np.float32)
    b = (0.1 * rng.randn(64)).astype(np.float32)
    want = R.upscale2d(torch.from_numpy(_oracle_conv(x, wt, b, R.SQRT2, True))).numpy()
    out = _run_conv(rt, x, wt, b, R.SQRT2, True, None, False, _lib.ALGO_TC, want_split=True, up2_out=True)
    hp = out.hi.float().cpu().numpy() + out.lo.float().cpu().numpy()
    assert hp.shape == (2, 34, 34, 64)
    wantp = np.pad(_nhwc(want), ((0, 0), (1, 1), (1, 1), (0, 0)), mode='reflect')
    assert _nmax(hp, wantp) <= 1e-4


@pytest.mark.parametrize('n,cin,cout,h,w', [(2, 64, 32, 16, 16), (2, 32, 16, 32, 32), (1, 64, 64, 8, 4),
                                            (3, 16, 16, 6, 10), (2, 256, 64, 4, 4)])
def test_conv_tc_upscale_subpixel_form(rt, n, cin, cout, h, w):
    """conv3x3(upscale2d(x)) on the tensor-core kernel: low-res REPLICATE-halo planes, 4*Cout phase GEMM."""
    from texturemixer_b200 import _lib
    from texturemixer_b200.runtime import Act
    rng = np.random.RandomState(n + cin + cout + h)
    x = rng.randn(n, cin, h, w).astype(np.float32)
    wt = rng.randn(3, 3, cin, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    want = _oracle_conv(x, wt, b, R.SQRT2, True, None, up2=True)
    out = _run_conv(rt, x, wt, b, R.SQRT2, True, None, True, _lib.ALGO_TC, want_split=True)
    assert out.f32.shape == (n, 2 * h, 2 * w, cout)
    assert _nmax(_nchw(out.f32.cpu().numpy()), want) <= 1e-4
    hp = out.hi.float().cpu().numpy() + out.lo.float().cpu().numpy()
    wantp = np.pad(_nhwc(want), ((0, 0), (1, 1), (1, 1), (0, 0)), mode='reflect')
    assert _nmax(hp, wantp) <= 1e-4
    # the exact-fp32 CUDA-core kernel reading through the upsampling agrees
    ff = _run_conv(rt, x, wt, b, R.SQRT2, True, None, True, _lib.ALGO_FFMA)
    assert _nmax(_nchw(ff.f32.cpu().numpy()), want) <= 1e-5


def test_conv_tc_replicate_halo_and_split_pack(rt):
    from texturemixer_b200 import _lib
    from texturemixer_b200.runtime import Act
    rng = np.random.RandomState(21)
    x = rng.randn(2, 32, 8, 12).astype(np.float32)
    wt = rng.randn(3, 3, 32, 32).astype(np.float32)
    b = (0.1 * rng.randn(32)).astype(np.float32)
    want = _oracle_conv(x, wt, b, R.SQRT2, True)
    a = Act(2, 8, 12, 32, f32=_dev(_nhwc(x)))
    out = rt.conv2d(a, _dev(wt), _dev(b), float(R.wscale_of(wt.shape)), 3, 32, lrelu=True, want_f32=False,
                    want_split=True, halo_out='replicate', algo=_lib.ALGO_TC)
    hp = out.hi.float().cpu().numpy() + out.lo.float().cpu().numpy()
    wantp = np.pad(_nhwc(want), ((0, 0), (1, 1), (1, 1), (0, 0)), mode='edge')
    assert out.halo == 'replicate' and _nmax(hp, wantp) <= 1e-4
    b2 = rt.split_pack(Act(2, 8, 12, 32, f32=_dev(_nhwc(x))), 'replicate')
    hp2 = b2.hi.float().cpu().numpy().astype(np.float64) + b2.lo.float().cpu().numpy()
    xp = np.pad(_nhwc(x), ((0, 0), (1, 1), (1, 1), (0, 0)), mode='edge')
    assert np.abs(hp2 - xp).max() <= 2.0 ** -16 * np.abs(xp).max()


@pytest.mark.parametrize('cout,tanh', [(16, True), (32, False)])
def test_conv_tc_fused_torgb(rt, cout, tanh):
    from texturemixer_b200 import _lib
    from texturemixer_b200.runtime import Act
    rng = np.random.RandomState(22)
    n, cin, h, w = 2, 16, 32, 16
    x = rng.randn(n, cin, h, w).astype(np.float32)
    wt = rng.randn(3, 3, cin, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    wr = rng.randn(1, 1, cout, 3).astype(np.float32)
    br = (0.1 * rng.randn(3)).astype(np.float32)
    y = _oracle_conv(x, wt, b, R.SQRT2, True)
    want = _oracle_conv(y, wr, br, 1.0, False)
    if tanh:
        want = np.tanh(want)
    a = Act(n, h, w, cin, f32=_dev(_nhwc(x)))
    out, img = rt.conv2d(a, _dev(wt), _dev(b), float(R.wscale_of(wt.shape)), 3, cout, lrelu=True, want_f32=True,
                         algo=_lib.ALGO_TC, torgb=(_dev(wr), _dev(br), float(R.wscale_of(wr.shape, 1.0)), 3, tanh))
    assert _nmax(_nchw(out.f32.cpu().numpy()), y) <= 1e-4
    assert img.shape == (n, 3, h, w) and _nmax(img.cpu().numpy(), want) <= 1e-4


# ---------------------------------------------------------------------- pointwise
def test_fromrgb_torgb_avgpool(rt):
    from texturemixer_b200.runtime import Act
    rng = np.random.RandomState(7)
    img = rng.uniform(-1, 1, (3, 3, 16, 12)).astype(np.float32)
    w = rng.randn(1, 1, 3, 16).astype(np.float32)
    b = (0.1 * rng.randn(16)).astype(np.float32)
    want = _oracle_conv(img, w, b, R.SQRT2, True)
    got = rt.fromrgb(_dev(img), _dev(w), _dev(b), float(R.wscale_of(w.shape)), 16, lrelu=True)
    assert _nmax(_nchw(got.f32.cpu().numpy()), want) <= 1e-6

    x = rng.randn(3, 16, 16, 12).astype(np.float32)
    w2 = rng.randn(1, 1, 16, 3).astype(np.float32)
    b2 = (0.1 * rng.randn(3)).astype(np.float32)
    want2 = np.tanh(_oracle_conv(x, w2, b2, 1.0, False))
    a = Act(3, 16, 12, 16, f32=_dev(_nhwc(x)))
    got2 = rt.torgb(a, _dev(w2), _dev(b2), float(R.wscale_of(w2.shape, 1.0)), 3, True)
    assert _nmax(got2.cpu().numpy(), want2) <= 2e-6

    want3 = R.downscale2d(torch.from_numpy(x)).numpy()
    got3 = rt.avgpool2(a)
    assert _nmax(_nchw(got3.f32.cpu().numpy()), want3) <= 1e-6


# ---------------------------------------------------------------------- latent blend (bit-exact)
def test_tiling_permutation_bit_exact(rt):
    from texturemixer_b200 import interp
    rng = np.random.RandomState(8)
    x = rng.randn(3, 128, 32, 32).astype(np.float32)
    np.random.seed(1000)
    idx = I.sample_schedule_indices(3, latent_res=32, scale_h=3, scale_w=3)
    want = I.tiling_permutation_gather(x, idx['h_forward'], idx['w_forward'])
    got = interp.tiling_permutation(_dev(x), 3, 3, idx['h_forward'], idx['w_forward'])
    assert np.array_equal(got.cpu().numpy(), want)
    # the reference's matrix-form arguments give the same canvas
    ph = np.stack([I.index_to_matrix_h(r) for r in idx['h_forward']])[:, None]
    pw = np.stack([I.index_to_matrix_w(c) for c in idx['w_forward']])[:, None]
    got2 = interp.tiling_permutation(_dev(x), 3, 3, ph, pw)
    assert np.array_equal(got2.cpu().numpy(), want)


def test_tiling_permutation_ragged(rt):
    from texturemixer_b200 import interp
    rng = np.random.RandomState(9)
    x = rng.randn(2, 5, 8, 8).astype(np.float32)
    np.random.seed(3)
    idx = I.sample_schedule_indices(2, latent_res=8, scale_h=2, scale_w=5)
    want = I.tiling_permutation_gather(x, idx['h_forward'], idx['w_forward'])
    got = interp.tiling_permutation(_dev(x), 2, 5, idx['h_forward'], idx['w_forward'])
    assert np.array_equal(got.cpu().numpy(), want)


def test_lerp_bit_exact(rt):
    from texturemixer_b200 import interp
    rng = np.random.RandomState(10)
    a = rng.randn(4, 16, 12, 12).astype(np.float32)
    b = rng.randn(4, 16, 12, 12).astype(np.float32)
    t = rng.uniform(0, 1, (4, 1, 1, 1)).astype(np.float32)
    want = a + (b - a) * t
    got = interp.lerp(_dev(a), _dev(b), _dev(t))
    assert np.array_equal(got.cpu().numpy(), want)


def test_blend_corners_bit_exact(rt):
    from texturemixer_b200 import interp
    rng = np.random.RandomState(11)
    n, c, r, sh, sw = 2, 16, 8, 3, 4
    H, W = r * sh, r * sw
    srcs = [rng.randn(n, c, r, r).astype(np.float32) for _ in range(4)]
    zgs = [rng.randn(n, c, 1, 1).astype(np.float32) for _ in range(4)]
    np.random.seed(4)
    idx_h = [np.stack([I.sample_index_h(H, 3, np.random.uniform) for _ in range(n)]) for _ in range(4)]
    idx_w = [np.stack([I.sample_index_w(W, 3, np.random.uniform) for _ in range(n)]) for _ in range(4)]
    mattes = I.linkern_for_weight_arbitrary_shape(H, W, r)
    canv = [I.tiling_permutation_gather(s, ih, iw) for s, ih, iw in zip(srcs, idx_h, idx_w)]
    want_zl = I.blend4(canv, mattes).astype(np.float32)
    want_zg = I.blend4([np.tile(z, (1, 1, H, W)) for z in zgs], mattes).astype(np.float32)
    zg, zl = interp.interpolate([_dev(z) for z in zgs], [_dev(s) for s in srcs], sh, sw, idx_h=idx_h, idx_w=idx_w,
                                latent_res=r)
    assert np.array_equal(zl.cpu().numpy(), want_zl)
    assert np.array_equal(zg.cpu().numpy(), want_zg)


def test_blend_nhwc_output_matches_nchw(rt):
    from texturemixer_b200 import _lib
    rng = np.random.RandomState(12)
    x = rng.randn(2, 128, 8, 8).astype(np.float32)
    idx_h = np.stack([rng.permutation(24) for _ in range(2)]).astype(np.int32)
    idx_w = np.stack([rng.permutation(40) for _ in range(2)]).astype(np.int32)
    buf = torch.zeros(2, 24, 40, 256, device='cuda')
    out = rt.latent_blend([_dev(x)], 24, 40, _lib.BLEND_COPY, idx_h=[_dev(idx_h)], idx_w=[_dev(idx_w)],
                          out_nhwc=buf, c_off=128, c_total=256)
    torch.cuda.synchronize()
    assert np.array_equal(buf[..., 128:].cpu().numpy(), _nhwc(out.cpu().numpy()))
    assert float(buf[..., :128].abs().max()) == 0.0


def test_errors_are_loud(rt):
    from texturemixer_b200 import _lib
    from texturemixer_b200.runtime import Act
    a = Act(1, 4, 4, 24, f32=torch.zeros(1, 4, 4, 24, device='cuda'))
    with pytest.raises(RuntimeError, match='Cin'):
        rt.conv2d(a, torch.zeros(3, 3, 24, 16, device='cuda'), None, 1.0, 3, 16, algo=_lib.ALGO_FFMA)
    n0 = rt.launch_count()
    rt.avgpool2(Act(1, 4, 4, 8, f32=torch.zeros(1, 4, 4, 8, device='cuda')))
    assert rt.launch_count() == n0 + 1


def test_weighted_sum_bit_exact(rt):
    """interp.weighted_sum == numpy's `np.sum(latents * weights, axis=0)` (util_scripts.py:1262,1268: float32 latents
    promoted to the float64 RBF weights, rounded once) and the float32 horizontal matte of :1337,1342."""
    from texturemixer_b200 import interp
    rng = np.random.RandomState(8)
    k, n, c, h, w = 5, 2, 6, 24, 40
    zl = [rng.randn(n, c, h, w).astype(np.float32) for _ in range(k)]
    zg = [rng.randn(n, c, 1, 1).astype(np.float32) for _ in range(k)]
    wts = np.stack([interp.gkern_for_weight_grid_shape_hybridization(h, w, 3.0 * i, 2.0 * i, 8.0, 6.0) for i in range(k)])
    wts = wts / wts.sum(axis=0, keepdims=True)
    got = interp.weighted_sum([torch.from_numpy(a).cuda() for a in zl], wts).cpu().numpy()
    want = np.zeros((n, c, h, w))
    for a, m in zip(zl, wts):
        want = want + a * m[None, None]
    assert np.array_equal(got, want.astype(np.float32))
    got_g = interp.weighted_sum([torch.from_numpy(a).cuda() for a in zg], wts).cpu().numpy()
    want_g = np.zeros((n, c, h, w))
    for a, m in zip(zg, wts):
        want_g = want_g + np.tile(a, (1, 1, h, w)) * m[None, None]
    assert np.array_equal(got_g, want_g.astype(np.float32))
    matt = interp.linkern_for_weight_horizontal([n, c, h, w], 8)                 # float32 matte: float32 arithmetic
    got_h = interp.weighted_sum([torch.from_numpy(zl[0]).cuda(), torch.from_numpy(zl[1]).cuda()],
                                np.stack([matt[0, 0], (np.float32(1.0) - matt)[0, 0]]), math_f32=True).cpu().numpy()
    assert np.array_equal(got_h, zl[0] * matt + zl[1] * (np.float32(1.0) - matt))
