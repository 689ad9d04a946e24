"""GPU parity of the network-level path (E_zg / E_zl / G_res through the
reference-style `Network` boundary) against the CPU oracle and the committed
golden vectors minted from the reference's own networks.py
(tests/golden/make_golden.py).  north_star tolerance: 1e-3 relative fp32
(normalised max error); we assert 5e-4 (bf16x3 products carry ~2^-16 relative
error each; 17 stacked convs measure 1-3e-4 normalised max)."""
import os

import numpy as np
import pytest
import torch

from oracle import interp_ref as I
from oracle import networks_ref as R

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SUBSAMPLE = 37
TOL = 5e-4


def _nmax(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


def _make(func, params, **extra):
    from texturemixer_b200.network import Network
    cfg = dict(R.CONFIG[func])
    cfg.update(extra)
    net = Network(func, func='networks.' + func, seed=0, num_channels=3, resolution=128, **cfg)
    assert list(net.vars.keys()) == list(params.keys())
    net.set_vars(params)
    return net


def _inputs(func, rng, n, sh=1, sw=1):
    if func == 'G_res':
        return [rng.randn(n, 128, 32 * sh, 32 * sw).astype(np.float32),
                rng.randn(n, 128, 32 * sh, 32 * sw).astype(np.float32)]
    return [rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)]


@pytest.mark.parametrize('func,n', [('E_zg', 2), ('E_zl', 2), ('G_res', 2)])
def test_network_matches_reference_golden(func, n):
    """Same seeded parameters and inputs as the golden file produced by the
    reference's networks.py; compares the stored (subsampled) outputs."""
    g = np.load(os.path.join(GOLDEN, 'networks.npz'))
    rng = np.random.RandomState(1000)
    params = R.init_params(func, rng, **R.CONFIG[func])
    ins = _inputs(func, rng, n)
    net = _make(func, params)
    outs = net.run(*ins, return_as_list=True)
    for i, a in enumerate(outs):
        assert list(a.shape) == g['%s_out%d_shape' % (func, i)].tolist()
        flat = a.reshape(-1)
        got = flat[::SUBSAMPLE] if flat.size > 4096 else flat
        scale = float(g['%s_out%d_absmax' % (func, i)][0])
        assert np.abs(got - g['%s_out%d' % (func, i)]).max() <= TOL * scale


@pytest.mark.parametrize('algo', ['auto', 'ffma'])
def test_generator_vs_oracle_with_taps(algo, monkeypatch):
    """cfg 2 recipe at a batch the oracle finishes in seconds, plus the pre-tanh check."""
    from texturemixer_b200.runtime import Runtime
    rng = np.random.RandomState(1000)
    params = R.init_params('G_res', rng, **R.CONFIG['G_res'])
    n = 3
    zg = np.tile(rng.randn(n, 128, 1, 1).astype(np.float32), (1, 1, 32, 32))
    zl = rng.randn(n, 128, 32, 32).astype(np.float32)
    with torch.no_grad():
        want = R.G_res(torch.from_numpy(zg), torch.from_numpy(zl), R.to_torch(params), **R.CONFIG['G_res']).numpy()
    net = _make('G_res', params)
    rt = Runtime.get(0)
    monkeypatch.setattr(rt, 'conv_algo', algo)
    got = net.run(zg, zl)
    assert got.shape == (n, 3, 128, 128)
    assert _nmax(got, want) <= (TOL if algo == "auto" else 5e-5)
    # pre-tanh comparison via atanh is ill-conditioned near +-1; compare in tanh space with rtol as well
    ok = np.isclose(got, want, rtol=1e-3, atol=1e-3 * np.abs(want).max())
    assert ok.mean() == 1.0


def test_generator_global_code_tiled_on_the_device():
    """A [N,C,1,1] global code - or the caller's stride-0 np.broadcast_to view of one - gives bit-identical images to
    the host-tiled array (run.py:375 np.tile) at half the H2D bytes; get_output_for takes the same short form."""
    rng = np.random.RandomState(4)
    params = R.init_params('G_res', rng, **R.CONFIG['G_res'])
    net = _make('G_res', params)
    n = 5
    zg1 = rng.randn(n, 128, 1, 1).astype(np.float32)
    zl = rng.randn(n, 128, 32, 32).astype(np.float32)
    want = net.run(np.tile(zg1, (1, 1, 32, 32)), zl, minibatch_size=2)
    assert np.array_equal(net.run(zg1, zl, minibatch_size=2), want)
    assert np.array_equal(net.run(np.broadcast_to(zg1, (n, 128, 32, 32)), zl, minibatch_size=3), want)
    got = net.get_output_for(torch.from_numpy(zg1).cuda(), torch.from_numpy(zl).cuda()).cpu().numpy()
    assert np.array_equal(got, want)


def test_generator_fully_convolutional_scale():
    """G_fcn view (run.py:273): same variables, scale_h x scale_w canvas."""
    from texturemixer_b200.network import Network
    rng = np.random.RandomState(1000)
    params = R.init_params('G_res', rng, **R.CONFIG['G_res'])
    G = _make('G_res', params)
    cfg = dict(R.CONFIG['G_res'], scale_h=2, scale_w=3)
    G_fcn = Network('G', func='networks.G_res', reuse=True, share_vars_with=G, num_channels=3, resolution=128, **cfg)
    assert G_fcn.input_shapes[0] == [None, 128, 64, 96] and G_fcn.output_shape == [None, 3, 256, 384]
    zg, zl = _inputs('G_res', rng, 1, 2, 3)
    with torch.no_grad():
        want = R.G_res(torch.from_numpy(zg), torch.from_numpy(zl), R.to_torch(params), **cfg).numpy()
    got = G_fcn.run(zg, zl)
    assert got.shape == (1, 3, 256, 384)
    assert _nmax(got, want) <= TOL
    with pytest.raises(ValueError):
        G.run(zg, zl)          # wrong canvas for the scale-1 view (set_shape check)


def test_encoders_vs_oracle():
    rng = np.random.RandomState(7)
    for func in ('E_zl', 'E_zg'):
        params = R.init_params(func, rng, **R.CONFIG[func])
        x = rng.uniform(-1, 1, (5, 3, 128, 128)).astype(np.float32)
        with torch.no_grad():
            want = R.NETWORKS[func](torch.from_numpy(x), R.to_torch(params), **R.CONFIG[func])
        net = _make(func, params)
        got = net.run(x, return_as_list=True)
        for a, b in zip(got, want):
            assert a.shape == tuple(b.shape)
            assert _nmax(a, b.numpy()) <= TOL


def test_cfg1_reconstruction_path():
    """BASELINE cfg 1: one crop, E_zg + E_zl -> G_res(tile(zg_mu), z_mu) (run.py:371-375)."""
    rng = np.random.RandomState(1000)
    P = {f: R.init_params(f, rng, **R.CONFIG[f]) for f in ('E_zg', 'E_zl', 'G_res')}
    img = rng.uniform(-1, 1, (1, 3, 128, 128)).astype(np.float32)
    with torch.no_grad():
        x = torch.from_numpy(img)
        zg_mu, _ = R.E_zg(x, R.to_torch(P['E_zg']), **R.CONFIG['E_zg'])
        z_mu, _ = R.E_zl(x, R.to_torch(P['E_zl']), **R.CONFIG['E_zl'])
        want = R.G_res(zg_mu.repeat(1, 1, 32, 32), z_mu, R.to_torch(P['G_res']), **R.CONFIG['G_res']).numpy()
    nets = {f: _make(f, P[f]) for f in P}
    dev = torch.from_numpy(img).cuda()
    zg_mu_d, _ = nets['E_zg'].get_output_for(dev)
    z_mu_d, _ = nets['E_zl'].get_output_for(dev)
    from texturemixer_b200 import interp
    zg_t = interp.tiling_permutation(zg_mu_d, 32, 32, None, None, pin_corners=False)       # tf.tile of run.py:375
    got = nets['G_res'].get_output_for(zg_t, z_mu_d).cpu().numpy()
    assert _nmax(got, want) <= TOL


def test_cfg4_interpolation_canvas_small():
    """cfg 4 pattern at reduced batch: 4 sources -> 4x4 tile grid canvases -> G_res(scale 4x4) -> 512x512."""
    from texturemixer_b200 import interp
    from texturemixer_b200.network import Network
    rng = np.random.RandomState(1000)
    params = R.init_params('G_res', rng, **R.CONFIG['G_res'])
    G = _make('G_res', params)
    cfg = dict(R.CONFIG['G_res'], scale_h=4, scale_w=4)
    G_fcn = Network('G', func='networks.G_res', reuse=True, share_vars_with=G, num_channels=3, resolution=128, **cfg)
    n = 1
    zls = [rng.randn(n, 128, 32, 32).astype(np.float32) for _ in range(4)]
    zgs = [rng.randn(n, 128, 1, 1).astype(np.float32) for _ in range(4)]
    np.random.seed(1000)
    idx_h = [np.stack([I.sample_index_h(128, 5, np.random.uniform) for _ in range(n)]) for _ in range(4)]
    idx_w = [np.stack([I.sample_index_w(128, 5, np.random.uniform) for _ in range(n)]) for _ in range(4)]
    mattes = I.linkern_for_weight_arbitrary_shape(128, 128, 32)
    canv = [I.tiling_permutation_gather(s, ih, iw) for s, ih, iw in zip(zls, idx_h, idx_w)]
    want_zl = I.blend4(canv, mattes).astype(np.float32)
    want_zg = I.blend4([np.tile(z, (1, 1, 128, 128)) for z in zgs], mattes).astype(np.float32)
    zg, zl = interp.interpolate([torch.from_numpy(z).cuda() for z in zgs], [torch.from_numpy(z).cuda() for z in zls],
                                4, 4, idx_h=idx_h, idx_w=idx_w)
    assert np.array_equal(zl.cpu().numpy(), want_zl) and np.array_equal(zg.cpu().numpy(), want_zg)
    with torch.no_grad():
        want = R.G_res(torch.from_numpy(want_zg), torch.from_numpy(want_zl), R.to_torch(params), **cfg).numpy()
    got = G_fcn.get_output_for(zg, zl).cpu().numpy()
    assert got.shape == (n, 3, 512, 512)
    assert _nmax(got, want) <= TOL


def test_full_size_properties_cfg2():
    """BASELINE cfg 2 at full size (batch 64): size-independent properties instead of the oracle.
    (a) batch independence: samples computed inside the batch equal the same samples computed alone;
    (b) determinism: two runs are bit-identical; (c) range: tanh output in (-1, 1), finite."""
    rng = np.random.RandomState(1000)
    params = R.init_params('G_res', rng, **R.CONFIG['G_res'])
    net = _make('G_res', params)
    zg = np.tile(rng.randn(64, 128, 1, 1).astype(np.float32), (1, 1, 32, 32))
    zl = rng.randn(64, 128, 32, 32).astype(np.float32)
    a = net.run(zg, zl)
    b = net.run(zg, zl)
    assert a.shape == (64, 3, 128, 128) and np.array_equal(a, b)
    assert np.isfinite(a).all() and np.abs(a).max() <= 1.0
    sub = net.run(zg[[5, 63]], zl[[5, 63]])
    assert np.abs(sub - a[[5, 63]]).max() <= 1e-6
    # the oracle on two samples of the full batch
    with torch.no_grad():
        want = R.G_res(torch.from_numpy(zg[:2]), torch.from_numpy(zl[:2]), R.to_torch(params),
                       **R.CONFIG['G_res']).numpy()
    assert _nmax(a[:2], want) <= TOL


def test_network_run_output_conversion():
    """tfutil.py:649-659: out_mul/out_add -> round -> saturate-cast to uint8."""
    rng = np.random.RandomState(3)
    params = R.init_params('G_res', rng, **R.CONFIG['G_res'])
    net = _make('G_res', params)
    zg, zl = _inputs('G_res', rng, 2)
    f = net.run(zg, zl)
    u = net.run(zg, zl, out_mul=127.5, out_add=127.5, out_dtype=np.uint8, minibatch_size=1)
    want = np.clip(np.rint(f * np.float32(127.5) + np.float32(127.5)), 0, 255).astype(np.uint8)
    assert u.dtype == np.uint8 and np.abs(u.astype(np.int32) - want.astype(np.int32)).max() <= 1


def test_network_run_num_gpus():
    """tfutil.py:644-661: Network.run(num_gpus=k) splits every minibatch over k devices of the process; the result
    equals the single-device run bit for bit (same kernels, same per-image arithmetic).  With one visible device
    k = 2 must fail loudly instead of silently running on one."""
    rng = np.random.RandomState(4)
    params = R.init_params('G_res', rng, **R.CONFIG['G_res'])
    net = _make('G_res', params)
    zg, zl = _inputs('G_res', rng, 6)
    if torch.cuda.device_count() < 2:
        with pytest.raises(RuntimeError, match='CUDA device'):
            net.run(zg, zl, num_gpus=2)
        one = net.run(zg, zl, num_gpus=1, minibatch_size=4)
        assert one.shape == (6, 3, 128, 128)
        return
    one = net.run(zg, zl, minibatch_size=4)
    two = net.run(zg, zl, num_gpus=2, minibatch_size=4)              # minibatches of 4 + 2, each split 2 ways
    assert two.shape == one.shape and np.array_equal(one, two)
    net.set_var('64x64/Residual_0/weight', net.get_var('64x64/Residual_0/weight') * 1.5) if \
        '64x64/Residual_0/weight' in net.vars else None
    name = next(iter(net.trainables))
    net.set_var(name, net.get_var(name) * 1.25)                      # replicas must follow a weight change
    assert np.array_equal(net.run(zg, zl), net.run(zg, zl, num_gpus=2))
    u = net.run(zg, zl, num_gpus=2, out_mul=127.5, out_add=127.5, out_dtype=np.uint8)
    assert u.dtype == np.uint8 and np.array_equal(u, net.run(zg, zl, out_mul=127.5, out_add=127.5, out_dtype=np.uint8))
    # a minibatch-dependent network: the critic's minibatch-stddev groups are formed inside each device's part
    pd = R.init_params('D_patch', rng, **R.CONFIG['D_patch'])
    D = _make('D_patch', pd)
    x = rng.uniform(-1, 1, (16, 3, 128, 128)).astype(np.float32)
    got = D.run(x, num_gpus=2, minibatch_size=16)
    want = np.concatenate([D.run(x[:8]), D.run(x[8:])], axis=0)
    assert np.array_equal(got, want)


def _variant_table():
    src = open(os.path.join(GOLDEN, 'make_golden.py')).read()
    ns = {}
    exec(src[src.index('VARIANTS = ['):src.index('def gen_network_variants')], ns)
    return ns['VARIANTS'], ns['variant_tag']


VARIANTS, variant_tag = _variant_table()


@pytest.mark.parametrize('func,lod,pn', VARIANTS)
def test_lod_and_pixelnorm_variants_match_reference_golden(func, lod, pn):
    """every branch of the progressive-growing tf.cond trees (integer and fractional lod:
    FromRGB / ToRGB heads of the lower resolutions, image down/up-scaling, fades) and use_pixelnorm, against the
    outputs of the reference's own networks.py (tests/golden/networks_variants.npz) and the oracle."""
    g = np.load(os.path.join(GOLDEN, 'networks_variants.npz'))
    n = 4 if func == 'D_patch' else 2
    rng = np.random.RandomState(1000)
    cfg = dict(R.CONFIG[func])
    extra = {'use_pixelnorm': True} if pn else {}
    cfg.update(extra)
    params = R.init_params(func, rng, **cfg)
    params['lod'] = np.float32(lod)
    ins = _inputs(func, rng, n)
    net = _make(func, params, **extra)
    assert net.lod == float(np.float32(lod))
    outs = net.run(*ins, return_as_list=True)
    tag = variant_tag(func, lod, pn)
    with torch.no_grad():
        want = R.NETWORKS[func](*[torch.from_numpy(a) for a in ins], R.to_torch(params), **cfg)
    want = want if isinstance(want, tuple) else (want,)
    for i, a in enumerate(outs):
        assert list(a.shape) == g['%s_out%d_shape' % (tag, i)].tolist()
        flat = a.reshape(-1)
        got = flat[::SUBSAMPLE] if flat.size > 4096 else flat
        scale = float(g['%s_out%d_absmax' % (tag, i)][0])
        assert np.abs(got - g['%s_out%d' % (tag, i)]).max() <= TOL * scale
        assert _nmax(a, want[i].numpy()) <= TOL


def test_output_conversion_kernel_bit_exact():
    """tfutil.py:649-659 on the device (tmx_convert_output): x * mul + add, avg-pool shrink, round half to even,
    saturate - bit-exact against numpy for uint8, and the shrink path against an fp32 mean."""
    from texturemixer_b200.network import _convert_output
    rng = np.random.RandomState(4)
    x = (rng.randn(3, 3, 16, 24) * 1.2).astype(np.float32)
    x[0, 0, 0, :4] = [1.0, -1.0, 0.00392157, np.nan]            # 255, 0, a .5 tie region, NaN -> 0
    xd = torch.from_numpy(x).cuda()
    u = _convert_output(xd, 127.5, 127.5, 1, np.uint8).cpu().numpy()
    v = x * np.float32(127.5) + np.float32(127.5)
    want = np.where(np.isnan(v), 0, np.clip(np.rint(v), 0, 255)).astype(np.uint8)
    assert u.dtype == np.uint8 and np.array_equal(u, want)
    ties = torch.tensor([[[[0.5, 1.5, 2.5, 3.5, 254.5, 255.5, -0.5, 300.0]]]], dtype=torch.float32).cuda()
    assert _convert_output(ties, 1.0, 0.0, 1, np.uint8).cpu().numpy().reshape(-1).tolist() == [0, 2, 2, 4, 254, 255, 0, 255]
    s = _convert_output(xd[1:], 2.0, 0.25, 4, None).cpu().numpy()
    ref = (x[1:] * np.float32(2.0) + np.float32(0.25)).reshape(2, 3, 4, 4, 6, 4).mean(axis=(3, 5))
    assert s.shape == (2, 3, 4, 6) and np.abs(s - ref).max() <= 1e-5
    i16 = _convert_output(xd[1:], 30000.0, 0.0, 1, np.int16).cpu().numpy()
    assert i16.dtype == np.int16 and np.array_equal(i16, np.clip(np.rint(x[1:] * np.float32(30000.0)), -32768, 32767).astype(np.int16))


def _fused_table():
    src = open(os.path.join(GOLDEN, 'make_golden.py')).read()
    ns = {}
    exec(src[src.index('FUSED_VARIANTS = ['):src.index('def gen_network_fused')], ns)
    return ns['FUSED_VARIANTS']


@pytest.mark.parametrize('func,lod', _fused_table())
def test_fused_scale_variants_match_reference_golden(func, lod):
    """fused_scale=True - `upscale2d_conv2d` (conv2d_transpose, networks.py:94-101) as the sub-pixel
    upsample+conv kernel over ZERO-halo planes with the flipped / channel-swapped kernel, `conv2d_downscale2d`
    (:142-148) as zero-padded conv + 2x2 average + bias/activation - against the reference's own networks.py."""
    g = np.load(os.path.join(GOLDEN, 'networks_fused.npz'))
    n = 4 if func == 'D_patch' else 2
    rng = np.random.RandomState(1000)
    cfg = dict(R.CONFIG[func], fused_scale=True)
    params = R.init_params(func, rng, **cfg)
    params['lod'] = np.float32(lod)
    ins = _inputs(func, rng, n)
    net = _make(func, params, fused_scale=True)
    tag = variant_tag(func, lod, False) + '_fused'
    assert list(net.vars.keys()) == [str(s) for s in g[tag + '_varnames']]
    outs = net.run(*ins, return_as_list=True)
    with torch.no_grad():
        want = R.NETWORKS[func](*[torch.from_numpy(a) for a in ins], R.to_torch(params), **cfg)
    want = want if isinstance(want, tuple) else (want,)
    for i, a in enumerate(outs):
        assert list(a.shape) == g['%s_out%d_shape' % (tag, i)].tolist()
        flat = a.reshape(-1)
        got = flat[::SUBSAMPLE] if flat.size > 4096 else flat
        scale = float(g['%s_out%d_absmax' % (tag, i)][0])
        assert np.abs(got - g['%s_out%d' % (tag, i)]).max() <= TOL * scale
        assert _nmax(a, want[i].numpy()) <= TOL
