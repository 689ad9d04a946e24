"""GPU parity of the training-step building blocks that exist so far: D_patch
forward (minibatch stddev, dense head), fused Adam / non-finite guard / EMA."""
import numpy as np
import pytest
import torch

from oracle import networks_ref as R
from oracle import optim_ref as O

pytestmark = pytest.mark.gpu


def _nmax(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


@pytest.fixture(scope='module')
def rt():
    from texturemixer_b200.runtime import Runtime
    return Runtime.get(0)


@pytest.mark.parametrize('n,group', [(8, 4), (4, 4), (2, 4), (12, 4)])
def test_mbstd_vs_oracle(rt, n, group):
    from texturemixer_b200.runtime import Act
    rng = np.random.RandomState(n)
    x = rng.randn(n, 32, 4, 4).astype(np.float32)
    want = R.minibatch_stddev_layer(torch.from_numpy(x), group).numpy()
    a = Act(n, 4, 4, 32, f32=torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1))).cuda())
    out, stat = rt.mbstd(a, group)
    got = out.f32.cpu().numpy().transpose(0, 3, 1, 2)
    assert got.shape == (n, 64, 4, 4)
    assert np.array_equal(got[:, :32], x)
    assert _nmax(got[:, 32], want[:, 32]) <= 1e-5
    assert np.abs(got[:, 33:]).max() == 0.0


@pytest.mark.parametrize('n,k,cout,lrelu', [(32, 8192, 512, True), (32, 512, 1, False), (5, 300, 70, True)])
def test_dense_vs_oracle(rt, n, k, cout, lrelu):
    rng = np.random.RandomState(k)
    x = rng.randn(n, k).astype(np.float32)
    w = rng.randn(k, cout).astype(np.float32)
    b = (0.1 * rng.randn(cout)).astype(np.float32)
    y = R.apply_bias(R.dense(torch.from_numpy(x), torch.from_numpy(w)), torch.from_numpy(b))
    if lrelu:
        y = R.leaky_relu(y)
    got = rt.dense(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(),
                   float(R.wscale_of(w.shape)), lrelu)
    assert _nmax(got.cpu().numpy(), y.numpy()) <= 2e-5


@pytest.mark.parametrize('n', [8, 4])
def test_discriminator_vs_oracle_and_golden(n):
    import os
    from texturemixer_b200.network import Network
    rng = np.random.RandomState(1000)
    params = R.init_params('D_patch', rng, **R.CONFIG['D_patch'])
    x = rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)
    with torch.no_grad():
        want = R.D_patch(torch.from_numpy(x), R.to_torch(params), **R.CONFIG['D_patch']).numpy()
    D = Network('D_rec', func='networks.D_patch', seed=0, num_channels=3, resolution=128, **R.CONFIG['D_patch'])
    D.set_vars(params)
    got = D.run(x)
    assert got.shape == (n, 1, 1, 1)
    assert _nmax(got, want) <= 5e-4
    if n == 8:   # the golden minted by the reference's own networks.py uses the same seed/inputs
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'networks.npz'))
        assert np.abs(got.reshape(-1) - g['D_patch_out0']).max() <= 5e-4 * float(g['D_patch_out0_absmax'][0])


def test_adam_nonfinite_ema_vs_oracle(rt):
    from texturemixer_b200.network import Network
    from texturemixer_b200.optim import Optimizer
    cfg = dict(fmap_base=1024, fmap_max=512, latent_res=32, latent_channels=128, use_pixelnorm=False, tanh_at_end=False)
    E = Network('E_zl', func='networks.E_zl', seed=3, num_channels=3, resolution=128, **cfg)
    Es = E.clone('Es_zl')
    ema = Es.setup_as_moving_average_of(E, beta=0.999)
    n = E.flat.numel()
    w = E.flat.cpu().numpy().copy()
    ws = w.copy()
    state = O.AdamState(n, 0.0, 0.99)
    opt = Optimizer(name='TrainEG', learning_rate=0.0015, beta1=0.0, beta2=0.99, epsilon=1e-8)
    g_dev = torch.zeros_like(E.flat)
    opt.register_gradients(E, g_dev)
    rng = np.random.RandomState(0)
    applied = []
    for step in range(5):
        g = (rng.randn(n) * 10.0 ** rng.uniform(-4, 1)).astype(np.float32)
        g[:64] = 0                                         # slot of the non-trainable `lod`: no gradient
        if step == 2:
            g[64 + rng.randint(n - 64)] = np.inf           # overflow -> the whole step is skipped
        g_dev.copy_(torch.from_numpy(g))
        flag = opt.apply_updates()
        ema()
        ok = O.optimizer_step(w, [g], state, 0.0015, 0.0, 0.99, 1e-8)
        O.ema_update(w, ws, 0.999)
        applied.append(ok)
        assert int(flag.item()) == (0 if ok else 1)
        got = E.flat.cpu().numpy()
        assert np.abs(got - w).max() <= 2e-6 * max(1.0, np.abs(w).max()), step
        assert np.abs(Es.flat.cpu().numpy() - ws).max() <= 2e-6 * max(1.0, np.abs(ws).max())
    assert applied == [True, True, False, True, True]
    # the cached tensor-core weight planes were invalidated by the step
    assert E._owner()._version > 0


@pytest.mark.parametrize('lod,mirror', [(0.0, False), (1.0, False), (1.25, True), (2.0, False), (0.5, True)])
def test_process_reals_vs_numpy(rt, lod, mirror):
    """run.py:68-102: dynamic range, mirror augmentation, FadeLOD, UpscaleLOD restated in numpy."""
    from texturemixer_b200.train import process_reals
    rng = np.random.RandomState(int(lod * 100) + 3)
    r = 128 // 2 ** int(np.floor(lod))
    x = rng.randint(0, 256, (5, 3, r, r)).astype(np.uint8)
    fade, orig = process_reals(torch.from_numpy(x).cuda(), lod, lr_mirror_augment=mirror, ud_mirror_augment=mirror,
                               rng=np.random.RandomState(9))
    scale = np.float32(2.0) / np.float32(255.0)
    y = x.astype(np.float32) * scale + (np.float32(-1.0) - np.float32(0.0) * scale)
    if mirror:
        draw = np.random.RandomState(9)
        f = draw.uniform(0.0, 1.0, 5) >= 0.5
        y = np.where(f[:, None, None, None], y[:, :, :, ::-1], y)
        f = draw.uniform(0.0, 1.0, 5) >= 0.5
        y = np.where(f[:, None, None, None], y[:, :, ::-1, :], y)
    box = y.reshape(5, 3, r // 2, 2, r // 2, 2).mean(axis=(3, 5), keepdims=True)
    box = np.tile(box, (1, 1, 1, 2, 1, 2)).reshape(5, 3, r, r)
    t = np.float32(lod) - np.floor(np.float32(lod))
    yf = y + (box - y) * t
    k = 2 ** int(np.floor(lod))
    up = lambda a: np.repeat(np.repeat(a, k, axis=2), k, axis=3)
    assert orig.shape == (5, 3, 128, 128) and fade.shape == (5, 3, 128, 128)
    assert np.array_equal(orig.cpu().numpy(), up(y))
    assert np.abs(fade.cpu().numpy() - up(yf)).max() <= 1e-6
