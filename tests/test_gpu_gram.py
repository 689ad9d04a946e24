"""GPU parity of the VGG-19 Gram-matrix loss (custom_vgg19.py:20-66, loss.py:29-35, 68-75, 148-160,
206-213, 248-257) against

 (a) torch for the Gram kernels themselves (tmx_gram_fwd / tmx_gram_l1 / tmx_gram_bwd, tmx_vgg_preprocess),
 (b) oracle/vgg_ref.py (pinned to the reference's own custom_vgg19.py + loss.py by tests/test_loss_golden.py) for the
     feature extractor, the Gram term and its image gradient - the gradient once flip-limited (ReLU's derivative is
     discontinuous at 0) and once with the DEVICE's ReLU branch masks fed to the oracle (<= 1e-3, north_star),
 (c) tests/golden/losses_gram.npz - the reference's loss.py + custom_vgg19.py executed on the shim with the seeded
     stand-in weights (the real vgg19.npy is not redistributable): every Gram term of EG_wgan and the variable
     gradients of E_zg / E_zl / G with the Gram terms on."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import networks_ref as R
from oracle import vgg_ref as V

from loss_case import (GOLDEN, GRAM_WEIGHT, LOSS_FUNCS, LOSS_NETS, golden_gradient, gram_alpha, loss_case_inputs,
                       subsample, vgg_standin_weights)

pytestmark = pytest.mark.gpu

TOL_TERM = 1e-3
TOL_GRAD_MASKED = 1e-3
TOL_GRAD_FLIPS = 2e-2


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@pytest.fixture(scope='module')
def rt():
    from texturemixer_b200.runtime import Runtime
    return Runtime.get(0)


# ---------------------------------------------------------------------- (a) kernels vs torch
@pytest.mark.parametrize('n,c,h,w', [(2, 64, 16, 16), (3, 128, 8, 8), (2, 512, 8, 8), (2, 48, 6, 6), (1, 256, 32, 32)])
def test_gram_kernels_vs_torch(rt, n, c, h, w):
    from texturemixer_b200 import _lib
    g = torch.Generator().manual_seed(n * 1000 + c)
    F = torch.randn(n, c, h, w, generator=g).cuda()
    T = torch.randn(n, c, c, generator=g).cuda()
    G = rt.empty(n, c, c)
    _lib.check(rt.lib.tmx_gram_fwd(rt.handle, _p(F), _p(G), n, c, h, w, rt.stream()), 'tmx_gram_fwd')
    Fd = F.double().reshape(n, c, h * w).requires_grad_(True)
    Gd = torch.matmul(Fd, Fd.transpose(1, 2)) / h / w
    assert _rel(G.cpu().numpy(), Gd.detach().cpu().numpy()) <= 1e-5

    # two targets: plain and batch-reversed with a device-side weight (loss.py:252-254)
    wdev = torch.tensor([0.3], dtype=torch.float32).cuda()
    S = rt.empty(n, c, c)
    sums = torch.zeros(n, dtype=torch.float32).cuda()
    coef, vs = 0.002 / (n * c * c), 0.002 / (c * c)
    _lib.check(rt.lib.tmx_gram_l1(rt.handle, _p(G), _p(T), _p(S), _p(sums), n, c, 1, coef, vs, 0, _p(wdev), 2,
                                  rt.stream()), 'tmx_gram_l1')
    _lib.check(rt.lib.tmx_gram_l1(rt.handle, _p(G), _p(T), _p(S), _p(sums), n, c, 0, coef, vs, 1, _p(wdev), 1,
                                  rt.stream()), 'tmx_gram_l1')
    Td = T.double()
    per = 0.7 * (Gd - Td.flip(0)).abs().mean(dim=(1, 2)) + 0.3 * (Gd - Td).abs().mean(dim=(1, 2))
    loss = (per * 0.002).mean()
    loss.backward()
    assert np.allclose(sums.cpu().numpy(), (per * 0.002).detach().cpu().numpy(), rtol=2e-5)

    dF = rt.empty(n, c, h, w)
    _lib.check(rt.lib.tmx_gram_bwd(rt.handle, _p(S), _p(F), _p(dF), n, c, h, w, rt.stream()), 'tmx_gram_bwd')
    torch.cuda.synchronize()
    assert _rel(dF.cpu().numpy().reshape(n, c, h * w), Fd.grad.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize('n,c,h,w', [(3, 64, 16, 16), (2, 128, 8, 8), (2, 256, 32, 32), (3, 512, 16, 16),
                                     (2, 64, 128, 128), (2, 512, 8, 8)])
def test_gram_tensor_core_kernels_vs_torch(rt, n, c, h, w):
    """tmx_gram_fwd_tc (weight-gradient kernel, sample = tap) and the per-sample-weight 1x1 conv of the gradient
    (TMX_CONV_W_PER_SAMPLE) against fp64 torch, from split-bf16 planes with a zero halo."""
    import types
    from texturemixer_b200.runtime import Act
    from texturemixer_b200.vgg import GramLoss
    g = torch.Generator().manual_seed(n * 77 + c + h)
    F = torch.relu(torch.randn(n, h, w, c, generator=g)).cuda()          # NHWC, like a VGG activation
    S = torch.randn(n, c, c, generator=g).cuda() * 1e-3
    act = rt.split_pack(Act(n, h, w, c, f32=F.clone()), 'zero')
    me = types.SimpleNamespace(rt=rt, use_tc=True)
    G = GramLoss._gram_of(me, act)
    dF = GramLoss._feature_gradient(me, act, S)
    torch.cuda.synchronize()
    Fd = F.double().reshape(n, h * w, c)
    Gd = torch.matmul(Fd.transpose(1, 2), Fd) / h / w
    assert _rel(G.cpu().numpy(), Gd.cpu().numpy()) <= 2e-5
    Sd = S.double()
    want = torch.matmul(Fd, (Sd + Sd.transpose(1, 2)).transpose(1, 2)) / h / w          # [n, p, i]
    assert _rel(dF.reshape(n, h * w, c).cpu().numpy(), want.cpu().numpy()) <= 2e-5


def test_vgg_preprocess_and_adjoint(rt):
    from texturemixer_b200 import _lib
    n, h, w = 2, 12, 20
    g = torch.Generator().manual_seed(5)
    img = (torch.rand(n, 3, h, w, generator=g) * 2 - 1).cuda()
    out = rt.empty(n, h, w, 16)
    _lib.check(rt.lib.tmx_vgg_preprocess(rt.handle, _p(img), _p(out), n, h, w, rt.stream()), 'tmx_vgg_preprocess')
    x = (img + 1.0) / 2.0 * 255.0
    want = torch.stack([x[:, 2] - V.VGG_MEAN[0], x[:, 1] - V.VGG_MEAN[1], x[:, 0] - V.VGG_MEAN[2]], dim=-1)
    assert torch.allclose(out[..., :3], want, rtol=0, atol=2e-5)   # values up to 255: 1 ulp = 1.5e-5 (fma contraction)
    assert float(out[..., 3:].abs().max()) == 0.0
    d = torch.randn(n, h, w, 16, generator=g).cuda()
    dimg = rt.empty(n, 3, h, w)
    _lib.check(rt.lib.tmx_vgg_preprocess_bwd(rt.handle, _p(d), _p(dimg), n, h, w, rt.stream()),
               'tmx_vgg_preprocess_bwd')
    want = torch.stack([d[..., 2], d[..., 1], d[..., 0]], dim=1) * 127.5
    assert torch.allclose(dimg, want, rtol=1e-6, atol=0)


# ---------------------------------------------------------------------- (b) feature extractor + term vs the oracle
@pytest.fixture(scope='module', params=['tc', 'ffma'])
def gram(request):
    """'tc': Gram matrices and their gradient GEMMs on the tensor cores (the default); 'ffma': the fp32 CUDA-core
    kernels on NCHW copies (what odd shapes fall back to)."""
    from texturemixer_b200.vgg import GramLoss
    g = GramLoss(vgg_standin_weights(), resolution=128, device=0)
    g.use_tc = request.param == 'tc'
    return g


def _images(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, 128, 128, generator=g) * 2 - 1


def test_vgg_features_and_grams_vs_oracle(gram):
    img = _images(2, 11)
    acts, gs = gram.grams(img.cuda())
    rt = gram.rt
    feats = [rt.nhwc_to_nchw(rt.split_unpack(a).f32) for a in acts]
    torch.cuda.synchronize()
    with torch.no_grad():
        want = V.vgg_features(img, vgg_standin_weights())
    assert [tuple(f.shape) for f in feats] == [tuple(want[k].shape) for k in V.GRAM_LAYERS]
    errs = {}
    for f, g, k in zip(feats, gs, V.GRAM_LAYERS):
        ref = want[k].numpy()
        errs[k] = (float(np.abs(f.cpu().numpy() - ref).max() / np.abs(ref).max()),
                   _rel(g.cpu().numpy(), V.gram_matrix(want[k]).numpy()))
    print('VGG-19 features (normalised max error) / Gram matrices (rel-L2) vs oracle:', errs)
    assert all(a <= 1e-3 and b <= 1e-3 for a, b in errs.values()), errs


def _oracle_term(img, real, weights, reverse=False, alpha_bar=None):
    x = img.clone().requires_grad_(True)
    fake_g = V.grams(x, weights)
    with torch.no_grad():
        real_g = V.grams(real, weights)
    if alpha_bar is None:
        per = V.multi_layer_diff(fake_g, real_g) * GRAM_WEIGHT
    else:
        rev = [torch.flip(m, dims=[0]) for m in real_g]
        per = ((1 - alpha_bar) * V.multi_layer_diff(fake_g, rev) + alpha_bar * V.multi_layer_diff(fake_g, real_g)) \
            * GRAM_WEIGHT
    per.mean().backward()
    return float(per.mean()), x.grad.numpy()


def test_gram_term_vs_oracle_autograd(gram):
    """value within 1e-3; image gradient flip-limited (a ReLU within forward rounding of 0 takes the other branch)."""
    n = 4
    fake, real = _images(n, 21), _images(n, 22)
    weights = vgg_standin_weights()
    _, real_gram = gram.grams(real.cuda())
    value, dimg = gram.term(fake.cuda(), [(real_gram, False, None, 0)], GRAM_WEIGHT)
    torch.cuda.synchronize()
    want, dwant = _oracle_term(fake, real, weights)
    assert abs(float(value) - want) <= TOL_TERM * abs(want), (float(value), want)
    err = _rel(dimg.cpu().numpy(), dwant)
    print('Gram term image gradient vs autograd (flip-limited): rel-L2', err)
    assert err <= TOL_GRAD_FLIPS, err

    # the two-target form of loss.py:252-254 with the batch-mean weight on the device
    abar = torch.tensor([0.37], dtype=torch.float32).cuda()
    value2, dimg2 = gram.term(fake.cuda(), [(real_gram, True, abar, 2), (real_gram, False, abar, 1)], GRAM_WEIGHT)
    torch.cuda.synchronize()
    want2, dwant2 = _oracle_term(fake, real, weights, alpha_bar=0.37)
    assert abs(float(value2) - want2) <= TOL_TERM * abs(want2), (float(value2), want2)
    assert _rel(dimg2.cpu().numpy(), dwant2) <= TOL_GRAD_FLIPS


class _ReluMaskFeed:
    def __init__(self, masks):
        self.q, self.flips, self.total = list(masks), 0, 0

    def __call__(self, x):
        m = self.q.pop(0)
        assert tuple(m.shape) == tuple(x.shape), (tuple(m.shape), tuple(x.shape))
        self.flips += int(((x.detach() > 0) != m).sum())
        self.total += m.numel()
        return x * m.to(x.dtype)


def test_gram_term_gradient_with_shared_relu_masks(gram, monkeypatch):
    """Both sides differentiate the same piecewise-linear function: <= 1e-3 relative L2 (north_star)."""
    from texturemixer_b200.backward import backward
    from texturemixer_b200 import _lib
    n = 2
    fake, real = _images(n, 31), _images(n, 32)
    weights = vgg_standin_weights()
    rt = gram.rt
    _, real_gram = gram.grams(real.cuda())
    tape = []
    gram.grams(fake.cuda(), tape=tape)
    masks = []
    for rec in tape:
        if rec['kind'] == 'conv':
            masks.append((rt.split_unpack(rec['y']).f32 > 0).permute(0, 3, 1, 2).cpu())
    assert len(masks) == 13                                     # conv1_1 .. conv5_1
    value, dimg = gram.term(fake.cuda(), [(real_gram, False, None, 0)], GRAM_WEIGHT)
    torch.cuda.synchronize()
    feed = _ReluMaskFeed(masks)
    x = fake.clone().requires_grad_(True)
    monkeypatch.setattr(V, 'relu', feed)
    fake_g = V.grams(x, weights)
    assert not feed.q
    monkeypatch.undo()
    with torch.no_grad():
        real_g = V.grams(real, weights)
    per = V.multi_layer_diff(fake_g, real_g) * GRAM_WEIGHT
    per.mean().backward()
    assert abs(float(value) - float(per.mean())) <= TOL_TERM * abs(float(per.mean()))
    err = _rel(dimg.cpu().numpy(), x.grad.numpy())
    print('Gram term image gradient, shared ReLU masks (%d of %d branches differ): rel-L2 %.3g'
          % (feed.flips, feed.total, err))
    assert err <= TOL_GRAD_MASKED, err


# ---------------------------------------------------------------------- (c) EG_wgan with the Gram terms vs the fixture
def test_eg_wgan_with_gram_vs_reference_golden(gram):
    from texturemixer_b200 import loss as dev_loss
    from texturemixer_b200.network import Network
    g = np.load(os.path.join(os.path.dirname(GOLDEN), 'losses_gram.npz'))
    g0 = np.load(GOLDEN)
    n, sh, sw, _ = (int(v) for v in g0['meta_n_sh_sw_stride'])
    params, reals, idx, crops, mixes = loss_case_inputs(n, sh, sw)
    nets = {}
    for k in LOSS_NETS:
        f = LOSS_FUNCS[k]
        nets[k] = Network(k, func='networks.' + f, seed=0, num_channels=3, resolution=128, **R.CONFIG[f])
        nets[k].set_vars(params[k])
    nets['G_fcn'] = Network('G', func='networks.G_res', reuse=True, share_vars_with=nets['G'], num_channels=3,
                            resolution=128, scale_h=sh, scale_w=sw, **R.CONFIG['G_res'])
    grads = {k: torch.zeros_like(nets[k].flat) for k in ('E_zg', 'E_zl', 'G')}
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rep = dev_loss.EG_wgan(nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], nets['G_fcn'], nets['D_interp'],
                           nets['D_blend'], dev(reals), idx, crops['eg_crop_interp'], crops['eg_crop_blend'],
                           dev(mixes['eg_mix']), grads, scale_h=sh, scale_w=sw, crop_aware=True, gram=gram,
                           gram_weight=GRAM_WEIGHT, gram_alpha=dev(gram_alpha(n)))
    torch.cuda.synchronize()
    rep = {k: float(v.reshape(-1)[0]) for k, v in rep.items()}
    for mine, ref in (('rec_gram', 'rec_gram_loss'), ('interp_gram', 'crop_interp_gram_loss'),
                      ('blend_gram', 'crop_blend_interp_gram_loss'), ('rec_G', 'rec_G_loss')):
        want = float(g['EGgram_term_Loss_' + ref].mean())
        assert abs(rep[mine] - want) <= TOL_TERM * max(abs(want), 1e-2), (mine, rep[mine], want)
    total = sum(rep[k] for k in ('rec_G', 'rec_pixel', 'interp_G', 'blend_G', 'rec_gram', 'interp_gram', 'blend_gram'))
    want = float(g['EGgram_loss'].mean())
    assert abs(total - want) <= TOL_TERM * abs(want), (total, want)
    worst = ('', 0.0)
    for scope in ('E_zg', 'E_zl', 'G'):
        for name in nets[scope].trainables:
            wantg, norm = golden_gradient(g, 'EGgram', scope, name)
            got = nets[scope].grad_view(grads[scope], name).cpu().numpy()
            if norm == 0.0:
                assert np.abs(got).max() == 0.0, (scope, name)
                continue
            err = _rel(subsample(got), wantg)
            worst = max(worst, (scope + '/' + name, err), key=lambda p: p[1])
            assert err <= TOL_GRAD_FLIPS, (scope, name, err)
    print('EG_wgan + Gram vs reference fixture: worst variable gradient rel-L2 (flip-limited)', worst)
