"""GPU parity of the E/G phase of the train step (loss.EG_wgan, loss.py:105-259, gram off): loss terms and every
variable gradient of E_zg, E_zl and G against the oracle's autograd, same crop offsets / mixing factors / index
vectors on both sides.  alpha = 1 checks the whole chain exactly; alpha = 0.2 is leaky-ReLU-flip limited (see
tests/test_gpu_backward.py)."""
import numpy as np
import pytest
import torch

from oracle import interp_ref as I
from oracle import loss_ref as L
from oracle import networks_ref as R

pytestmark = pytest.mark.gpu


def _rel_l2(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


@pytest.mark.parametrize('alpha,sh,sw,kl', [(1.0, 2, 2, 0.0), (0.2, 3, 3, 0.0), (1.0, 2, 2, 1.0)])
def test_eg_loss_and_gradients_vs_autograd(alpha, sh, sw, kl, monkeypatch):
    from texturemixer_b200 import loss as dev_loss
    from texturemixer_b200 import runtime
    from texturemixer_b200.network import Network
    monkeypatch.setattr(runtime, 'LRELU_ALPHA', alpha)
    monkeypatch.setattr(R, 'leaky_relu', lambda x, a=alpha: torch.maximum(x * a, x) if a != 1.0 else x)
    rng = np.random.RandomState(1000)
    n = 4
    names = ['E_zg', 'E_zl', 'G', 'D_rec', 'D_interp', 'D_blend']
    funcs = dict(E_zg='E_zg', E_zl='E_zl', G='G_res', D_rec='D_patch', D_interp='D_patch', D_blend='D_patch')
    params = {k: R.init_params(funcs[k], rng, **R.CONFIG[funcs[k]]) for k in names}
    if alpha == 1.0:
        for k in names:          # without the damping of the leaky ReLU the random nets blow up: tame the weights
            for v in params[k]:
                if v.endswith('weight'):
                    params[k][v] = params[k][v] * np.float32(0.5)
    reals = rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)
    np.random.seed(1000)
    idx = I.sample_schedule_indices(n, latent_res=32, scale_h=sh, scale_w=sw)
    crop_i = (int(rng.randint(0, 128 * sh - 128 + 1)), int(rng.randint(0, 128 * sw - 128 + 1)))
    crop_b = (int(rng.randint(0, 128 * sh - 128 + 1)), int(rng.randint(0, 128 * sw - 128 + 1)))
    mix = rng.uniform(0, 1, (n, 1, 1, 1)).astype(np.float32)

    # ---- oracle
    torch.set_num_threads(max(1, torch.get_num_threads()))
    P = {k: R.to_torch(params[k], requires_grad=k in ('E_zg', 'E_zl', 'G')) for k in names}
    cfg = dict(R.CONFIG)
    loss, terms = L.EG_wgan(P, torch.from_numpy(reals), idx, crop_i, crop_b, torch.from_numpy(mix), scale_h=sh,
                            scale_w=sw, cfg=cfg, kl_weight=kl)
    loss.mean().backward()

    # ---- device
    nets = {}
    for k in names:
        nets[k] = Network(k, func='networks.' + funcs[k], seed=0, num_channels=3, resolution=128, **R.CONFIG[funcs[k]])
        nets[k].set_vars(params[k])
    G_fcn = Network('G', func='networks.G_res', reuse=True, share_vars_with=nets['G'], num_channels=3, resolution=128,
                    scale_h=sh, scale_w=sw, **R.CONFIG['G_res'])
    grads = {k: torch.zeros_like(nets[k].flat) for k in ('E_zg', 'E_zl', 'G')}
    rep = dev_loss.EG_wgan(nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], G_fcn, nets['D_interp'],
                           nets['D_blend'], torch.from_numpy(reals).cuda(), idx, crop_i, crop_b,
                           torch.from_numpy(mix).cuda(), grads, scale_h=sh, scale_w=sw, kl_weight=kl)
    torch.cuda.synchronize()
    ltol = 2e-3 if alpha == 1.0 else 1e-2
    keys = [('rec_G', 'rec_G'), ('rec_pixel', 'rec_pixel'), ('interp_G', 'interp_G'), ('blend_G', 'blend_G')]
    if kl > 0:                                                   # KL regulariser on both encoders (loss.py:163-171)
        keys += [('KL_zg', 'KL_zg'), ('KL_zl', 'KL_zl')]
    for k, key in keys:
        want = float(terms[key].mean())
        got = float(rep[k].reshape(-1)[0])
        assert abs(got - want) <= ltol * max(1.0, abs(want)), (k, got, want)
    gtol = 2e-3 if alpha == 1.0 else 3e-2
    worst = ('', 0.0)
    for k in ('E_zg', 'E_zl', 'G'):
        for name, t in P[k].items():
            if name == 'lod' or t.grad is None:
                continue
            want = t.grad.numpy()
            if np.abs(want).max() == 0:
                continue
            got = nets[k].grad_view(grads[k], name).cpu().numpy()
            err = _rel_l2(got, want)
            if err > worst[1]:
                worst = (k + '/' + name, err)
            assert err <= gtol, (k, name, err)
    print('alpha', alpha, 'worst variable gradient rel-L2', worst)


@pytest.mark.parametrize('alpha,fused', [(1.0, False), (0.2, False), (1.0, True)])
def test_critic_wgangp_double_backward_vs_autograd(alpha, fused, monkeypatch):
    """D_*_wgangp (loss.py:303-521): every variable gradient of the critic, including the gradient penalty's
    second-order terms (tangent x adjoint weight gradients + minibatch-stddev curvature), vs create_graph autograd."""
    from texturemixer_b200 import loss as dev_loss
    from texturemixer_b200 import runtime
    from texturemixer_b200.network import Network
    monkeypatch.setattr(runtime, 'LRELU_ALPHA', alpha)
    monkeypatch.setattr(R, 'leaky_relu', lambda x, a=alpha: torch.maximum(x * a, x) if a != 1.0 else x)
    rng = np.random.RandomState(5)
    n = 8
    dcfg = dict(R.CONFIG['D_patch'], fused_scale=fused)      # fused: conv2d_downscale2d critics (networks.py:142-148)
    params = R.init_params('D_patch', rng, **dcfg)
    reals = rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)
    fakes = np.tanh(rng.randn(n, 3, 128, 128)).astype(np.float32)
    mix = rng.uniform(0, 1, (n, 1, 1, 1)).astype(np.float32)
    P = R.to_torch(params, dtype=torch.float64, requires_grad=True)
    loss, terms = L.D_wgangp(P, torch.from_numpy(fakes).double(), torch.from_numpy(reals).double(),
                             torch.from_numpy(mix).double(), cfg=dict(R.CONFIG, D_patch=dcfg))
    loss.mean().backward()
    D = Network('D_rec', func='networks.D_patch', seed=0, num_channels=3, resolution=128, **dcfg)
    D.set_vars(params)
    fg = torch.zeros_like(D.flat)
    rep = dev_loss.D_wgangp(D, torch.from_numpy(fakes).cuda(), torch.from_numpy(reals).cuda(),
                            torch.from_numpy(mix).cuda(), fg)
    torch.cuda.synchronize()
    ltol = 2e-3 if alpha == 1.0 else 2e-2
    for k in ('D_loss', 'gradient_penalty', 'epsilon_penalty'):
        want, got = float(terms[k].mean()), float(rep[k].reshape(-1)[0])
        assert abs(got - want) <= ltol * max(1e-3, abs(want)), (k, got, want)
    gtol = (4e-3 if fused else 3e-3) if alpha == 1.0 else 5e-2     # fused: bias gradients pass through the 2x2 average
    worst = ('', 0.0)
    for name, t in P.items():
        if name == 'lod' or t.grad is None or float(t.grad.abs().max()) == 0:
            continue
        err = _rel_l2(D.grad_view(fg, name).cpu().numpy(), t.grad.numpy())
        if err > worst[1]:
            worst = (name, err)
        assert err <= gtol, (name, err)
    print('alpha', alpha, 'critic worst variable gradient rel-L2', worst)


@pytest.mark.parametrize('crops', [((0, 0), (255, 255)), ((57, 131), (200, 56)), ((130, 3), (128, 129))])
def test_crop_aware_g_fcn_equals_whole_canvas(crops):
    """Decoding only the latent window each random_crop depends on (loss.crop_window; SURVEY Appendix C note) against
    decoding the whole 3x3 canvas, both on the device: same crops, same loss terms, same variable gradients up to the
    summation order of the weight-gradient split-K (the per-pixel arithmetic is identical)."""
    from texturemixer_b200 import loss as dev_loss
    from texturemixer_b200.network import Network
    rng = np.random.RandomState(11)
    n, sh, sw = 4, 3, 3
    names = ['E_zg', 'E_zl', 'G', 'D_rec', 'D_interp', 'D_blend']
    funcs = dict(E_zg='E_zg', E_zl='E_zl', G='G_res', D_rec='D_patch', D_interp='D_patch', D_blend='D_patch')
    nets = {k: Network(k, func='networks.' + funcs[k], seed=20 + i, num_channels=3, resolution=128,
                       **R.CONFIG[funcs[k]]) for i, k in enumerate(names)}
    for net in nets.values():
        for vn, v in net.trainables.items():
            if vn.endswith('/bias'):
                net.set_var(vn, 0.1 * rng.randn(*v.shape).astype(np.float32))
    G_fcn = Network('G', func='networks.G_res', reuse=True, share_vars_with=nets['G'], num_channels=3, resolution=128,
                    scale_h=sh, scale_w=sw, **R.CONFIG['G_res'])
    reals = torch.from_numpy(rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)).cuda()
    np.random.seed(5)
    idx = I.sample_schedule_indices(n, latent_res=32, scale_h=sh, scale_w=sw)
    mix = torch.from_numpy(rng.uniform(0, 1, (n, 1, 1, 1)).astype(np.float32)).cuda()
    out = {}
    for aware in (False, True):
        fwd = dev_loss.EGForward(nets['E_zg'], nets['E_zl'], nets['G'], G_fcn, reals, idx, mix, sh, sw,
                                 crop_interp=crops[0] if aware else None, crop_blend=crops[1] if aware else None)
        assert (fwd.win['interp'] is not None) == aware
        assert tuple(fwd.interp.shape[2:]) == ((160, 160) if aware else (384, 384))     # 40 latent pixels of tail
        imgs = (fwd.crop('interp', crops[0]).clone(), fwd.crop('blend', crops[1]).clone())
        grads = {k: torch.zeros_like(nets[k].flat) for k in ('E_zg', 'E_zl', 'G')}
        rep = dev_loss.EG_backward(fwd, nets['D_rec'], nets['D_interp'], nets['D_blend'], crops[0], crops[1], grads)
        torch.cuda.synchronize()
        out[aware] = (imgs, grads, {k: float(v.reshape(-1)[0]) for k, v in rep.items()})
    for a, b in zip(out[False][0], out[True][0]):
        assert float((a - b).abs().max()) <= 1e-6              # same per-pixel arithmetic (expected: bit-identical)
        print('crop pixels bit-identical:', bool(torch.equal(a, b)))
    for k, v in out[False][2].items():
        assert abs(out[True][2][k] - v) <= 1e-5 * max(1.0, abs(v)), k
    for k in ('E_zg', 'E_zl', 'G'):
        a, b = out[False][1][k].double(), out[True][1][k].double()
        assert float((a - b).norm() / a.norm()) <= 1e-5, k
    with pytest.raises(ValueError):
        fwd.crop('interp', (crops[0][0] + 1, crops[0][1]))      # a window serves exactly the crop it was planned for
