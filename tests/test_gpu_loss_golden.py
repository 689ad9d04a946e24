"""GPU parity of the device losses (texturemixer_b200/loss.py) against

 (a) tests/golden/losses.npz - the reference's own loss.py (/root/reference/loss.py:105-259, 303-521) executed
     unmodified on oracle/tfshim (tests/golden/make_golden.py gen_losses): every loss term, and every variable
     gradient (strided subsample + norm) of EG_wgan and the three D_*_wgangp; the product runs as it ships
     (crop-aware G_fcn).  Gradients of a leaky-ReLU network are discontinuous in the pre-activations, so against a
     fixture this comparison is limited by the few pre-activations that lie within forward rounding of zero and
     take the other branch on the device ("flips"): the tolerance is the flip-limited one.
 (b) the oracle (oracle/loss_ref.py, pinned bit-level to the same fixture by tests/test_loss_golden.py) evaluated
     with the DEVICE's leaky-ReLU branch masks, so that both sides differentiate the same piecewise-linear
     function: every variable gradient within 1e-3 relative L2 (BASELINE north_star tolerance)."""
import numpy as np
import pytest
import torch

from oracle import loss_ref as L
from oracle import networks_ref as R

from loss_case import GOLDEN, LOSS_FUNCS, LOSS_NETS, MODE_CASES, loss_case_inputs, golden_gradient, mode_noise, subsample

pytestmark = pytest.mark.gpu

TOL_TERM = 1e-3          # loss terms, relative (north_star)
TOL_GRAD_MASKED = 1e-3   # variable gradients with shared branch masks, relative L2 (north_star)
TOL_GRAD_FLIPS = 2e-2    # variable gradients against the fixture: leaky-ReLU-flip limited (see module docstring)


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


@pytest.fixture(scope='module')
def case():
    from texturemixer_b200.network import Network
    g = np.load(GOLDEN)
    n, sh, sw, _ = (int(v) for v in g['meta_n_sh_sw_stride'])
    params, reals, idx, crops, mixes = loss_case_inputs(n, sh, sw)
    nets = {}
    for k in LOSS_NETS:
        f = LOSS_FUNCS[k]
        nets[k] = Network(k, func='networks.' + f, seed=0, num_channels=3, resolution=128, **R.CONFIG[f])
        nets[k].set_vars(params[k])
    nets['G_fcn'] = Network('G', func='networks.G_res', reuse=True, share_vars_with=nets['G'], num_channels=3,
                            resolution=128, scale_h=sh, scale_w=sw, **R.CONFIG['G_res'])
    return dict(g=g, params=params, reals=reals, idx=idx, crops=crops, mixes=mixes, nets=nets, sh=sh, sw=sw, n=n)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---------------------------------------------------------------------- device evaluations
def _device_eg(c, crop_aware, sink=None):
    from texturemixer_b200 import loss as dev_loss
    nets = c['nets']
    grads = {k: torch.zeros_like(nets[k].flat) for k in ('E_zg', 'E_zl', 'G')}
    dev_loss.TAPE_SINK = sink
    try:
        rep = dev_loss.EG_wgan(nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], nets['G_fcn'], nets['D_interp'],
                               nets['D_blend'], _dev(c['reals']), c['idx'], c['crops']['eg_crop_interp'],
                               c['crops']['eg_crop_blend'], _dev(c['mixes']['eg_mix']), grads, scale_h=c['sh'],
                               scale_w=c['sw'], crop_aware=crop_aware)
        torch.cuda.synchronize()
    finally:
        dev_loss.TAPE_SINK = None
    return {k: float(v.reshape(-1)[0]) for k, v in rep.items()}, grads


def _device_critic(c, which, crop_aware, sink=None):
    """D_rec / D_interp / D_blend_wgangp as the trainer evaluates them: fakes from the E/G forward (no tape), then
    loss.D_wgangp."""
    from texturemixer_b200 import loss as dev_loss
    nets = c['nets']
    x = _dev(c['reals'])
    fwd = dev_loss.EGForward(nets['E_zg'], nets['E_zl'], nets['G'], nets['G_fcn'], x, c['idx'],
                             _dev(c['mixes']['eg_mix']), c['sh'], c['sw'], defer_canvases=True)
    if which == 'D_rec':
        fake, gp = fwd.rec, 'd_rec_gp'
    elif which == 'D_interp':
        fake, gp = dev_loss.fcn_fake(nets['G_fcn'], fwd, 'interp', c['crops']['d_interp_crop'],
                                     crop_aware=crop_aware), 'd_interp_gp'
    else:
        fake, gp = dev_loss.fcn_fake(nets['G_fcn'], fwd, 'blend', c['crops']['d_blend_crop'],
                                     _dev(c['mixes']['d_blend_mix']), crop_aware=crop_aware), 'd_blend_gp'
    fg = torch.zeros_like(nets[which].flat)
    dev_loss.TAPE_SINK = sink
    try:
        rep = dev_loss.D_wgangp(nets[which], fake, x, _dev(c['mixes'][gp]), fg)
        torch.cuda.synchronize()
    finally:
        dev_loss.TAPE_SINK = None
    return {k: float(v.reshape(-1)[0]) for k, v in rep.items()}, fg


# ---------------------------------------------------------------------- (a) against the reference-minted fixture
def _grads_vs_golden(c, tag, scopes, flat_grads, tol):
    g, nets = c['g'], c['nets']
    worst = ('', 0.0)
    for scope in scopes:
        for name in nets[scope].trainables:
            want, norm = golden_gradient(g, tag, scope, name)
            got = nets[scope].grad_view(flat_grads[scope], name).cpu().numpy()
            if norm == 0.0:
                assert np.abs(got).max() == 0.0, (scope, name)          # unused lod heads
                continue
            assert abs(np.linalg.norm(got.astype(np.float64)) - norm) <= tol * norm, (scope, name)
            err = _rel(subsample(got), want)
            worst = max(worst, (scope + '/' + name, err), key=lambda p: p[1])
            assert err <= tol, (tag, scope, name, err)
    return worst


def test_eg_wgan_vs_reference_golden(case):
    c = case
    rep, grads = _device_eg(c, crop_aware=True)
    for mine, ref in (('rec_G', 'rec_G_loss'), ('rec_pixel', 'rec_pixel_loss'), ('interp_G', 'crop_interp_G_loss'),
                      ('blend_G', 'crop_blend_interp_G_loss')):
        want = float(c['g']['EG_term_Loss_' + ref].mean())
        assert abs(rep[mine] - want) <= TOL_TERM * max(1.0, abs(want)), (mine, rep[mine], want)
    total = sum(rep[k] for k in ('rec_G', 'rec_pixel', 'interp_G', 'blend_G'))
    assert abs(total - float(c['g']['EG_loss'].mean())) <= TOL_TERM * abs(float(c['g']['EG_loss'].mean()))
    worst = _grads_vs_golden(c, 'EG', ('E_zg', 'E_zl', 'G'), grads, TOL_GRAD_FLIPS)
    print('EG_wgan vs reference fixture: worst variable gradient rel-L2 (flip-limited)', worst)


CRITIC_TERMS = dict(
    D_rec=dict(D_loss='rec_D_loss', gradient_penalty='rec_gradient_penalty', epsilon_penalty='rec_epsilon_penalty'),
    D_interp=dict(D_loss='crop_interp_D_loss', gradient_penalty='crop_interp_gradient_penalty',
                  epsilon_penalty='crop_interp_epsilon_penalty'),
    D_blend=dict(D_loss='crop_blend_interp_D_loss', gradient_penalty='crop_blend_interp_gradient_penalty',
                 epsilon_penalty='crop_blend_interp_epsilon_penalty'))


@pytest.mark.parametrize('which', ['D_rec', 'D_interp', 'D_blend'])
def test_critic_loss_vs_reference_golden(case, which):
    c = case
    rep, fg = _device_critic(c, which, crop_aware=True)
    for mine, ref in CRITIC_TERMS[which].items():
        want = float(c['g']['%s_term_Loss_%s' % (which, ref)].mean())
        assert abs(rep[mine] - want) <= TOL_TERM * max(abs(want), 1e-2), (which, mine, rep[mine], want)
    worst = _grads_vs_golden(c, which, (which,), {which: fg}, TOL_GRAD_FLIPS)
    print(which, 'vs reference fixture: worst variable gradient rel-L2 (flip-limited)', worst)


# ---------------------------------------------------------------------- (b) shared leaky-ReLU branch masks
def _masks_of(net, tape):
    """Branch masks (NCHW / [N,C] bool, CPU) of every activated layer of one recorded device evaluation, in
    evaluation order == the order in which the oracle calls leaky_relu for the same network."""
    rt = net.rt
    out = []
    for rec in tape:
        kind = rec['kind']
        if kind == 'fromrgb':
            out.append((rec['y'].f32 > 0).permute(0, 3, 1, 2).cpu())
        elif kind == 'conv' and rec['act']:
            out.append((rt.split_unpack(rec['y']).f32 > 0).permute(0, 3, 1, 2).cpu())
        elif kind == 'dense' and rec['act']:
            out.append((rec['y'] > 0).cpu())
    return out


class _MaskFeed:
    """Stand-in for oracle.networks_ref.leaky_relu: the same function wherever the oracle's own branch agrees with
    the device's (everywhere except pre-activations within rounding of zero), and the device's branch there."""

    def __init__(self, masks, alpha=0.2):
        self.q, self.alpha, self.flips, self.total = list(masks), alpha, 0, 0

    def __call__(self, x, alpha=None):
        m = self.q.pop(0)
        assert tuple(m.shape) == tuple(x.shape), (tuple(m.shape), tuple(x.shape))
        self.flips += int(((x.detach() > 0) != m).sum())
        self.total += m.numel()
        return x * torch.where(m, 1.0, self.alpha).to(x.dtype)


def test_eg_wgan_gradients_with_shared_masks(case, monkeypatch):
    c = case
    sink = []
    rep, grads = _device_eg(c, crop_aware=False, sink=sink)      # whole canvases: mask shapes == the oracle's
    tapes = {}
    for label, net, tape in sink:
        tapes.setdefault(label, []).append((net, tape))
    order = [tapes['E_zg'][0], tapes['E_zl'][0], tapes['G_rec'][0], tapes['critic_fixed'][0], tapes['G_interp'][0],
             tapes['critic_fixed'][1], tapes['G_blend'][0], tapes['critic_fixed'][2]]     # oracle call order
    feed = _MaskFeed([m for net, tape in order for m in _masks_of(net, tape)])
    del sink, tapes, order
    monkeypatch.setattr(R, 'leaky_relu', feed)
    P = {k: R.to_torch(c['params'][k], requires_grad=k in ('E_zg', 'E_zl', 'G')) for k in LOSS_NETS}
    loss, terms = L.EG_wgan(P, torch.from_numpy(c['reals']), c['idx'], c['crops']['eg_crop_interp'],
                            c['crops']['eg_crop_blend'], torch.from_numpy(c['mixes']['eg_mix']), scale_h=c['sh'],
                            scale_w=c['sw'])
    loss.mean().backward()
    assert not feed.q
    for k in ('rec_G', 'rec_pixel', 'interp_G', 'blend_G'):
        want = float(terms[k].mean())
        assert abs(rep[k] - want) <= TOL_TERM * max(1.0, abs(want)), (k, rep[k], want)
    worst = ('', 0.0)
    for k in ('E_zg', 'E_zl', 'G'):
        for name, t in P[k].items():
            if name == 'lod' or t.grad is None or float(t.grad.abs().max()) == 0:
                continue
            err = _rel(c['nets'][k].grad_view(grads[k], name).cpu().numpy(), t.grad.numpy())
            worst = max(worst, (k + '/' + name, err), key=lambda p: p[1])
            assert err <= TOL_GRAD_MASKED, (k, name, err)
    print('EG_wgan, shared masks (%d of %d branches differ): worst variable gradient rel-L2' % (feed.flips, feed.total),
          worst)


@pytest.mark.parametrize('which', ['D_rec', 'D_interp', 'D_blend'])
def test_critic_gradients_with_shared_masks(case, which, monkeypatch):
    c = case
    sink = []
    rep, fg = _device_critic(c, which, crop_aware=True, sink=sink)
    by = {label: (net, tape) for label, net, tape in sink}
    order = [by['critic_fake'], by['critic_real'], by['critic_mixed']]                    # oracle call order
    feed = _MaskFeed([m for net, tape in order for m in _masks_of(net, tape)])
    del sink, by, order
    P = {k: R.to_torch(c['params'][k], requires_grad=(k == which)) for k in LOSS_NETS}
    x = torch.from_numpy(c['reals'])
    mixes = {k: torch.from_numpy(v) for k, v in c['mixes'].items()}
    # the fakes are constants of the critic's loss: evaluate them with the oracle's own branches, then feed masks
    with torch.no_grad():
        if which == 'D_rec':
            zg, zl = L._fakes(P, x, R.CONFIG)
            fake = R.G_res(zg.repeat(1, 1, 32, 32), zl, P['G'], **R.CONFIG['G_res'])
            gp = 'd_rec_gp'
        else:
            fake, gp = None, 'd_interp_gp' if which == 'D_interp' else 'd_blend_gp'
    if fake is None:
        # reuse the oracle's own composition for the fake, with the critic evaluations redirected to the mask feed
        orig = L.D_wgangp
        monkeypatch.setattr(L, 'D_wgangp', lambda P_D, fakes, reals, mf, **kw: (fakes, None))
        if which == 'D_interp':
            fake, _ = L.D_interp_wgangp(P, x, c['idx'], c['crops']['d_interp_crop'], mixes[gp], c['sh'], c['sw'])
        else:
            fake, _ = L.D_blend_wgangp(P, x, c['idx'], c['crops']['d_blend_crop'], mixes['d_blend_mix'], mixes[gp],
                                       c['sh'], c['sw'])
        monkeypatch.setattr(L, 'D_wgangp', orig)
    monkeypatch.setattr(R, 'leaky_relu', feed)
    loss, terms = L.D_wgangp(P[which], fake, x, mixes[gp])
    loss.mean().backward()
    assert not feed.q
    for k in ('D_loss', 'gradient_penalty', 'epsilon_penalty'):
        want = float(terms[k].mean())
        assert abs(rep[k] - want) <= TOL_TERM * max(abs(want), 1e-2), (which, k, rep[k], want)
    worst = ('', 0.0)
    for name, t in P[which].items():
        if name == 'lod' or t.grad is None or float(t.grad.abs().max()) == 0:
            continue
        err = _rel(c['nets'][which].grad_view(fg, name).cpu().numpy(), t.grad.numpy())
        worst = max(worst, (name, err), key=lambda p: p[1])
        assert err <= TOL_GRAD_MASKED, (which, name, err)
    print(which, 'shared masks (%d of %d branches differ): worst variable gradient rel-L2' % (feed.flips, feed.total),
          worst)


# ---------------------------------------------------------------------- (c) config-off interpolation modes
@pytest.mark.parametrize('crop_aware', [True, False])
@pytest.mark.parametrize('tag,zg,zl', MODE_CASES)
def test_eg_wgan_interp_modes_vs_reference_golden(tag, zg, zl, crop_aware):
    """zg_interp_variational = 'variational', zl_interp_variational = 'hard' | 'variational' | 'random'
    (loss.py:176-193, 218-235) on the device - sampled canvases (tmx_latent_noise_fwd / _bwd), gradients into the
    encoders' mu AND log_sigma outputs - against tests/golden/losses_modes.npz (the reference's loss.py on the shim)."""
    import os
    from texturemixer_b200 import loss as dev_loss
    from texturemixer_b200.network import Network
    g = np.load(os.path.join(os.path.dirname(GOLDEN), 'losses_modes.npz'))
    n, sh, sw, _ = (int(v) for v in g['meta_n_sh_sw_stride'])
    params, reals, idx, crops, mixes = loss_case_inputs(n, sh, sw)
    nets = {}
    for k in LOSS_NETS:
        f = LOSS_FUNCS[k]
        nets[k] = Network(k, func='networks.' + f, seed=0, num_channels=3, resolution=128, **R.CONFIG[f])
        nets[k].set_vars(params[k])
    nets['G_fcn'] = Network('G', func='networks.G_res', reuse=True, share_vars_with=nets['G'], num_channels=3,
                            resolution=128, scale_h=sh, scale_w=sw, **R.CONFIG['G_res'])
    modes = dev_loss.interp_modes(zg, zl, {k: _dev(v) for k, v in mode_noise(n, 128, 32, sh, sw).items()})
    grads = {k: torch.zeros_like(nets[k].flat) for k in ('E_zg', 'E_zl', 'G')}
    rep = dev_loss.EG_wgan(nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], nets['G_fcn'], nets['D_interp'],
                           nets['D_blend'], _dev(reals), idx, crops['eg_crop_interp'], crops['eg_crop_blend'],
                           _dev(mixes['eg_mix']), grads, scale_h=sh, scale_w=sw, crop_aware=crop_aware, modes=modes)
    torch.cuda.synchronize()
    rep = {k: float(v.reshape(-1)[0]) for k, v in rep.items()}
    for mine, ref in (('rec_G', 'rec_G_loss'), ('rec_pixel', 'rec_pixel_loss'), ('interp_G', 'crop_interp_G_loss'),
                      ('blend_G', 'crop_blend_interp_G_loss')):
        want = float(g['EG%s_term_Loss_%s' % (tag, ref)].mean())
        assert abs(rep[mine] - want) <= TOL_TERM * max(1.0, abs(want)), (mine, rep[mine], want)
    c = dict(g=g, nets=nets)
    worst = _grads_vs_golden(c, 'EG' + tag, ('E_zg', 'E_zl', 'G'), grads, TOL_GRAD_FLIPS)
    print('EG_wgan zg=%s zl=%s crop_aware=%s vs reference fixture: worst variable gradient rel-L2 (flip-limited)'
          % (zg, zl, crop_aware), worst)
