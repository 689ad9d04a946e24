"""Pin the loss oracle (oracle/loss_ref.py) against tests/golden/losses.npz - the reference's OWN loss.py
(/root/reference/loss.py:105-259, 303-521) executed unmodified on oracle/tfshim by tests/golden/make_golden.py
gen_losses: per-sample losses, every autosummary'd term, and d mean(loss) / d variable for the variables each
optimizer owns (run.py:321-324).  CPU only; the device path is checked against the same fixture in
tests/test_gpu_loss_golden.py."""
import numpy as np
import pytest
import torch

from oracle import loss_ref as L
from oracle import networks_ref as R

from loss_case import (GOLDEN, GRAM_WEIGHT, MODE_CASES, loss_case_inputs, golden_gradient, gram_alpha, mode_noise, subsample,
                       vgg_standin_weights)


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))


@pytest.fixture(scope='module')
def case():
    g = np.load(GOLDEN)
    n, sh, sw, stride = (int(v) for v in g['meta_n_sh_sw_stride'])
    params, reals, idx, crops, mixes = loss_case_inputs(n, sh, sw)
    for k, v in crops.items():                      # the recipe reproduces the stored draws
        assert tuple(g['draw_' + k]) == v
    for k, v in mixes.items():
        assert np.array_equal(g['draw_' + k], v)
    return g, params, torch.from_numpy(reals), idx, crops, {k: torch.from_numpy(v) for k, v in mixes.items()}, (sh, sw)


def _check_grads(g, tag, P, scopes, tol):
    worst = ('', 0.0)
    for scope in scopes:
        for name, t in P[scope].items():
            if name == 'lod':
                continue
            want, norm = golden_gradient(g, tag, scope, name)
            got = np.zeros(tuple(t.shape), np.float32) if t.grad is None else t.grad.numpy()
            assert abs(np.linalg.norm(got.astype(np.float64)) - norm) <= tol * max(norm, 1e-12), (scope, name)
            if norm == 0.0:
                assert np.abs(got).max() == 0.0, (scope, name)      # unused lod heads: zero on both sides
                continue
            err = _rel(subsample(got), want)
            worst = max(worst, (scope + '/' + name, err), key=lambda p: p[1])
            assert err <= tol, (scope, name, err)
    return worst


def test_eg_wgan_matches_reference_loss_py(case):
    g, params, reals, idx, crops, mixes, (sh, sw) = case
    P = {k: R.to_torch(params[k], requires_grad=k in ('E_zg', 'E_zl', 'G')) for k in params}
    loss, terms = L.EG_wgan(P, reals, idx, crops['eg_crop_interp'], crops['eg_crop_blend'], mixes['eg_mix'],
                            scale_h=sh, scale_w=sw)
    loss.mean().backward()
    assert np.allclose(loss.detach().numpy(), g['EG_loss'], rtol=2e-6, atol=0)
    for mine, ref in (('rec_G', 'rec_G_loss'), ('rec_pixel', 'rec_pixel_loss'), ('interp_G', 'crop_interp_G_loss'),
                      ('blend_G', 'crop_blend_interp_G_loss')):
        want = g['EG_term_Loss_' + ref]
        assert np.allclose(terms[mine].detach().numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max()), mine
    worst = _check_grads(g, 'EG', P, ('E_zg', 'E_zl', 'G'), 1e-4)
    print('EG_wgan: worst variable gradient rel-L2 vs the reference code', worst)


@pytest.mark.parametrize('which', ['D_rec', 'D_interp', 'D_blend'])
def test_critic_losses_match_reference_loss_py(case, which):
    g, params, reals, idx, crops, mixes, (sh, sw) = case
    P = {k: R.to_torch(params[k], requires_grad=(k == which)) for k in params}
    if which == 'D_rec':
        loss, terms = L.D_rec_wgangp(P, reals, mixes['d_rec_gp'])
        names = dict(D_loss='rec_D_loss', gradient_penalty='rec_gradient_penalty', epsilon_penalty='rec_epsilon_penalty')
    elif which == 'D_interp':
        loss, terms = L.D_interp_wgangp(P, reals, idx, crops['d_interp_crop'], mixes['d_interp_gp'], sh, sw)
        names = dict(D_loss='crop_interp_D_loss', gradient_penalty='crop_interp_gradient_penalty',
                     epsilon_penalty='crop_interp_epsilon_penalty')
    else:
        loss, terms = L.D_blend_wgangp(P, reals, idx, crops['d_blend_crop'], mixes['d_blend_mix'], mixes['d_blend_gp'],
                                       sh, sw)
        names = dict(D_loss='crop_blend_interp_D_loss', gradient_penalty='crop_blend_interp_gradient_penalty',
                     epsilon_penalty='crop_blend_interp_epsilon_penalty')
    loss.mean().backward()
    assert np.allclose(loss.detach().numpy(), g[which + '_loss'], rtol=1e-5, atol=0)
    for mine, ref in names.items():
        want = g['%s_term_Loss_%s' % (which, ref)]
        assert np.allclose(terms[mine].detach().numpy(), want, rtol=1e-4, atol=1e-5 * max(np.abs(want).max(), 1e-3)), mine
    worst = _check_grads(g, which, P, (which,), 1e-4)
    print(which, 'worst variable gradient rel-L2 vs the reference code', worst)


def test_eg_wgan_with_gram_terms_matches_reference_code(case):
    """EG_wgan WITH the VGG-19 Gram terms (config.py:64 gram_weight = 0.002): tests/golden/losses_gram.npz comes from
    the reference's loss.py + custom_vgg19.py run unmodified on the shim (stand-in weights; tensorflow_vgg's base
    class restated in oracle/tfshim/tensorflow_vgg) - incl. the [N,1,1,N] broadcast of loss.py:254."""
    import os
    _, params, reals, idx, crops, mixes, (sh, sw) = case
    g = np.load(os.path.join(os.path.dirname(GOLDEN), 'losses_gram.npz'))
    alpha = gram_alpha(reals.shape[0])
    assert np.array_equal(g['draw_eg_gram_alpha'], alpha)
    P = {k: R.to_torch(params[k], requires_grad=k in ('E_zg', 'E_zl', 'G')) for k in params}
    loss, terms = L.EG_wgan(P, reals, idx, crops['eg_crop_interp'], crops['eg_crop_blend'], mixes['eg_mix'],
                            scale_h=sh, scale_w=sw, gram_weight=GRAM_WEIGHT, vgg=vgg_standin_weights(),
                            gram_alpha=torch.from_numpy(alpha))
    loss.mean().backward()
    assert tuple(loss.shape) == tuple(g['EGgram_loss'].shape) == (4, 1, 1, 4)
    assert np.allclose(loss.detach().numpy(), g['EGgram_loss'], rtol=1e-5, atol=0)
    for mine, ref in (('rec_gram', 'rec_gram_loss'), ('interp_gram', 'crop_interp_gram_loss'),
                      ('blend_gram', 'crop_blend_interp_gram_loss'), ('rec_G', 'rec_G_loss')):
        want = g['EGgram_term_Loss_' + ref]
        assert np.allclose(terms[mine].detach().numpy(), want, rtol=1e-4, atol=1e-5 * np.abs(want).max()), mine
    worst = _check_grads(g, 'EGgram', P, ('E_zg', 'E_zl', 'G'), 2e-4)
    print('EG_wgan + Gram: worst variable gradient rel-L2 vs the reference code', worst)


@pytest.mark.parametrize('tag,zg,zl', MODE_CASES)
def test_eg_wgan_interp_modes_match_reference_code(tag, zg, zl):
    """The config-off interpolation modes (zg_interp_variational = 'variational'; zl_interp_variational = 'hard' |
    'variational' | 'random', loss.py:176-193, 218-235): tests/golden/losses_modes.npz comes from the reference's
    loss.py with its tf.random_normal draws fed from tests/loss_case.mode_noise in the graph's own call order."""
    import os
    g = np.load(os.path.join(os.path.dirname(GOLDEN), 'losses_modes.npz'))
    n, sh, sw, _ = (int(v) for v in g['meta_n_sh_sw_stride'])
    params, reals, idx, crops, mixes = loss_case_inputs(n, sh, sw)
    noise = {k: torch.from_numpy(v) for k, v in mode_noise(n, 128, 32, sh, sw).items()}
    P = {k: R.to_torch(params[k], requires_grad=k in ('E_zg', 'E_zl', 'G')) for k in params}
    loss, terms = L.EG_wgan(P, torch.from_numpy(reals), idx, crops['eg_crop_interp'], crops['eg_crop_blend'],
                            torch.from_numpy(mixes['eg_mix']), scale_h=sh, scale_w=sw, zg_mode=zg, zl_mode=zl,
                            noise=noise)
    loss.mean().backward()
    assert np.allclose(loss.detach().numpy(), g['EG%s_loss' % tag], rtol=1e-5, atol=0)
    for mine, ref in (('rec_G', 'rec_G_loss'), ('interp_G', 'crop_interp_G_loss'), ('blend_G', 'crop_blend_interp_G_loss')):
        want = g['EG%s_term_Loss_%s' % (tag, ref)]
        assert np.allclose(terms[mine].detach().numpy(), want, rtol=1e-4, atol=1e-5 * np.abs(want).max()), mine
    worst = _check_grads(g, 'EG' + tag, P, ('E_zg', 'E_zl', 'G'), 2e-4)
    print('EG_wgan zg=%s zl=%s: worst variable gradient rel-L2 vs the reference code' % (zg, zl), worst)
