"""ORACLE (test infrastructure): torch restatement of the E/G loss of the train step,
`loss.EG_wgan` (/root/reference/loss.py:105-259) with the reference config
(config.py:50-68: zg 'hard', zl 'permutational', block_size 0) and
gram_weight = 0 (VGG weights are not available) - autograd supplies the
gradients the device backward is checked against.

The random draws of the graph (crop offsets of `random_crop`, loss.py:78-90, and the
blend `mixing_factors`, loss.py:237) are explicit arguments so that both sides see
the same values; index vectors replace the 0/1 permutation matrices."""
import torch

from . import networks_ref as R


def tiling_permutation(x, scale_h, scale_w, idx_h, idx_w):
    """loss.py:92-100 in gather form (differentiable).  x [N,C,h,w]; idx_h [N,h*sh]; idx_w [N,w*sw] int."""
    n, c, h, w = x.shape
    t = x.repeat(1, 1, scale_h, scale_w)
    ih = torch.as_tensor(idx_h, dtype=torch.long)
    iw = torch.as_tensor(idx_w, dtype=torch.long)
    t = torch.stack([t[b][:, ih[b]][:, :, iw[b]] for b in range(n)])          # P_h @ tile(X) @ P_w
    t1 = torch.cat([x, t[:, :, :h, w:-w], x], dim=3)
    t2 = t[:, :, h:-h, :]
    t3 = torch.cat([x, t[:, :, -h:, w:-w], x], dim=3)
    return torch.cat([t1, t2, t3], dim=2)


def crop(images, yx, size):
    """loss.py:78-90 with the drawn offset made explicit: one (y, x) for the whole batch."""
    y, x = yx
    return images[:, :, y:y + size[0], x:x + size[1]]


def zg_canvas(zg_mu, zg_ls, H, W, mode='hard', eps=None):
    """loss.py:175-180 / 217-222 (callers pass the batch-reversed codes for the blend's second branch)."""
    if mode == 'hard':
        return zg_mu.repeat(1, 1, H, W)
    assert mode == 'variational'
    return (eps * torch.exp(zg_ls) + zg_mu).repeat(1, 1, H, W)


def zl_canvas(zl_mu, zl_ls, scale_h, scale_w, idx_h, idx_w, mode='permutational', eps=None):
    """loss.py:181-194 / 223-236.  `eps` [N,C,H,W]: the tf.random_normal draws laid out on the canvas ('random' uses
    the non-corner regions: the reference concatenates three separately drawn blocks, loss.py:189-192)."""
    if mode == 'permutational':
        return tiling_permutation(zl_mu, scale_h, scale_w, idx_h, idx_w)
    if mode == 'hard':
        return zl_mu.repeat(1, 1, scale_h, scale_w)
    if mode == 'variational':
        return eps * torch.exp(zl_ls.repeat(1, 1, scale_h, scale_w)) + zl_mu.repeat(1, 1, scale_h, scale_w)
    assert mode == 'random'
    h, w = zl_mu.shape[2:]
    r1 = torch.cat([zl_mu, eps[:, :, :h, w:-w], zl_mu], dim=3)
    r2 = eps[:, :, h:-h, :]
    r3 = torch.cat([zl_mu, eps[:, :, -h:, w:-w], zl_mu], dim=3)
    return torch.cat([r1, r2, r3], dim=2)


def EG_wgan(P, reals, idx, crop_interp, crop_blend, mixing_factors, scale_h=3, scale_w=3, rec_G_weight=1.0,
            pixel_weight=200.0, kl_weight=0.0, interp_G_weight=1.0, blend_interp_G_weight=1.0, cfg=None,
            gram_weight=0.0, vgg=None, gram_alpha=None, zg_mode='hard', zl_mode='permutational', noise=None):
    """P: dict of parameter dicts for 'E_zg','E_zl','G','D_rec','D_interp','D_blend'.  Returns the per-sample
    loss vector [N] (the optimizer differentiates its mean, run.py:321) and a dict of named terms.
    gram_weight > 0 adds the VGG-19 Gram terms (loss.py:148-160, 206-213, 248-257) with `vgg` = the weight dict
    ({layer: [filter, bias]}) and `gram_alpha` = the [N,1,1,1] draw of loss.py:253; the loss then has the
    reference's [N,1,1,N] shape (see vgg_ref.blend_gram_term).
    zg_mode / zl_mode: zg_interp_variational / zl_interp_variational (config.py:61-62: 'hard' / 'permutational');
    `noise`: {'zg_f', 'zl_f', 'zg_b', 'zl_b'} standard-normal tensors for the sampling modes (see zl_canvas)."""
    cfg = cfg or R.CONFIG
    noise = noise or {}
    zg_mu, zg_ls = R.E_zg(reals, P['E_zg'], **cfg['E_zg'])                           # loss.py:119
    zl_mu, zl_ls = R.E_zl(reals, P['E_zl'], **cfg['E_zl'])                           # loss.py:126
    lat = zl_mu.shape[2]
    rec = R.G_res(zg_mu.repeat(1, 1, lat, lat), zl_mu, P['G'], **cfg['G_res'])        # loss.py:130
    terms = {}
    loss = 0
    if rec_G_weight > 0:
        terms['rec_G'] = (-R.D_patch(rec, P['D_rec'], **cfg['D_patch'])).mean(dim=(1, 2, 3)) * rec_G_weight
        loss = loss + terms['rec_G']
    if pixel_weight > 0:
        terms['rec_pixel'] = (rec - reals).abs().mean(dim=(1, 2, 3)) * pixel_weight  # loss.py:143
        loss = loss + terms['rec_pixel']
    if gram_weight > 0:                                                               # loss.py:149-160
        from . import vgg_ref as V
        real_gram = V.grams(reals, vgg)
        terms['rec_gram'] = V.multi_layer_diff(V.grams(rec, vgg), real_gram) * gram_weight
        loss = loss + terms['rec_gram']
    if kl_weight > 0:                                                                 # loss.py:163-171
        for tag, mu, ls in (('KL_zg', zg_mu, zg_ls), ('KL_zl', zl_mu, zl_ls)):
            terms[tag] = -0.5 * (1 + 2 * ls - mu ** 2 - torch.exp(2 * ls)).mean(dim=(1, 2, 3)) * kl_weight
            loss = loss + terms[tag]
    g_cfg = dict(cfg['G_res'], scale_h=scale_h, scale_w=scale_w)
    zg_c = zg_canvas(zg_mu, zg_ls, lat * scale_h, lat * scale_w, zg_mode, noise.get('zg_f'))       # loss.py:175-180
    zl_c = zl_canvas(zl_mu, zl_ls, scale_h, scale_w, idx['h_forward'], idx['w_forward'], zl_mode,
                     noise.get('zl_f'))                                                             # loss.py:181-194
    size = reals.shape[2:]
    if interp_G_weight > 0:
        interp = R.G_res(zg_c, zl_c, P['G'], **g_cfg)                                 # loss.py:197
        crop_i = crop(interp, crop_interp, size)
        terms['interp_G'] = (-R.D_patch(crop_i, P['D_interp'], **cfg['D_patch'])).mean(dim=(1, 2, 3)) * interp_G_weight
        loss = loss + terms['interp_G']
        if gram_weight > 0:                                                           # loss.py:206-213
            terms['interp_gram'] = V.multi_layer_diff(V.grams(crop_i, vgg), real_gram) * gram_weight
            loss = loss + terms['interp_gram']
    if blend_interp_G_weight > 0:
        zg_r = zg_canvas(torch.flip(zg_mu, dims=[0]), torch.flip(zg_ls, dims=[0]), lat * scale_h, lat * scale_w,
                         zg_mode, noise.get('zg_b'))                                              # loss.py:217-222
        zl_r = zl_canvas(torch.flip(zl_mu, dims=[0]), torch.flip(zl_ls, dims=[0]), scale_h, scale_w,
                         idx['h_backward'], idx['w_backward'], zl_mode, noise.get('zl_b'))        # loss.py:223-236
        t = mixing_factors
        bzg = zg_r + (zg_c - zg_r) * t                                                            # loss.py:238
        bzl = zl_r + (zl_c - zl_r) * t
        blend = R.G_res(bzg, bzl, P['G'], **g_cfg)                                                # loss.py:240
        crop_b = crop(blend, crop_blend, size)
        terms['blend_G'] = (-R.D_patch(crop_b, P['D_blend'], **cfg['D_patch'])).mean(dim=(1, 2, 3)) * blend_interp_G_weight
        loss = loss + terms['blend_G']
        if gram_weight > 0:                                                                       # loss.py:248-257
            terms['blend_gram'] = V.blend_gram_term(V.grams(crop_b, vgg), real_gram, gram_alpha, gram_weight)
            loss = loss + terms['blend_gram']
    return loss, terms


def D_wgangp(P_D, fakes, reals, mixing_factors, wgan_lambda=10.0, wgan_epsilon=0.001, wgan_target=1.0, cfg=None):
    """The critic loss of loss.py:303-346 (D_rec_wgangp; D_interp_/D_blend_wgangp, :351-521, differ only in how
    the fake images are produced).  `fakes` are constants here (D's optimizer only sees D's variables,
    run.py:322-324).  Returns the per-sample loss [N] and the named terms; the gradient penalty differentiates
    through the gradient (create_graph), like tf.gradients of loss.py:333 inside tfutil.py:299."""
    cfg = cfg or R.CONFIG
    s_f = R.D_patch(fakes, P_D, **cfg['D_patch'])
    s_r = R.D_patch(reals, P_D, **cfg['D_patch'])
    terms = {'D_loss': (s_f - s_r).mean(dim=(1, 2, 3))}                                    # loss.py:323
    mixed = (reals + (fakes - reals) * mixing_factors).detach().requires_grad_(True)      # loss.py:329-330
    s_m = R.D_patch(mixed, P_D, **cfg['D_patch'])
    (g,) = torch.autograd.grad(s_m.sum(), mixed, create_graph=True)                       # loss.py:332-333
    norms = torch.sqrt((g * g).sum(dim=(1, 2, 3)))
    terms['gradient_penalty'] = (norms - wgan_target) ** 2 * (wgan_lambda / wgan_target ** 2)   # loss.py:335-336
    terms['epsilon_penalty'] = (s_r * s_r).mean(dim=(1, 2, 3)) * wgan_epsilon              # loss.py:342
    loss = terms['D_loss'] + terms['gradient_penalty'] + terms['epsilon_penalty']
    return loss, terms


# ---------------------------------------------------------------------- the three critic losses, fakes included
def _fakes(P, reals, cfg):
    """Encoders + reconstruction shared by the three critic losses (loss.py:308-320, 360-369, 435-444)."""
    zg_mu, _ = R.E_zg(reals, P['E_zg'], **cfg['E_zg'])
    zl_mu, _ = R.E_zl(reals, P['E_zl'], **cfg['E_zl'])
    return zg_mu, zl_mu


def D_rec_wgangp(P, reals, mixing_factors, cfg=None, **kw):
    """loss.py:303-346.  P: parameter dicts of 'E_zg','E_zl','G','D_rec'.  The fakes are constants for the critic's
    optimizer (run.py:322), so they are evaluated without autograd."""
    cfg = cfg or R.CONFIG
    with torch.no_grad():
        zg_mu, zl_mu = _fakes(P, reals, cfg)
        lat = zl_mu.shape[2]
        rec = R.G_res(zg_mu.repeat(1, 1, lat, lat), zl_mu, P['G'], **cfg['G_res'])        # loss.py:320
    return D_wgangp(P['D_rec'], rec, reals, mixing_factors, cfg=cfg, **kw)


def D_interp_wgangp(P, reals, idx, crop_yx, mixing_factors, scale_h=3, scale_w=3, cfg=None, **kw):
    """loss.py:351-421 ('hard' zg, 'permutational' zl)."""
    cfg = cfg or R.CONFIG
    with torch.no_grad():
        zg_mu, zl_mu = _fakes(P, reals, cfg)
        lat = zl_mu.shape[2]
        zg_c = zg_mu.repeat(1, 1, lat * scale_h, lat * scale_w)                           # loss.py:373
        zl_c = tiling_permutation(zl_mu, scale_h, scale_w, idx['h_forward'], idx['w_forward'])   # loss.py:391
        img = R.G_res(zg_c, zl_c, P['G'], **dict(cfg['G_res'], scale_h=scale_h, scale_w=scale_w))  # loss.py:394
        fake = crop(img, crop_yx, reals.shape[2:])                                        # loss.py:395
    return D_wgangp(P['D_interp'], fake, reals, mixing_factors, cfg=cfg, **kw)


def D_blend_wgangp(P, reals, idx, crop_yx, blend_mixing_factors, mixing_factors, scale_h=3, scale_w=3, cfg=None, **kw):
    """loss.py:426-521: the blend draws its own mixing factors (loss.py:489), then the penalty's (loss.py:505)."""
    cfg = cfg or R.CONFIG
    with torch.no_grad():
        zg_mu, zl_mu = _fakes(P, reals, cfg)
        lat = zl_mu.shape[2]
        zg_c = zg_mu.repeat(1, 1, lat * scale_h, lat * scale_w)
        zl_c = tiling_permutation(zl_mu, scale_h, scale_w, idx['h_forward'], idx['w_forward'])
        zg_r = torch.flip(zg_mu, dims=[0]).repeat(1, 1, lat * scale_h, lat * scale_w)      # loss.py:470
        zl_r = tiling_permutation(torch.flip(zl_mu, dims=[0]), scale_h, scale_w, idx['h_backward'],
                                  idx['w_backward'])                                       # loss.py:488
        t = blend_mixing_factors
        img = R.G_res(zg_r + (zg_c - zg_r) * t, zl_r + (zl_c - zl_r) * t, P['G'],
                      **dict(cfg['G_res'], scale_h=scale_h, scale_w=scale_w))              # loss.py:490-494
        fake = crop(img, crop_yx, reals.shape[2:])
    return D_wgangp(P['D_blend'], fake, reals, mixing_factors, cfg=cfg, **kw)
