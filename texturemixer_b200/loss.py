"""E/G phase of the train step: `loss.EG_wgan` (loss.py:105-259) evaluated and
differentiated on the device.  Reference config (config.py:50-68): zg 'hard', zl
'permutational', kl_weight = 0; gram_weight is forced to 0 (VGG-19 weights are not
redistributable, SURVEY §2 - stated deviation).

The forward builds exactly the reference graph (encoders once, G at scale 1, G_fcn
twice on the 3x3 canvases, the three critics as fixed functions); the backward is the
explicit reverse of it: critic input gradients -> crop adjoint -> G / G_fcn backward
-> lerp / tiling_permutation / tile adjoints -> encoder backward.  Variable gradients
are accumulated into one flat buffer per network (E_zg, E_zl, G) for `Optimizer`.

`random_crop` offsets (loss.py:78-90) and the blend `mixing_factors` (loss.py:237) are
passed in by the caller, who owns the RNG (one shared offset per rank like the reference)."""
import ctypes as C

import numpy as np
import torch

from . import _lib, interp
from .backward import backward
from .runtime import Runtime


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _tile_code(rt, z, H, W):
    """tf.tile of a [N,C,1,1] code over an H x W canvas (loss.py:130,176)."""
    return rt.latent_blend([z.contiguous()], H, W, _lib.BLEND_COPY)


def _row_sum(rt, x, rows, length, out=None, scale=1.0, accumulate=False):
    if out is None:
        out = rt.empty(rows)
    _lib.check(rt.lib.tmx_row_sum(rt.handle, _ptr(x), _ptr(out), rows, length, float(scale), int(accumulate),
                                  rt.stream()), 'tmx_row_sum')
    return out


def _gather_bwd(rt, dcanvas, dsrc, idx_h, idx_w, pins=(0, 0), reverse=False):
    n, c, H, W = dcanvas.shape
    h, w = dsrc.shape[2:]
    _lib.check(rt.lib.tmx_latent_gather_bwd(rt.handle, _ptr(dcanvas.contiguous()), _ptr(dsrc), _ptr(idx_h), _ptr(idx_w),
                                            n, c, h, w, H, W, pins[0], pins[1], int(reverse), rt.stream()),
               'tmx_latent_gather_bwd')


def _dev_idx(rt, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(rt.device)


def _crop_adjoint(dcrop, full_hw, yx):
    """Adjoint of random_crop (loss.py:78-90): zeros outside the window (pure data movement)."""
    n, c, h, w = dcrop.shape
    full = torch.zeros(n, c, full_hw[0], full_hw[1], dtype=torch.float32, device=dcrop.device)
    full[:, :, yx[0]:yx[0] + h, yx[1]:yx[1] + w].copy_(dcrop)
    return full


def EG_wgan(E_zg, E_zl, G, D_rec, G_fcn, D_interp, D_blend, reals, idx, crop_interp, crop_blend, mixing_factors,
            grads, scale_h=3, scale_w=3, rec_G_weight=1.0, pixel_weight=200.0, interp_G_weight=1.0,
            blend_interp_G_weight=1.0):
    """One evaluation + differentiation of mean(EG_loss) (run.py:321).
    reals: [N,3,R,R] fp32 device tensor in [-1,1]; idx: dict of int32 index vectors (interp.sample_schedule_indices);
    crop_*: (y, x); mixing_factors: [N,1,1,1] fp32 device tensor; grads: {'E_zg','E_zl','G'} -> flat gradient
    buffers (accumulated into).  Returns a dict of per-term batch means (device scalars)."""
    rt = Runtime.get(reals.device)
    n, _, res, _ = reals.shape
    inv_n = 1.0 / n
    report = {}

    # ---------------- forward
    t_zg, t_zl = [], []
    zg_mu, _ = E_zg.get_output_for(reals, tape=t_zg)
    zl_mu, _ = E_zl.get_output_for(reals, tape=t_zl)
    c, lat = zl_mu.shape[1], zl_mu.shape[2]
    H, W = lat * scale_h, lat * scale_w
    pins = interp._corner_pins(scale_h, scale_w)

    t_rec, t_drec = [], []
    rec = G.get_output_for(_tile_code(rt, zg_mu, lat, lat), zl_mu, tape=t_rec)
    d_rec_img = None
    if rec_G_weight > 0:
        s = D_rec.get_output_for(rec, tape=t_drec)
        report['rec_G'] = _row_sum(rt, s, 1, n, scale=-rec_G_weight * inv_n)
        (d_rec_img,) = backward(D_rec, t_drec, [torch.full_like(s, -rec_G_weight * inv_n)], None, param_grads=False)
    if pixel_weight > 0:
        l1 = rt.empty(*rec.shape)
        lsum = torch.zeros(1, dtype=torch.float32, device=rt.device)
        per = rec[0].numel()
        _lib.check(rt.lib.tmx_loss_l1_grad(rt.handle, _ptr(rec), _ptr(reals.contiguous()), _ptr(l1), _ptr(lsum),
                                           rec.numel(), pixel_weight * inv_n / per, rt.stream()), 'tmx_loss_l1_grad')
        report['rec_pixel'] = lsum * (pixel_weight * inv_n / per)
        if d_rec_img is None:
            d_rec_img = l1
        else:
            d_rec_img = _add(rt, d_rec_img, l1)
    dzg_tiled, dzl = backward(G, t_rec, [d_rec_img], grads['G'])
    dzg = _row_sum(rt, dzg_tiled, n * c, lat * lat)                      # adjoint of the 32x32 tile of zg
    dzl = dzl.contiguous()

    if interp_G_weight > 0 or blend_interp_G_weight > 0:
        ih_f, iw_f = _dev_idx(rt, idx['h_forward']), _dev_idx(rt, idx['w_forward'])
        zg_c = _tile_code(rt, zg_mu, H, W)
        zl_c = rt.latent_blend([zl_mu.contiguous()], H, W, _lib.BLEND_COPY, idx_h=[ih_f], idx_w=[iw_f],
                               pin_rows=pins[0], pin_cols=pins[1])
    if interp_G_weight > 0:
        t_g, t_d = [], []
        img = G_fcn.get_output_for(zg_c, zl_c, tape=t_g)
        y0, x0 = crop_interp
        cr = img[:, :, y0:y0 + res, x0:x0 + res].contiguous()
        s = D_interp.get_output_for(cr, tape=t_d)
        report['interp_G'] = _row_sum(rt, s, 1, n, scale=-interp_G_weight * inv_n)
        (dcr,) = backward(D_interp, t_d, [torch.full_like(s, -interp_G_weight * inv_n)], None, param_grads=False)
        dzg_c, dzl_c = backward(G_fcn, t_g, [_crop_adjoint(dcr, img.shape[2:], crop_interp)], grads['G'])
        _row_sum(rt, dzg_c, n * c, H * W, out=dzg, accumulate=True)
        _gather_bwd(rt, dzl_c, dzl, ih_f, iw_f, pins)
        del t_g, t_d, img
    if blend_interp_G_weight > 0:
        ih_b, iw_b = _dev_idx(rt, idx['h_backward']), _dev_idx(rt, idx['w_backward'])
        zero1 = torch.zeros_like(zg_mu)
        # tf.reverse(axis=[0]) of the sources is folded into the gather (src_reverse)
        zg_r = rt.latent_blend([zg_mu.contiguous()], H, W, _lib.BLEND_COPY, src_reverse=1)
        zl_r = rt.latent_blend([zl_mu.contiguous()], H, W, _lib.BLEND_COPY, idx_h=[ih_b], idx_w=[iw_b],
                               pin_rows=pins[0], pin_cols=pins[1], src_reverse=1)
        t = mixing_factors.reshape(-1).contiguous()
        bzg = rt.latent_blend([zg_r, zg_c], H, W, _lib.BLEND_LERP, t=t)          # lerp(reverse, forward, t), loss.py:238
        bzl = rt.latent_blend([zl_r, zl_c], H, W, _lib.BLEND_LERP, t=t)
        t_g, t_d = [], []
        img = G_fcn.get_output_for(bzg, bzl, tape=t_g)
        y0, x0 = crop_blend
        cr = img[:, :, y0:y0 + res, x0:x0 + res].contiguous()
        s = D_blend.get_output_for(cr, tape=t_d)
        report['blend_G'] = _row_sum(rt, s, 1, n, scale=-blend_interp_G_weight * inv_n)
        (dcr,) = backward(D_blend, t_d, [torch.full_like(s, -blend_interp_G_weight * inv_n)], None, param_grads=False)
        dbzg, dbzl = backward(G_fcn, t_g, [_crop_adjoint(dcr, img.shape[2:], crop_blend)], grads['G'])
        zero_c = torch.zeros_like(dbzg)
        # adjoint of lerp: d forward = t * d, d reverse = d - t * d   (same fp32 ops as autograd of a + (b - a) * t)
        for d, dsrc_kind in ((dbzg, 'zg'), (dbzl, 'zl')):
            d_fwd = rt.latent_blend([zero_c, d.contiguous()], H, W, _lib.BLEND_LERP, t=t)
            d_rev = rt.latent_blend([d.contiguous(), zero_c], H, W, _lib.BLEND_LERP, t=t)
            if dsrc_kind == 'zg':
                _row_sum(rt, d_fwd, n * c, H * W, out=dzg, accumulate=True)
                tmp = _row_sum(rt, d_rev, n * c, H * W).view(n, c, 1, 1)
                _gather_bwd(rt, tmp, dzg.view(n, c, 1, 1), None, None, (0, 0), reverse=True)
            else:
                _gather_bwd(rt, d_fwd, dzl, ih_f, iw_f, pins)
                _gather_bwd(rt, d_rev, dzl, ih_b, iw_b, pins, reverse=True)
        del t_g, t_d, img, zero1

    # ---------------- encoders
    backward(E_zl, t_zl, [dzl, None], grads['E_zl'], want_input_grads=False)
    backward(E_zg, t_zg, [dzg.view(n, c, 1, 1), None], grads['E_zg'], want_input_grads=False)
    return report


def _add(rt, a, b):
    out = rt.empty(*a.shape)
    _lib.check(rt.lib.tmx_add_f32(rt.handle, _ptr(a.contiguous()), _ptr(b.contiguous()), _ptr(out), a.numel(),
                                  rt.stream()), 'tmx_add_f32')
    return out
