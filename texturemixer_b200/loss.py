"""E/G phase of the train step: `loss.EG_wgan` (loss.py:105-259) evaluated and
differentiated on the device.  Reference config (config.py:50-68): zg 'hard', zl
'permutational'; the KL term (kl_weight, 0 in the reference config) is supported; the VGG-19 Gram terms
(gram_weight, 0.002 in the reference config) run when a `vgg.GramLoss` is supplied - the weight file itself is not
redistributable, so benchmarks state whether they ran with a stand-in or without the term.

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
from . import runtime
from .backward import backward, conv_wgrad_into
from .runtime import Act
from .runtime import Runtime


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _tile_code(rt, z, H, W):
    """tf.tile of a [N,C,1,1] code over an H x W canvas (loss.py:130,176)."""
    return rt.latent_blend([z.contiguous()], H, W, _lib.BLEND_COPY)


def _row_sum(rt, x, rows, length, out=None, scale=1.0, accumulate=False, square=False):
    if out is None:
        out = rt.empty(rows)
    _lib.check(rt.lib.tmx_row_sum(rt.handle, _ptr(x), _ptr(out), rows, length, float(scale), int(accumulate),
                                  int(square), rt.stream()), 'tmx_row_sum')
    return out


def _gather_bwd(rt, dcanvas, dsrc, idx_h, idx_w, pins=(0, 0), reverse=False):
    n, c, H, W = dcanvas.shape
    h, w = dsrc.shape[2:]
    _lib.check(rt.lib.tmx_latent_gather_bwd(rt.handle, _ptr(dcanvas.contiguous()), _ptr(dsrc), _ptr(idx_h), _ptr(idx_w),
                                            n, c, h, w, H, W, pins[0], pins[1], int(reverse), rt.stream()),
               'tmx_latent_gather_bwd')


def _gather_bwd_win(rt, d, win, H, W, dsrc, idx_h, idx_w, pins=(0, 0), reverse=False):
    """`_gather_bwd(_embed(d, win))` without the zero canvas: scatter the window gradient from where it lies."""
    if win is None:
        return _gather_bwd(rt, d, dsrc, idx_h, idx_w, pins, reverse)
    n, c, wh, ww = d.shape
    h, w = dsrc.shape[2:]
    oy, ox = win[0], win[1]
    assert (wh, ww) == tuple(win[2:]), ((wh, ww), tuple(win))
    _lib.check(rt.lib.tmx_latent_gather_bwd_window(rt.handle, _ptr(d.contiguous()), _ptr(dsrc), _ptr(idx_h), _ptr(idx_w),
                                                   n, c, h, w, H, W, wh, ww, int(oy), int(ox),
                                                   _ptr(getattr(win, 'dev', None)), pins[0], pins[1], int(reverse),
                                                   rt.stream()), 'tmx_latent_gather_bwd_window')


def _dev_idx(rt, a):
    if isinstance(a, torch.Tensor):
        return a.to(device=rt.device, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(rt.device)


# Test hook: when set to a list, every forward tape recorded by this module is appended as (label, network, tape) in
# evaluation order - the parity tests read the leaky-ReLU branch masks of the device evaluation from it and hand them
# to the oracle, so that both sides differentiate the SAME piecewise-linear function (tests/test_gpu_loss_golden.py).
TAPE_SINK = None


def _sink(label, net, tape):
    if TAPE_SINK is not None:
        TAPE_SINK.append((label, net, tape))


G_CONTEXT = 14   # latent pixels of context one output pixel of G_res sees on each side: 12 convs at the latent
#                  resolution (networks.py:427-446) + 2 at 2x + 2 at 4x (:447-457) = 12 + 1 + 0.5, rounded up


def crop_window(yx, res, lat, H, W):
    """Latent window (oy, ox, wh, ww) of the H x W canvas outside of which nothing reaches the res x res image crop
    at pixel offset yx (SURVEY Appendix C note).  D only ever sees `random_crop(G_fcn(canvas))` (loss.py:198,241), the
    offset is drawn before decoding and every layer of G_res is local, so decoding just the window gives the same crop
    pixels and - gradients being zero outside the crop's cone of dependence - the same gradients; where the window
    ends inside the canvas its REFLECT padding differs from the true neighbours, which perturbs only pixels within
    G_CONTEXT of that edge, and those are outside the cone by construction.  None = whole canvas."""
    up = res // lat
    if up != 4 or res % lat:
        return None
    need = lat + 1 + 2 * G_CONTEXT                # an unaligned crop touches lat + 1 latent pixels

    def axis(c0, L):
        win = -(-need // lat) * lat               # a whole number of tiles: G_fcn at a smaller (scale_h, scale_w)
        if win >= L:
            return 0, L
        return min(max(c0 // up - G_CONTEXT, 0), L - win), win
    oy, wh = axis(yx[0], H)
    ox, ww = axis(yx[1], W)
    if wh == H and ww == W:
        return None
    return oy, ox, wh, ww


TAIL_CONTEXT = 2   # latent pixels of context the blocks ABOVE the latent resolution add: 2 convs at 2x + 2 at 4x = 1.5
TAIL_SIZE = 40     # >= (crop footprint 33) + 2 * TAIL_CONTEXT, a multiple of 8 (whole MMA tile rows at 2x and 4x)


MID_CONTEXT = 6    # after the 4th residual block: 2 residual convs + Conv0 + Conv1 at the latent resolution + the tail's 2
MID_SIZE = 48      # >= 33 + 2 * MID_CONTEXT, a multiple of 16


def mid_window(yx, res, lat, win, H, W):
    """Intermediate level of the crop-aware evaluation, applied after G_res's fourth residual block (8 of the 12
    latent-resolution convs done): (oy, ox, h, w) relative to the trunk window `win`.  None = keep everything."""
    up = res // lat
    if up != 4:
        return None
    oy0, ox0, wh, ww = (0, 0, H, W) if win is None else win
    if wh <= MID_SIZE and ww <= MID_SIZE:
        return None

    def axis(c0, o0, L):
        if L <= MID_SIZE:
            return 0, L
        return min(max(c0 // up - o0 - MID_CONTEXT - 1, 0), L - MID_SIZE), MID_SIZE
    oy, h = axis(yx[0], oy0, wh)
    ox, w = axis(yx[1], ox0, ww)
    return oy, ox, h, w


def compose_window(outer, inner, H, W):
    """Absolute (oy, ox, h, w) of window `inner` given relative to `outer` (either may be None)."""
    if inner is None:
        return outer
    oy0, ox0 = (0, 0) if outer is None else outer[:2]
    return oy0 + inner[0], ox0 + inner[1], inner[2], inner[3]


def tail_window(yx, res, lat, win, H, W):
    """Second level of the crop-aware evaluation: (oy, ox, h, w) in latent pixels RELATIVE to the trunk window `win`
    (or to the canvas when win is None) that the up-sampling blocks of G_res need for the crop at pixel offset yx -
    the crop's footprint plus TAIL_CONTEXT.  Same argument as `crop_window`: an edge of this window that is not an
    edge of the canvas lies >= TAIL_CONTEXT latent pixels from the footprint, and where it coincides with an interior
    edge of the trunk window the footprint is >= G_CONTEXT away from it by construction.  None = keep everything."""
    up = res // lat
    if up != 4:
        return None
    oy0, ox0, wh, ww = (0, 0, H, W) if win is None else win
    if wh <= TAIL_SIZE and ww <= TAIL_SIZE:
        return None

    def axis(c0, o0, L):
        if L <= TAIL_SIZE:
            return 0, L
        return min(max(c0 // up - o0 - TAIL_CONTEXT, 0), L - TAIL_SIZE), TAIL_SIZE
    oy, h = axis(yx[0], oy0, wh)
    ox, w = axis(yx[1], ox0, ww)
    return oy, ox, h, w


class Window(tuple):
    """(oy, ox, h, w) of a crop-aware window.  `dev` = optional int32 device view {oy, ox}: when set, the kernels
    read the offset from there instead of baking the host integers into the launch, so that the train step is the
    same sequence of launches whatever the random_crop offsets are and can replay as CUDA graphs (train.GraphedStep).
    The SIZE is always static."""

    def __new__(cls, oy, ox, h, w, dev=None):
        self = super().__new__(cls, (int(oy), int(ox), int(h), int(w)))
        self.dev = dev
        return self


def plan_crop(yx, res, lat, H, W, crop_aware=True, lod=0.0):
    """Everything one random_crop of a G_fcn image needs, as host integers: {'win': trunk window of the latent
    canvas, 'mid': window after the fourth residual block (relative to win), 'tail': window of the up-sampling
    blocks (relative to win + mid), 'img': (y0, x0, res, res) of the crop inside the decoded image}; windows are
    None where the whole extent is kept."""
    win = crop_window(yx, res, lat, H, W) if crop_aware else None
    mid = tail = None
    win_abs = win
    if crop_aware and lod <= 2.0:                  # windows stay aligned to the upscaled low-res pixels
        mid = mid_window(yx, res, lat, win, H, W)
        win_abs = compose_window(win, mid, H, W)
        tail = tail_window(yx, res, lat, win_abs, H, W)
    y0, x0 = image_offset(yx, res // lat, win_abs, tail)
    return dict(yx=(int(yx[0]), int(yx[1])), win=win, mid=mid, tail=tail, img=(y0, x0, res, res))


PLAN_KEYS = ('win', 'mid', 'tail', 'img')


def plan_offsets(plan):
    """The 8 int32 offsets of a plan in PLAN_KEYS order (0, 0 for an absent window)."""
    out = []
    for k in PLAN_KEYS:
        w = plan[k]
        out += [0, 0] if w is None else [int(w[0]), int(w[1])]
    return out


def plan_on_device(plan, dev_offsets):
    """The same plan with `Window`s whose offsets are read from `dev_offsets` (int32 device tensor of 8 values laid
    out by `plan_offsets`)."""
    out = dict(yx=plan['yx'])
    for i, k in enumerate(PLAN_KEYS):
        w = plan[k]
        out[k] = None if w is None else Window(w[0], w[1], w[2], w[3], dev=dev_offsets[2 * i:2 * i + 2])
    return out


def _win(rt, x, win):
    return x if win is None else rt.window(x, win)


ZG_MODES = ('hard', 'variational')
ZL_MODES = ('permutational', 'hard', 'variational', 'random')


def interp_modes(zg='hard', zl='permutational', noise=None):
    """zg_interp_variational / zl_interp_variational of loss.py:105-259, 351-521 (config.py:61-62: 'hard' /
    'permutational').  `noise`: the graph's tf.random_normal draws made explicit, standard-normal device tensors -
    'zg_f' [N,C,1,1] (zg 'variational'), 'zl_f' [N,C,H,W] (zl 'variational' | 'random': the draws of the
    non-corner tiles at their canvas positions), and 'zg_b' / 'zl_b' for the reversed branch of the blend."""
    if zg not in ZG_MODES or zl not in ZL_MODES:
        raise ValueError('interpolation modes: zg in %r, zl in %r (got %r, %r)' % (ZG_MODES, ZL_MODES, zg, zl))
    return dict(zg=zg, zl=zl, noise=dict(noise or {}))


def _noise(rt, mode, mu, ls, eps, H, W, reverse):
    n, c, sh, sw = mu.shape
    eh, ew = eps.shape[2:]
    out = rt.empty(n, c, H, W)
    _lib.check(rt.lib.tmx_latent_noise_fwd(rt.handle, mode, _ptr(mu.contiguous()), _ptr(None if ls is None else
                                                                                         ls.contiguous()),
                                           _ptr(eps.contiguous()), _ptr(out), n, c, sh, sw, eh, ew, H, W, int(reverse),
                                           rt.stream()), 'tmx_latent_noise_fwd')
    return out


def _noise_bwd(rt, mode, g, ls, eps, dmu, dls, reverse):
    n, c, sh, sw = dmu.shape
    eh, ew = eps.shape[2:]
    H, W = g.shape[2:]
    _lib.check(rt.lib.tmx_latent_noise_bwd(rt.handle, mode, _ptr(g.contiguous()), _ptr(None if ls is None else
                                                                                       ls.contiguous()),
                                           _ptr(eps.contiguous()), _ptr(dmu), _ptr(dls), n, c, sh, sw, eh, ew, H, W,
                                           int(reverse), rt.stream()), 'tmx_latent_noise_bwd')


def _zg_canvas(rt, zg_mu, zg_ls, wh, ww, modes, which, reverse):
    """The global-code canvas (window-sized: a 1x1 code tiles any extent) of the forward ('f') / reversed ('b') branch."""
    if modes is None or modes['zg'] == 'hard':                            # loss.py:176, 218
        return rt.latent_blend([zg_mu.contiguous()], wh, ww, _lib.BLEND_COPY, src_reverse=int(reverse))
    return _noise(rt, 1, zg_mu, zg_ls, modes['noise']['zg_' + which], wh, ww, reverse)      # loss.py:178-180, 220-222


def _zl_canvas(rt, zl_mu, zl_ls, H, W, pins, ih, iw, win, modes, which, reverse):
    mode = 'permutational' if modes is None else modes['zl']
    if mode == 'permutational':                                           # loss.py:194, 236
        full = rt.latent_blend([zl_mu.contiguous()], H, W, _lib.BLEND_COPY, idx_h=[ih], idx_w=[iw], pin_rows=pins[0],
                               pin_cols=pins[1], src_reverse=int(reverse))
    elif mode == 'hard':                                                  # tf.tile, loss.py:182, 224
        n = zl_mu.shape[0]
        ident = lambda L: torch.arange(L, dtype=torch.int32, device=rt.device).repeat(n, 1).contiguous()   # noqa: E731
        full = rt.latent_blend([zl_mu.contiguous()], H, W, _lib.BLEND_COPY, idx_h=[ident(H)], idx_w=[ident(W)],
                               src_reverse=int(reverse))
    else:                                                                 # loss.py:183-193, 225-235
        full = _noise(rt, 1 if mode == 'variational' else 2, zl_mu, zl_ls, modes['noise']['zl_' + which], H, W, reverse)
    return _win(rt, full, win)


def fcn_canvases(rt, zg_mu, zl_mu, H, W, pins, ih_f, iw_f, win, blend=None, modes=None, zg_ls=None, zl_ls=None):
    """The latent canvases G_fcn decodes (loss.py:176-194 interpolation; :218-239 blend when `blend` =
    (ih_b, iw_b, t)), restricted to `win`; `modes` = interp_modes(...) for the config-off variants."""
    wh, ww = (H, W) if win is None else win[2:]
    zg_c = _zg_canvas(rt, zg_mu, zg_ls, wh, ww, modes, 'f', False)
    zl_c = _zl_canvas(rt, zl_mu, zl_ls, H, W, pins, ih_f, iw_f, win, modes, 'f', False)
    if blend is None:
        return zg_c, zl_c
    ih_b, iw_b, t = blend
    # tf.reverse(axis=[0]) of the sources is folded into the gather (src_reverse)
    zg_r = _zg_canvas(rt, zg_mu, zg_ls, wh, ww, modes, 'b', True)
    zl_r = _zl_canvas(rt, zl_mu, zl_ls, H, W, pins, ih_b, iw_b, win, modes, 'b', True)
    bzg = rt.latent_blend([zg_r, zg_c], wh, ww, _lib.BLEND_LERP, t=t)      # lerp(reverse, forward, t), loss.py:238
    bzl = rt.latent_blend([zl_r, zl_c], wh, ww, _lib.BLEND_LERP, t=t)
    return bzg, bzl


def image_offset(yx, up, win, tail):
    """Pixel offset of the crop at canvas offset yx inside the image decoded from trunk window `win` and tail
    window `tail` (either may be None)."""
    oy = (0 if win is None else win[0]) + (0 if tail is None else tail[0])
    ox = (0 if win is None else win[1]) + (0 if tail is None else tail[1])
    return yx[0] - up * oy, yx[1] - up * ox


def fcn_scale(canvas, lat):
    """G_fcn's (scale_h, scale_w) for a canvas or a window of one: the same variables decode any whole number of
    tiles (the reference builds one G_fcn per scale over the shared scope, run.py:273, util_scripts.py:380-384)."""
    return dict(scale_h=canvas.shape[2] // lat, scale_w=canvas.shape[3] // lat)


def _crop_adjoint(rt, dcrop, full_hw, img_win):
    """Adjoint of random_crop (loss.py:78-90): zeros outside the window (pure data movement)."""
    return rt.window_embed(dcrop, img_win, full_hw[0], full_hw[1])


class EGForward:
    """The E/G side of `EG_wgan` run once with tapes: encoders, reconstruction, interpolated and blended canvases.
    Its images depend only on the E/G variables, which do not change between the critic update and the E/G update of
    one step (run.py:511-512), so a critic phase on the SAME reals can reuse `rec` as its fake.
    With `record=False` nothing is taped (the critics' fakes, loss.py:308-320, when their minibatch is not the E/G
    phase's)."""

    def __init__(self, E_zg, E_zl, G, G_fcn, reals, idx, mixing_factors, scale_h=3, scale_w=3, need_interp=True,
                 need_blend=True, crop_interp=None, crop_blend=None, defer_canvases=False, plans=None, record=True,
                 modes=None):
        """crop_interp / crop_blend: the (y, x) offsets the E/G loss will crop at; when given, G_fcn decodes only the
        latent window those crops depend on (`crop_window`) - same crop pixels, same gradients, 44 % of the work at
        the reference's 3x3 canvases.  None decodes the whole canvas.  `plans` = {'interp' | 'blend': plan_crop(...)}
        supplies the windows ready-made (the trainer's: their offsets live in device memory)."""
        rt = self.rt = Runtime.get(reals.device)
        self.nets = (E_zg, E_zl, G, G_fcn)
        self.modes = modes               # interp_modes(...) or None = the reference config ('hard', 'permutational')
        self.reals, self.scale = reals, (scale_h, scale_w)
        self.n = reals.shape[0]
        res = reals.shape[2]
        tape = (lambda: []) if record else (lambda: None)
        self.t_zg, self.t_zl, self.t_rec, self.t_int, self.t_bl = tape(), tape(), tape(), tape(), tape()
        self.zg_mu, self.zg_ls = E_zg.get_output_for(reals, tape=self.t_zg)
        self.zl_mu, self.zl_ls = E_zl.get_output_for(reals, tape=self.t_zl)
        _sink('E_zg', E_zg, self.t_zg)
        _sink('E_zl', E_zl, self.t_zl)
        zg_mu, zl_mu = self.zg_mu, self.zl_mu
        self.c, self.lat = zl_mu.shape[1], zl_mu.shape[2]
        lat = self.lat
        self.H, self.W = H, W = lat * scale_h, lat * scale_w
        self.pins = interp._corner_pins(scale_h, scale_w)
        self.rec = G.get_output_for(_tile_code(rt, zg_mu, lat, lat), zl_mu, tape=self.t_rec)
        _sink('G_rec', G, self.t_rec)
        self.interp = self.blend = None
        plans = dict(plans or {})
        for k, yx in (('interp', crop_interp), ('blend', crop_blend)):
            if k not in plans:
                plans[k] = plan_crop((0, 0) if yx is None else yx, res, lat, H, W, crop_aware=yx is not None,
                                     lod=G_fcn.lod)
                if yx is None:
                    plans[k]['yx'] = None        # whole canvas decoded: any crop may be taken from it
        self.plans = plans
        self.win = {k: plans[k]['win'] for k in plans}
        self.mid = {k: plans[k]['mid'] for k in plans}
        self.tail = {k: plans[k]['tail'] for k in plans}
        self._need = (need_interp, need_blend)
        if need_interp or need_blend:
            self.ih_f, self.iw_f = _dev_idx(rt, idx['h_forward']), _dev_idx(rt, idx['w_forward'])
            self.ih_b, self.iw_b = _dev_idx(rt, idx['h_backward']), _dev_idx(rt, idx['w_backward'])
        if need_blend:
            self.t = mixing_factors.reshape(-1).contiguous()
        if not defer_canvases:
            self.decode_canvases()

    def decode_canvases(self):
        """The two taped G_fcn evaluations (interpolated and blended canvases, loss.py:176-246).  Separate from the
        constructor so that the trainer can start the critic phase - which only needs the encoders' codes and the
        reconstruction - before them and let the two overlap."""
        rt, (E_zg, E_zl, G, G_fcn) = self.rt, self.nets
        zg_mu, zl_mu, H, W, pins, lat = self.zg_mu, self.zl_mu, self.H, self.W, self.pins, self.lat
        need_interp, need_blend = self._need
        if need_interp and self.interp is None:
            zg_c, zl_c = fcn_canvases(rt, zg_mu, zl_mu, H, W, pins, self.ih_f, self.iw_f, self.win['interp'],
                                      modes=self.modes, zg_ls=self.zg_ls, zl_ls=self.zl_ls)
            self.interp = G_fcn.get_output_for(zg_c, zl_c, tape=self.t_int, mid_window=self.mid['interp'],
                                               tail_window=self.tail['interp'], **fcn_scale(zl_c, lat))
            _sink('G_interp', G_fcn, self.t_int)
        if need_blend and self.blend is None:
            bzg, bzl = fcn_canvases(rt, zg_mu, zl_mu, H, W, pins, self.ih_f, self.iw_f, self.win['blend'],
                                    blend=(self.ih_b, self.iw_b, self.t), modes=self.modes, zg_ls=self.zg_ls,
                                    zl_ls=self.zl_ls)
            self.blend = G_fcn.get_output_for(bzg, bzl, tape=self.t_bl, mid_window=self.mid['blend'],
                                              tail_window=self.tail['blend'], **fcn_scale(bzl, lat))
            _sink('G_blend', G_fcn, self.t_bl)

    def image_window(self, which, yx):
        """(y0, x0, res, res) of crop `yx` inside the decoded image of `which` (== yx when the whole canvas was
        decoded); carries the device offset when the plan does."""
        plan = self.plans[which]
        res = self.reals.shape[2]
        if plan['yx'] is None:
            return Window(yx[0], yx[1], res, res)
        if tuple(yx) != tuple(plan['yx']):
            raise ValueError('%s was decoded for the crop at %r only (crop-aware G_fcn); asked for %r'
                             % (which, plan['yx'], yx))
        return plan['img']

    def window_offset(self, which, yx):
        """Pixel offset (host integers) of crop `yx` inside the decoded image of `which`."""
        return tuple(self.image_window(which, yx)[:2])

    def crop(self, which, yx):
        img = self.interp if which == 'interp' else self.blend
        return self.rt.window(img, self.image_window(which, yx))


def fcn_fake(G_fcn, fwd, which, yx, mix=None, crop_aware=True, plan=None, modes=None):
    """Fake images of the canvas critics (no tape): the crop at `yx` of G_fcn's image of the interpolated (`which` =
    'interp', loss.py:391-395) or blended ('blend', loss.py:466-495) canvas built from the codes of `fwd`, decoding
    only the latent window the crop depends on (`crop_window`).  D_blend_wgangp draws its own mixing factors
    (loss.py:489), D_interp_wgangp its own crop offset (loss.py:395): neither can reuse the E/G images once those
    are decoded crop-aware.  `plan` = plan_crop(yx, ...) ready-made (device-resident offsets)."""
    rt = fwd.rt
    res = fwd.reals.shape[2]
    if plan is None:
        plan = plan_crop(yx, res, fwd.lat, fwd.H, fwd.W, crop_aware=crop_aware, lod=G_fcn.lod)
    blend = None if which == 'interp' else (fwd.ih_b, fwd.iw_b, mix.reshape(-1).contiguous())
    zg_c, zl_c = fcn_canvases(rt, fwd.zg_mu, fwd.zl_mu, fwd.H, fwd.W, fwd.pins, fwd.ih_f, fwd.iw_f, plan['win'], blend,
                              modes=modes if modes is not None else fwd.modes, zg_ls=fwd.zg_ls, zl_ls=fwd.zl_ls)
    img = G_fcn.get_output_for(zg_c, zl_c, mid_window=plan['mid'], tail_window=plan['tail'],
                               **fcn_scale(zl_c, fwd.lat))
    return rt.window(img, plan['img'])


def critic_input_gradient(D, images, weight):
    """One adversarial term of `EG_wgan` (loss.py:135-136, 194-199, 241-246): the critic as a FIXED function.
    -> (mean over the batch of -weight * D(images) [device scalar], its gradient w.r.t. images [N,3,R,R])."""
    rt = D.rt
    n = images.shape[0]
    tape = []
    s = D.get_output_for(images, tape=tape)
    _sink('critic_fixed', D, tape)
    term = _row_sum(rt, s, 1, n, scale=-weight / n)
    (d,) = backward(D, tape, [torch.full_like(s, -weight / n)], None, param_grads=False)
    return term, d


def _zg_canvas_bwd(rt, fwd, d, dzg, lsg, which, reverse):
    """Adjoint of `_zg_canvas`: d [N,C,wh,ww] -> dzg [N*C] (+ lsg['zg'] for the 'variational' mode)."""
    n, c = fwd.n, fwd.c
    wh, ww = d.shape[2:]
    modes = fwd.modes
    if modes is None or modes['zg'] == 'hard':
        if not reverse:
            _row_sum(rt, d, n * c, wh * ww, out=dzg, accumulate=True)
        else:
            tmp = _row_sum(rt, d, n * c, wh * ww).view(n, c, 1, 1)
            _gather_bwd(rt, tmp, dzg.view(n, c, 1, 1), None, None, (0, 0), reverse=True)
        return
    if 'zg' not in lsg:
        lsg['zg'] = torch.zeros_like(fwd.zg_ls)
    _noise_bwd(rt, 1, d, fwd.zg_ls, modes['noise']['zg_' + which], dzg.view(n, c, 1, 1), lsg['zg'], reverse)


def _zl_canvas_bwd(rt, fwd, d, win, dzl, lsg, ih, iw, which, reverse):
    """Adjoint of `_zl_canvas`: d = gradient w.r.t. the window `win` of the canvas -> dzl (+ lsg['zl'])."""
    H, W, pins = fwd.H, fwd.W, fwd.pins
    mode = 'permutational' if fwd.modes is None else fwd.modes['zl']
    if mode == 'permutational':
        _gather_bwd_win(rt, d, win, H, W, dzl, ih, iw, pins, reverse=reverse)
    elif mode == 'hard':
        _gather_bwd_win(rt, d, win, H, W, dzl, None, None, (0, 0), reverse=reverse)
    else:
        g = d if win is None else rt.window_embed(d, win, H, W)
        if mode == 'variational' and 'zl' not in lsg:
            lsg['zl'] = torch.zeros_like(fwd.zl_ls)
        _noise_bwd(rt, 1 if mode == 'variational' else 2, g, fwd.zl_ls, fwd.modes['noise']['zl_' + which], dzl,
                   lsg.get('zl'), reverse)


def gram_terms(fwd, crop_interp, crop_blend, gram, gram_weight, gram_alpha, interp_G_weight=1.0,
               blend_interp_G_weight=1.0, reals_fade=None):
    """The VGG-19 Gram terms of EG_wgan (loss.py:148-160, 206-213, 248-257) and their image gradients:
    {'rec' | 'interp' | 'blend': (term, d term / d image)}.  They depend on the generated images and the reals only -
    not on the critics - so the trainer evaluates them while the critics' gradient all-reduce is in flight."""
    rt = fwd.rt
    reals = fwd.reals if reals_fade is None else reals_fade
    n = fwd.n
    out = {}
    _, real_gram = gram.grams(reals.contiguous())
    out['rec'] = gram.term(fwd.rec, [(real_gram, False, None, 0)], gram_weight)
    if interp_G_weight > 0:
        out['interp'] = gram.term(fwd.crop('interp', crop_interp), [(real_gram, False, None, 0)], gram_weight)
    if blend_interp_G_weight > 0:
        # loss.py:252-255 AS WRITTEN: (1 - alpha) [N,1,1,1] * multi_layer_diff [N] broadcasts to [N,1,1,N], so what
        # the optimizer differentiates is mean(1 - alpha) * mean_j A_j + mean(alpha) * mean_j B_j, A against the
        # batch-reversed real Gram matrices, B against the real ones (kept as the reference has it)
        abar = _row_sum(rt, gram_alpha.reshape(-1).contiguous(), 1, n, scale=1.0 / n)
        out['blend'] = gram.term(fwd.crop('blend', crop_blend),
                                 [(real_gram, True, abar, 2), (real_gram, False, abar, 1)], gram_weight)
    return out


def EG_backward(fwd, D_rec, D_interp, D_blend, crop_interp, crop_blend, grads, rec_G_weight=1.0, pixel_weight=200.0,
                interp_G_weight=1.0, blend_interp_G_weight=1.0, reals_fade=None, critic_grads=None, kl_weight=0.0,
                gram=None, gram_weight=0.0, gram_alpha=None, gram_grads=None):
    """Loss terms of `EG_wgan` on the images of `fwd` and the whole reverse pass into `grads`.  `reals_fade`: the
    target of the pixel loss (loss.py:143) when it differs from what the encoders saw (fractional lod).
    `critic_grads`: optional {'rec' | 'interp' | 'blend': critic_input_gradient(...)} evaluated by the caller (the
    trainer runs them on parallel streams).
    `gram` (a vgg.GramLoss) with gram_weight > 0 adds the VGG-19 Gram terms of loss.py:148-160, 206-213, 248-257;
    `gram_alpha` = the [N,1,1,1] uniform draw of loss.py:253; `gram_grads` = gram_terms(...) evaluated by the caller."""
    rt = fwd.rt
    E_zg, E_zl, G, G_fcn = fwd.nets
    reals, n, c, lat, H, W, pins = fwd.reals, fwd.n, fwd.c, fwd.lat, fwd.H, fwd.W, fwd.pins
    if reals_fade is not None:
        reals = reals_fade
    inv_n = 1.0 / n
    report = {}
    rec = fwd.rec
    d_rec_img = None
    cg = critic_grads or {}
    if rec_G_weight > 0:
        report['rec_G'], d_rec_img = cg['rec'] if 'rec' in cg else critic_input_gradient(D_rec, rec, rec_G_weight)
    if pixel_weight > 0:
        l1 = rt.empty(*rec.shape)
        lsum = torch.zeros(1, dtype=torch.float32, device=rt.device)
        per = rec[0].numel()
        _lib.check(rt.lib.tmx_loss_l1_grad(rt.handle, _ptr(rec), _ptr(reals.contiguous()), _ptr(l1), _ptr(lsum),
                                           rec.numel(), pixel_weight * inv_n / per, rt.stream()), 'tmx_loss_l1_grad')
        report['rec_pixel'] = _row_sum(rt, lsum, 1, 1, scale=pixel_weight * inv_n / per)
        d_rec_img = l1 if d_rec_img is None else _add(rt, d_rec_img, l1)
    use_gram = gram is not None and gram_weight > 0
    if use_gram:                                                          # loss.py:149-160
        gt = gram_grads if gram_grads is not None else gram_terms(
            fwd, crop_interp, crop_blend, gram, gram_weight, gram_alpha, interp_G_weight, blend_interp_G_weight,
            reals_fade=reals_fade)
        report['rec_gram'], dg = gt['rec']
        d_rec_img = dg if d_rec_img is None else _add(rt, d_rec_img, dg)
    dzg_tiled, dzl = backward(G, fwd.t_rec, [d_rec_img], grads['G'])
    dzg = _row_sum(rt, dzg_tiled, n * c, lat * lat)                      # adjoint of the 32x32 tile of zg
    dzl = dzl.contiguous()
    lsg = {}                             # gradients w.r.t. the encoders' log_sigma outputs ('variational' modes only)
    if interp_G_weight > 0:
        report['interp_G'], dcr = cg['interp'] if 'interp' in cg else \
            critic_input_gradient(D_interp, fwd.crop('interp', crop_interp), interp_G_weight)
        if use_gram:                                                      # loss.py:206-213
            report['interp_gram'], dg = gt['interp']
            dcr = _add(rt, dcr, dg)
        dzg_c, dzl_c = backward(G_fcn, fwd.t_int, [_crop_adjoint(rt, dcr, fwd.interp.shape[2:],
                                                                 fwd.image_window('interp', crop_interp))], grads['G'])
        _zg_canvas_bwd(rt, fwd, dzg_c, dzg, lsg, 'f', False)
        _zl_canvas_bwd(rt, fwd, dzl_c, fwd.win['interp'], dzl, lsg, fwd.ih_f, fwd.iw_f, 'f', False)
    if blend_interp_G_weight > 0:
        t = fwd.t
        report['blend_G'], dcr = cg['blend'] if 'blend' in cg else \
            critic_input_gradient(D_blend, fwd.crop('blend', crop_blend), blend_interp_G_weight)
        if use_gram:
            report['blend_gram'], dg = gt['blend']
            dcr = _add(rt, dcr, dg)
        dbzg, dbzl = backward(G_fcn, fwd.t_bl, [_crop_adjoint(rt, dcr, fwd.blend.shape[2:],
                                                              fwd.image_window('blend', crop_blend))], grads['G'])
        zero_c = torch.zeros_like(dbzg)
        win = fwd.win['blend']
        wh, ww = dbzg.shape[2:]
        # adjoint of lerp: d forward = t * d, d reverse = d - t * d   (same fp32 ops as autograd of a + (b - a) * t)
        for d, dsrc_kind in ((dbzg, 'zg'), (dbzl, 'zl')):
            d_fwd = rt.latent_blend([zero_c, d.contiguous()], wh, ww, _lib.BLEND_LERP, t=t)
            d_rev = rt.latent_blend([d.contiguous(), zero_c], wh, ww, _lib.BLEND_LERP, t=t)
            if dsrc_kind == 'zg':
                _zg_canvas_bwd(rt, fwd, d_fwd, dzg, lsg, 'f', False)
                _zg_canvas_bwd(rt, fwd, d_rev, dzg, lsg, 'b', True)
            else:
                _zl_canvas_bwd(rt, fwd, d_fwd, win, dzl, lsg, fwd.ih_f, fwd.iw_f, 'f', False)
                _zl_canvas_bwd(rt, fwd, d_rev, win, dzl, lsg, fwd.ih_b, fwd.iw_b, 'b', True)
    dzg_ls, dzl_ls = lsg.get('zg'), lsg.get('zl')       # d loss / d log_sigma from the 'variational' canvases, if any
    dzg = dzg.view(n, c, 1, 1)
    if kl_weight > 0:                                                     # loss.py:163-171
        (report['KL_zg'], dzg, kg) = _kl(rt, fwd.zg_mu, fwd.zg_ls, dzg, kl_weight, 'KL_zg')
        (report['KL_zl'], dzl, kl_) = _kl(rt, fwd.zl_mu, fwd.zl_ls, dzl, kl_weight, 'KL_zl')
        dzg_ls = kg if dzg_ls is None else _add(rt, dzg_ls, kg)
        dzl_ls = kl_ if dzl_ls is None else _add(rt, dzl_ls, kl_)
    backward(E_zl, fwd.t_zl, [dzl, dzl_ls], grads['E_zl'], want_input_grads=False)
    backward(E_zg, fwd.t_zg, [dzg, dzg_ls], grads['E_zg'], want_input_grads=False)
    return report


def _kl(rt, mu, ls, dmu_in, kl_weight, name):
    """KL term of one encoder: (batch mean of the term, dL/dmu incl. the incoming gradient, dL/dlog_sigma)."""
    mu, ls = mu.contiguous(), ls.contiguous()
    total = mu.numel()
    dmu, dls, val = rt.empty(*mu.shape), rt.empty(*mu.shape), rt.empty(total)
    _lib.check(rt.lib.tmx_kl_terms(rt.handle, _ptr(mu), _ptr(ls), _ptr(dmu), _ptr(dls), _ptr(val), total,
                                   float(kl_weight) / total, rt.stream()), 'tmx_kl_terms')
    term = _row_sum(rt, val, 1, total, scale=-0.5 * float(kl_weight) / total)
    return term, _add(rt, dmu_in.contiguous().view(*mu.shape), dmu), dls


def EG_wgan(E_zg, E_zl, G, D_rec, G_fcn, D_interp, D_blend, reals, idx, crop_interp, crop_blend, mixing_factors,
            grads, scale_h=3, scale_w=3, rec_G_weight=1.0, pixel_weight=200.0, interp_G_weight=1.0,
            blend_interp_G_weight=1.0, crop_aware=True, kl_weight=0.0, gram=None, gram_weight=0.0, gram_alpha=None,
            modes=None):
    """One evaluation + differentiation of mean(EG_loss) (loss.py:105-259, run.py:321).
    reals: [N,3,R,R] fp32 device tensor in [-1,1]; idx: dict of int32 index vectors (interp.sample_schedule_indices);
    crop_*: (y, x); mixing_factors: [N,1,1,1] fp32 device tensor; grads: {'E_zg','E_zl','G'} -> flat gradient
    buffers (accumulated into).  Returns a dict of per-term batch means (device scalars)."""
    fwd = EGForward(E_zg, E_zl, G, G_fcn, reals, idx, mixing_factors, scale_h, scale_w,
                    need_interp=interp_G_weight > 0, need_blend=blend_interp_G_weight > 0,
                    crop_interp=crop_interp if crop_aware else None, crop_blend=crop_blend if crop_aware else None,
                    modes=modes)
    return EG_backward(fwd, D_rec, D_interp, D_blend, crop_interp, crop_blend, grads, rec_G_weight, pixel_weight,
                       interp_G_weight, blend_interp_G_weight, kl_weight=kl_weight, gram=gram, gram_weight=gram_weight,
                       gram_alpha=gram_alpha)


def _add(rt, a, b):
    out = rt.empty(*a.shape)
    _lib.check(rt.lib.tmx_add_f32(rt.handle, _ptr(a.contiguous()), _ptr(b.contiguous()), _ptr(out), a.numel(),
                                  rt.stream()), 'tmx_add_f32')
    return out


# ====================================================================== critic phase (WGAN-GP)
def _mask(rt, t, y, n, h, w, c):
    """t * lrelu'(y): the activation of the mask-frozen (linearised) network.  y: fp32 map or an Act."""
    from .backward import _mask_of
    kw = _mask_of(rt, y) if isinstance(y, Act) else dict(y_f32=y)
    _, out = rt.grad_prepare(t, n, h, w, c, src_kind=1, want_planes=False, want_f32=True, **kw)
    return out


def tangent_forward(D, tape, v):
    """Forward of the linearised critic (no biases, leaky-ReLU branches frozen at the recorded evaluation) on the
    tangent image v [N,3,R,R]: J_D(x) v.  Returns {tape position: tangent INPUT of that parametrised layer} and the
    (primal, tangent) pair at the minibatch-stddev layer."""
    rt = D.rt
    tang, tin, mb = {}, {}, None
    timg = {}          # tangents of the pooled copies of the input image (progressive growing, lod > 0)
    for pos, rec in enumerate(tape[:-1]):
        kind = rec['kind']
        if kind == 'imgpool':
            from .networks import _pool_image
            timg[id(rec['y'])] = _pool_image(rt, timg.get(id(rec['x']), v), rec['factor'])
        elif kind == 'lerp':
            from .networks import _lerp_lod
            ta, tb = rt.split_unpack(tang[id(rec['a'])]), rt.split_unpack(tang[id(rec['b'])])
            y = rec['y']
            tang[id(y)] = Act(y.n, y.h, y.w, y.c, f32=_lerp_lod(rt, ta.f32, tb.f32, rec['t']))
        elif kind == 'fromrgb':
            y = rec['y']
            vin = timg.get(id(rec['img']), v)
            t = rt.fromrgb(vin, D.vars[rec['w']].value, None, rec['wscale'], rec['cout'], lrelu=False)
            t.f32 = _mask(rt, t.f32, y.f32, y.n, y.h, y.w, y.c)
            tin[pos] = vin
            tang[id(y)] = t
        elif kind == 'conv':
            xin, y = tang[id(rec['x'])], rec['y']
            wv = D.vars[rec['w']]
            prepared = D.prepared_weights(wv, rec['wscale'], rec['k'], rec['cin'], rec['cout'], cin_pad=xin.c)
            if rec['up2']:
                raise NotImplementedError('tangent of an upsampling conv (not part of D_patch)')
            out = rt.conv2d(xin, wv.value, None, rec['wscale'], rec['k'], rec['cout'], lrelu=False, want_f32=True,
                            want_split=False, algo=_lib.ALGO_TC, prepared=prepared,
                            halo_in='zero' if rec.get('halo') == 'zero' else None)
            if rec['act']:
                out.f32 = _mask(rt, out.f32, y, y.n, y.h, y.w, y.c)
            tin[pos] = xin                      # carries the padded planes (REFLECT / ZERO) the conv just packed
            tang[id(y)] = out
        elif kind == 'bias_act':                # fused conv2d_downscale2d: the bias drops out, the mask stays
            t, y = rt.split_unpack(tang[id(rec['x'])]), rec['y']
            f = _mask(rt, t.f32, y.f32, y.n, y.h, y.w, y.c) if rec['act'] else t.f32
            tang[id(y)] = Act(y.n, y.h, y.w, y.c, f32=f)
        elif kind == 'pool':
            tang[id(rec['y'])] = rt.avgpool2(tang[id(rec['x'])])
        elif kind == 'mbstd':
            x, y, xd = rec['x'], rec['y'], rt.split_unpack(tang[id(rec['x'])])
            g = min(rec['group'], x.n)
            yd = Act(y.n, y.h, y.w, y.c, f32=rt.empty(y.n, y.h, y.w, y.c))
            sdot = rt.empty(x.n // g)
            _lib.check(rt.lib.tmx_mbstd_tangent(rt.handle, _ptr(rt.split_unpack(x).f32), _ptr(xd.f32), _ptr(yd.f32),
                                                _ptr(sdot), x.n, x.h, x.w, x.c, y.c, rec['group'], rt.stream()),
                       'tmx_mbstd_tangent')
            tang[id(y)] = yd
            mb = (pos, x, xd)
        elif kind == 'flatten':
            a = rt.split_unpack(tang[id(rec['x'])])
            tang[id(rec['y'])] = rt.nhwc_to_nchw(a.f32)
        elif kind == 'dense':
            xin, y = tang[id(rec['x'])], rec['y']
            n = y.shape[0]
            out = rt.dense(xin.view(n, -1), D.vars[rec['w']].value, None, rec['wscale'], lrelu=False)
            if rec['act']:
                out = _mask(rt, out, y, n, 1, 1, y.shape[1])
            tin[pos] = xin
            tang[id(y)] = out
        else:
            raise NotImplementedError('tangent of tape record %r' % kind)
    return tin, mb


def gradient_penalty(D, mixed, flat_grad, wgan_lambda=10.0, wgan_target=1.0):
    """mean_n lambda (||grad_x D(x_n)|| - target)^2 / target^2 (loss.py:327-337) and its gradient w.r.t. D's
    variables, accumulated into flat_grad.  Double backward without autograd:
      1. forward at `mixed` (tape) and first backward of sum(scores) w.r.t. the images, keeping every layer's adjoint
      2. per-sample coefficient c_n = d mean(penalty) / d||g_n|| / ||g_n||; tangent seed v_n = c_n g_n
      3. tangent forward of the mask-frozen network on v
      4. weight gradients = adjoint (step 1) x tangent input (step 3) for every conv / dense / FromRGB layer
      5. curvature of the minibatch-stddev statistic -> an extra fp32 gradient at its input, pushed through an
         ordinary backward of the layers below it."""
    rt = D.rt
    n = mixed.shape[0]
    tape, adj = [], {}
    s = D.get_output_for(mixed, tape=tape)
    _sink('critic_mixed', D, tape)
    (g,) = backward(D, tape, [torch.ones_like(s)], None, param_grads=False, adjoints=adj)
    per = g[0].numel()
    sq = _row_sum(rt, g, n, per, square=True)
    pen, coef = rt.empty(n), rt.empty(n)
    _lib.check(rt.lib.tmx_gp_coefficients(rt.handle, _ptr(sq), _ptr(pen), _ptr(coef), n, float(wgan_lambda),
                                          float(wgan_target), rt.stream()), 'tmx_gp_coefficients')
    v = rt.empty(*g.shape)
    _lib.check(rt.lib.tmx_scale_rows(rt.handle, _ptr(g), _ptr(coef), _ptr(v), n, per, rt.stream()), 'tmx_scale_rows')
    tin, mb = tangent_forward(D, tape, v)
    for pos, rec in enumerate(tape[:-1]):
        if pos not in adj or pos not in tin:
            continue
        kind = rec['kind']
        if kind == 'conv':
            x = tin[pos]
            conv_wgrad_into(rt, D, dict(rec, x=x), (x.hi, x.lo), adj[pos], D.grad_view(flat_grad, rec['w']))
        elif kind == 'fromrgb':
            img = tin[pos]
            nn, cimg, h, w = img.shape
            _lib.check(rt.lib.tmx_fromrgb_bwd(rt.handle, _ptr(img), _ptr(adj[pos]), _ptr(D.vars[rec['w']].value),
                                              float(rec['wscale']), _ptr(D.grad_view(flat_grad, rec['w'])), None, nn,
                                              cimg, h, w, rec['cout'], rt.stream()), 'tmx_fromrgb_bwd')
        elif kind == 'dense':
            dy, y = adj[pos]
            xin = tin[pos]
            _lib.check(rt.lib.tmx_dense_wgrad(rt.handle, _ptr(xin), _ptr(dy), _ptr(y),
                                              _ptr(D.grad_view(flat_grad, rec['w'])), None, y.shape[0],
                                              xin.numel() // y.shape[0], y.shape[1], float(rec['wscale']),
                                              int(rec['act']), runtime.LRELU_ALPHA, rt.stream()), 'tmx_dense_wgrad')
    if mb is not None:
        pos, x, xd = mb
        q = rt.empty(x.n, x.h, x.w, x.c)
        _lib.check(rt.lib.tmx_mbstd_curvature(rt.handle, _ptr(x.f32), _ptr(xd.f32), _ptr(adj[pos]), _ptr(q), x.n, x.h,
                                              x.w, x.c, tape[pos]['group'], rt.stream()), 'tmx_mbstd_curvature')
        backward(D, tape, [None], flat_grad, want_input_grads=False, param_grads=True, seeds=[(x, q)])
    return _row_sum(rt, pen, 1, n, scale=1.0 / n)


def D_wgangp(D, fakes, reals, mixing_factors, flat_grad, wgan_lambda=10.0, wgan_epsilon=0.001, wgan_target=1.0):
    """The critic loss shared by D_rec_wgangp / D_interp_wgangp / D_blend_wgangp (loss.py:303-521); they differ
    only in how `fakes` is produced (reconstruction, interpolation crop, blend crop):
        mean(D(fake) - D(real)) + lambda (||grad D(mixed)|| - 1)^2 + eps D(real)^2,  mixed = lerp(real, fake, t).
    Differentiates mean over the batch w.r.t. D's variables into flat_grad; returns the term means."""
    rt = D.rt
    n = reals.shape[0]
    rep = {}
    t_f, t_r = [], []
    s_f = D.get_output_for(fakes.contiguous(), tape=t_f)
    _sink('critic_fake', D, t_f)
    backward(D, t_f, [torch.full_like(s_f, 1.0 / n)], flat_grad, want_input_grads=False)
    s_r = D.get_output_for(reals.contiguous(), tape=t_r)
    _sink('critic_real', D, t_r)
    seed = rt.empty(*s_r.shape)                       # d/ds_real [ -s/N + eps s^2 / N ]
    _lib.check(rt.lib.tmx_axpb(rt.handle, _ptr(s_r), _ptr(seed), n, 2.0 * wgan_epsilon / n, -1.0 / n, rt.stream()),
               'tmx_axpb')
    backward(D, t_r, [seed], flat_grad, want_input_grads=False)
    rep['D_loss'] = _add(rt, _row_sum(rt, s_f, 1, n, scale=1.0 / n), _row_sum(rt, s_r, 1, n, scale=-1.0 / n))
    rep['epsilon_penalty'] = _row_sum(rt, s_r, 1, n, scale=wgan_epsilon / n, square=True)
    per = reals[0].numel()
    mixed = rt.latent_blend([reals.contiguous().view(n, 1, 1, per), fakes.contiguous().view(n, 1, 1, per)], 1, per,
                            _lib.BLEND_LERP, t=mixing_factors.reshape(-1).contiguous()).view(*reals.shape)
    rep['gradient_penalty'] = gradient_penalty(D, mixed, flat_grad, wgan_lambda, wgan_target)
    return rep
