"""Reverse pass over a forward tape: the hand-written counterpart of `tf.gradients`
(tfutil.py:299, loss.py:285,333) for the networks of networks.py.  No autograd:
`Network.get_output_for(..., tape=[])` records one entry per layer (saved input
planes, fp32 output, variable names); `backward()` walks it backwards and launches

    tmx_grad_prepare   padding / pooling adjoints, residual add, leaky-ReLU mask, bias gradient
    tmx_conv2d_wgrad   weight gradient (tcgen05, MN-major operands)
    tmx_conv2d_dgrad   data gradient   (tcgen05, LIN mode)
    tmx_torgb_bwd / tmx_fromrgb_bwd    the 1x1 RGB heads

Weight/bias gradients are ACCUMULATED into `flat_grad`, a fp32 buffer laid out like
`net.flat` (what Optimizer.register_gradients takes).  A gradient with respect to an
activation travels in one of three forms:
    ('f32', t)            NHWC fp32 at the activation's resolution
    ('grid', g, fold)     a consumer's dgrad output on the zero-ringed grid, to be folded
                          (0 REFLECT, 1 REPLICATE [consumer read through upscale2d], 2 none)
    ('pool', t)           NHWC fp32 at half resolution (consumer was downscale2d)
    ('ready', planes, t)  already combined, masked and split by the consumer's fused data gradient
                          (tmx_conv2d_dgrad_gp): dz planes on the zero-ringed grid (+ fp32 NHWC when needed)"""
import ctypes as C
import os

import torch

from . import _lib, runtime


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Grads:
    """Pending gradient contributions per activation (keyed by object identity)."""

    def __init__(self):
        self.by_id = {}

    def add(self, act, contrib):
        self.by_id.setdefault(id(act), []).append(contrib)

    def pop(self, act):
        return self.by_id.pop(id(act), [])


def _mask_of(rt, y):
    """What the leaky-ReLU mask of activation `y` is read from: the bf16 `hi` plane of its split form when the forward
    kept one at the activation's own resolution (sign(hi) == sign(y): bf16 shares fp32's exponent range; half the
    bytes, and the training forward need not write the fp32 map at all), else the fp32 map."""
    if y.hi is not None and tuple(y.hi.shape) == (y.n, y.h + 2, y.w + 2, y.c) and not os.environ.get('TMX_MASK_F32'):
        return dict(y_hi=y.hi)
    return dict(y_f32=rt.split_unpack(y).f32)


def _combine(rt, act, contribs, want_planes, want_f32, mask_y=None, dbias=None, phase_pack=False, alpha=None):
    """All contributions to dL/d(act) -> (planes on the zero-ringed grid, fp32 NHWC), masked by lrelu'(mask_y);
    `mask_y` is an fp32 map or the dict returned by `_mask_of`."""
    if len(contribs) == 1 and contribs[0][0] == 'ready':     # the consumer's fused data gradient did all of it
        return contribs[0][1], contribs[0][2]
    mask_kw = mask_y if isinstance(mask_y, dict) else dict(y_f32=mask_y)
    n, h, w, c = act.n, act.h, act.w, act.c
    main = [x for x in contribs if x[0] in ('grid', 'pool')]
    f32s = [x[1] for x in contribs if x[0] == 'f32']
    if len(main) > 1:
        raise NotImplementedError('more than one grid/pool gradient contribution to one activation')
    add = None
    if main:
        if len(f32s) > 1:
            raise NotImplementedError('more than one fp32 addend next to a grid contribution')
        add = f32s[0] if f32s else None
        if main[0][0] == 'grid':
            src, kind, fold = main[0][1], 0, main[0][2]
        else:
            src, kind, fold = main[0][1], 2, 2
    else:
        if not f32s:
            return None, None
        if len(f32s) > 2:
            raise NotImplementedError('more than two fp32 gradient contributions to one activation')
        src, kind, fold = f32s[0], 1, 2
        add = f32s[1] if len(f32s) == 2 else None
    return rt.grad_prepare(src, n, h, w, c, kind, fold=fold, add=add, want_planes=want_planes,
                           want_f32=want_f32, dbias=dbias, phase_pack=phase_pack, alpha=alpha, **mask_kw)


# records whose output gradient a consuming conv's fused data gradient may prepare (backward.fused_dgrad)
_GP_PRODUCERS = ('conv', 'fromrgb', 'pool', 'window', 'concat')


def gp_fusion_maps(tape):
    """Tape pre-pass of the fused data gradient: ({id(activation): tape position of its EARLIEST consumer},
    {id(activation): the record that produced it, for the producer kinds the fused kernel can stand in for}).  The
    reverse pass walks the tape backwards, so the earliest consumer is the last one to contribute a gradient."""
    first_use, producer = {}, {}
    for pos, rec in enumerate(tape):
        for key in ('x', 'residual', 'a', 'b'):
            a = rec.get(key)
            if isinstance(a, runtime.Act):
                first_use.setdefault(id(a), pos)
        for a in (rec.get('inputs') or ()):
            if isinstance(a, runtime.Act):
                first_use.setdefault(id(a), pos)
        if rec['kind'] in _GP_PRODUCERS and isinstance(rec.get('y'), runtime.Act):
            producer[id(rec['y'])] = rec
    return first_use, producer


def _scaled(rt, x, a):
    """a * x (fp32, any shape)."""
    x = x.contiguous()
    out = rt.empty(*x.shape)
    _lib.check(rt.lib.tmx_axpb(rt.handle, _ptr(x), _ptr(out), x.numel(), float(a), 0.0, rt.stream()), 'tmx_axpb')
    return out


def _accumulate(rt, table, key, g):
    """table[key] += g for gradients that reach one tensor along several paths."""
    cur = table.get(key)
    if cur is None:
        table[key] = g
    else:
        out = rt.empty(*g.shape)
        _lib.check(rt.lib.tmx_add_f32(rt.handle, _ptr(cur.contiguous()), _ptr(g.contiguous()), _ptr(out), g.numel(),
                                      rt.stream()), 'tmx_add_f32')
        table[key] = out


def _block_sum(rt, g, factor):
    """Adjoint of upscale2d of an NCHW image (networks.py:80-88): sum over each factor x factor block."""
    n, c, h, w = g.shape
    out = rt.empty(n, c, h // factor, w // factor)
    _lib.check(rt.lib.tmx_convert_output(rt.handle, _ptr(g.contiguous()), _ptr(out), n * c, h, w,
                                         float(factor * factor), 0.0, int(factor), 0, rt.stream()),
               'tmx_convert_output')     # mean of (x * f^2) over the block == the block sum (f^2 is a power of two)
    return out


def _spread(rt, g, factor):
    """Adjoint of the VALID average pool of an NCHW image (networks.py:131-136): g / f^2 copied to the block."""
    n, c, h, w = g.shape
    dev = g.device
    ih = torch.arange(h * factor, dtype=torch.int32, device=dev).div_(factor, rounding_mode='floor')
    iw = torch.arange(w * factor, dtype=torch.int32, device=dev).div_(factor, rounding_mode='floor')
    up = rt.latent_blend([g.contiguous()], h * factor, w * factor, _lib.BLEND_COPY,
                         idx_h=[ih.repeat(n, 1).contiguous()], idx_w=[iw.repeat(n, 1).contiguous()])
    return _scaled(rt, up, 1.0 / (factor * factor))


def conv_wgrad_into(rt, net, rec, x_planes, dz, dw):
    """Weight gradient of a (non-upsampling) conv record into `dw` (HWIO view of the flat gradient).  A conv whose
    input was channel-padded (513 -> 576 after minibatch stddev) goes through a padded scratch gradient."""
    x = rec['x']
    if x.c == rec['cin']:
        rt.conv_wgrad(x_planes, dz, x.n, x.h, x.w, x.c, rec['cout'], rec['k'], rec['wscale'], dw)
        return
    taps = rec['k'] * rec['k']
    tmp = torch.zeros(taps, x.c, rec['cout'], dtype=torch.float32, device=rt.device)
    rt.conv_wgrad(x_planes, dz, x.n, x.h, x.w, x.c, rec['cout'], rec['k'], rec['wscale'], tmp)
    dwv = dw.view(taps, rec['cin'], rec['cout'])
    for t in range(taps):      # rows of the real channels are a contiguous prefix of each tap
        a, b = dwv[t], tmp[t, :rec['cin']]
        _lib.check(rt.lib.tmx_add_f32(rt.handle, _ptr(a), _ptr(b), _ptr(a), a.numel(), rt.stream()), 'tmx_add_f32')


def backward(net, tape, out_grads, flat_grad, want_input_grads=True, param_grads=True, adjoints=None, seeds=None):
    """Differentiate one recorded evaluation of `net`.
    out_grads : list matching the network outputs (NCHW fp32 device tensors, or None for unused outputs)
    flat_grad : fp32 buffer like net.flat; variable gradients are accumulated into it
    param_grads=False: only propagate to the inputs (a fixed critic inside the E/G loss: flat_grad may be None)
    adjoints : optional dict filled with the adjoint that met each parametrised layer, keyed by tape position
               (conv: dz planes on the zero-ringed grid; fromrgb: masked dz fp32; dense: (dy, y)) - the WGAN-GP
               double backward pairs them with tangent activations (loss.gradient_penalty)
    seeds    : optional [(activation, fp32 NHWC gradient)] injected at intermediate activations
    Returns the list of gradients w.r.t. the network inputs (NCHW fp32; None for image inputs of the
    encoders unless the first layer's data gradient is defined, i.e. for D_patch/FromRGB: NCHW image gradient)."""
    rt = net.rt
    assert tape and tape[-1]['kind'] == 'outputs'
    outs = tape[-1]['tensors']
    assert len(out_grads) == len(outs)
    grads = _Grads()
    for act, g in (seeds or []):
        grads.add(act, ('f32', g))
    by_tensor = {id(t): g for t, g in zip(outs, out_grads) if g is not None}
    input_grads = {}

    scratch = {}

    def gview(name):
        if param_grads:
            return net.grad_view(flat_grad, name)
        v = net.vars[name]                      # throw-away target for kernels that always emit a weight gradient
        if v.size not in scratch:
            scratch[v.size] = torch.zeros(v.size, dtype=torch.float32, device=rt.device)
        return scratch[v.size]

    tgrads = {}        # gradients w.r.t. plain tensors (dense head), keyed by tensor identity

    # fused data gradient + grad_prepare (tmx_conv2d_dgrad_gp): possible when the conv at hand is the LAST contributor
    # to its input activation (every other consumer sits later on the tape, i.e. has already been differentiated) and
    # that activation is the plain output of another conv record, whose mask / bias gradient / residual need are then
    # known at the time of the data gradient
    fuse_gp = not os.environ.get('TMX_NO_FUSED_GP')
    first_use, producer = gp_fusion_maps(tape) if fuse_gp else ({}, {})

    def fused_dgrad(pos, rec, x, dz, wt, hs, ws_, cin_g, ng, k, fold):
        """Data gradient of `rec` w.r.t. `x` with x's grad_prepare folded in; False when not applicable."""
        prod = producer.get(id(x)) if fuse_gp else None
        if prod is None or prod.get('up2') or first_use.get(id(x)) != pos:
            return False
        pending = grads.by_id.get(id(x), [])
        if any(cn[0] != 'f32' for cn in pending) or len(pending) > 1:
            return False
        kind = prod['kind']
        if kind == 'conv':          # what the producer's own record would ask of _combine
            mask = _mask_of(rt, x) if prod['act'] else {}
            dbias = gview(prod['b']) if (param_grads and prod['b'] is not None) else None
            planes, f32 = True, prod['residual'] is not None
        elif kind == 'fromrgb':     # fp32 only: the 1x1 image head differentiates on the fp32 map
            mask, dbias, planes, f32 = dict(y_f32=x.f32), (gview(prod['b']) if param_grads else None), False, True
        else:                       # pool / window / concat: the plain fp32 gradient
            mask, dbias, planes, f32 = {}, None, False, True
        out = rt.conv_dgrad_gp(dz, x.n, hs, ws_, cin_g, ng, k, wt, fold, add=pending[0][1] if pending else None,
                               want_f32=f32, want_planes=planes, dbias=dbias, alpha=prod.get('alpha'), **mask)
        if out is None:
            return False
        grads.pop(x)
        grads.add(x, ('ready', out[0], out[1]))
        return True

    for pos in range(len(tape) - 2, -1, -1):
        rec = tape[pos]
        kind = rec['kind']
        if kind == 'tanh':                 # trailing tanh of G_res at lod != 0 (networks.py:482-483)
            g = by_tensor.pop(id(rec['y']), None)
            if g is not None:
                dx = rt.empty(*g.shape)
                _lib.check(rt.lib.tmx_tanh_bwd(rt.handle, _ptr(g.contiguous()), _ptr(rec['y']), _ptr(dx), g.numel(),
                                               rt.stream()), 'tmx_tanh_bwd')
                _accumulate(rt, by_tensor, id(rec['x']), dx)
        elif kind == 'imgup':              # upscale2d of an image head (networks.py:476-477)
            g = by_tensor.pop(id(rec['y']), None)
            if g is not None:
                _accumulate(rt, by_tensor, id(rec['x']), _block_sum(rt, g, rec['factor']))
        elif kind == 'imglerp':            # fade between two image heads: a + (b - a) t
            g = by_tensor.pop(id(rec['y']), None)
            if g is not None:
                _accumulate(rt, by_tensor, id(rec['a']), _scaled(rt, g, 1.0 - rec['t']))
                _accumulate(rt, by_tensor, id(rec['b']), _scaled(rt, g, rec['t']))
        elif kind == 'lerp':               # fade between two feature maps (networks.py:281,373,573)
            contribs = grads.pop(rec['y'])
            if contribs:
                _, g = _combine(rt, rec['y'], contribs, want_planes=False, want_f32=True)
                grads.add(rec['a'], ('f32', _scaled(rt, g, 1.0 - rec['t'])))
                grads.add(rec['b'], ('f32', _scaled(rt, g, rec['t'])))
        elif kind == 'pixelnorm':          # pixel_norm (networks.py:170-172)
            contribs = grads.pop(rec['y'])
            if contribs:
                _, g = _combine(rt, rec['y'], contribs, want_planes=False, want_f32=True)
                x = rec['x']
                dx = rt.empty(x.n, x.h, x.w, x.c)
                _lib.check(rt.lib.tmx_pixel_norm_bwd(rt.handle, _ptr(x.f32), _ptr(g), _ptr(dx), x.n * x.h * x.w, x.c,
                                                     float(rec['eps']), rt.stream()), 'tmx_pixel_norm_bwd')
                grads.add(x, ('f32', dx))
        elif kind == 'bias_act':           # [lrelu](x + b) behind the fused conv2d_downscale2d (networks.py:142-148)
            contribs = grads.pop(rec['y'])
            if contribs:
                y = rec['y']
                _, g = _combine(rt, y, contribs, want_planes=False, want_f32=True, mask_y=y.f32 if rec['act'] else None,
                                dbias=gview(rec['b']) if param_grads else None)
                grads.add(rec['x'], ('f32', g))
        elif kind == 'window':             # G_res(tail_window=...): slice of a feature map -> zeros outside
            contribs = grads.pop(rec['y'])
            if contribs:
                _, g = _combine(rt, rec['y'], contribs, want_planes=False, want_f32=True)
                x = rec['x']
                grads.add(x, ('f32', rt.window_embed(g, rec['win'], x.h, x.w, nhwc=True)))
        elif kind == 'vggpre':             # custom_vgg19.py:31-40 input scaling (16-channel padded NHWC)
            contribs = grads.pop(rec['y'])
            if contribs and want_input_grads:
                y = rec['y']
                _, g = _combine(rt, y, contribs, want_planes=False, want_f32=True)
                img = rec['img']
                n, _, h, w = img.shape
                dimg = rt.empty(n, 3, h, w)
                _lib.check(rt.lib.tmx_vgg_preprocess_bwd(rt.handle, _ptr(g), _ptr(dimg), n, h, w, rt.stream()),
                           'tmx_vgg_preprocess_bwd')
                _accumulate(rt, input_grads, ('img', id(img)), dimg)
        elif kind == 'imgpool':            # the input image pooled for a lower level of detail (networks.py:278,281)
            g = input_grads.pop(('img', id(rec['y'])), None)
            if g is not None:
                _accumulate(rt, input_grads, ('img', id(rec['x'])), _spread(rt, g, rec['factor']))
        elif kind == 'slice':
            g = by_tensor.get(id(rec['out']))
            if g is not None:
                a = rec['x']
                key = ('slicebuf', id(a))
                buf = input_grads.get(key)
                if buf is None:
                    buf = torch.zeros(a.n, a.h, a.w, a.c, dtype=torch.float32, device=rt.device)
                    input_grads[key] = buf
                    grads.add(a, ('f32', buf))
                rt.nchw_to_nhwc(g.contiguous(), out=buf, c_off=rec['c_off'], c_total=a.c)
        elif kind == 'torgb':
            g = by_tensor.get(id(rec['img']))
            if g is None:
                continue
            a = rt.split_unpack(rec['x'])
            dy = rt.empty(a.n, a.h, a.w, a.c)
            w = net.vars[rec['w']]
            _lib.check(rt.lib.tmx_torgb_bwd(rt.handle, _ptr(g.contiguous()), _ptr(rec['img']), _ptr(a.f32),
                                            _ptr(w.value), float(rec['wscale']), _ptr(dy), _ptr(gview(rec['w'])),
                                            _ptr(gview(rec['b'])), a.n, a.h, a.w, a.c, w.shape[3], int(rec['tanh']),
                                            rt.stream()), 'tmx_torgb_bwd')
            grads.add(a, ('f32', dy))
        elif kind == 'dense':
            g = by_tensor.get(id(rec['y']))
            if g is None and rec.get('alias') is not None:
                g = by_tensor.get(id(rec['alias']))
            if g is None:
                g = tgrads.pop(id(rec['y']), None)
            if g is None:
                continue
            y, xin = rec['y'], rec['x']
            n, cout = y.shape
            kdim = xin.numel() // n
            g = g.contiguous()
            if adjoints is not None:
                adjoints[pos] = (g, y)
            wv = net.vars[rec['w']]
            if param_grads:
                _lib.check(rt.lib.tmx_dense_wgrad(rt.handle, _ptr(xin), _ptr(g), _ptr(y), _ptr(gview(rec['w'])),
                                                  _ptr(gview(rec['b'])), n, kdim, cout, float(rec['wscale']),
                                                  int(rec['act']), runtime.LRELU_ALPHA, rt.stream()),
                           'tmx_dense_wgrad')
            dx = rt.empty(*xin.shape)
            _lib.check(rt.lib.tmx_dense_bwd_input(rt.handle, _ptr(g.contiguous()), _ptr(y), _ptr(wv.value),
                                                  float(rec['wscale']), _ptr(dx), n, kdim, cout, int(rec['act']),
                                                  runtime.LRELU_ALPHA, rt.stream()), 'tmx_dense_bwd_input')
            tgrads[id(xin)] = dx
        elif kind == 'flatten':
            g = tgrads.pop(id(rec['y']), None)
            if g is None:
                continue
            a = rec['x']
            grads.add(a, ('f32', rt.nchw_to_nhwc(g.view(a.n, a.c, a.h, a.w))))
        elif kind == 'mbstd':
            contribs = grads.pop(rec['y'])
            if not contribs:
                continue
            y, x = rec['y'], rec['x']
            _, dy = _combine(rt, y, contribs, want_planes=False, want_f32=True)
            g = min(rec['group'], x.n)
            dx = rt.empty(x.n, x.h, x.w, x.c)
            ds = rt.empty(x.n // g)
            _lib.check(rt.lib.tmx_mbstd_bwd(rt.handle, _ptr(rt.split_unpack(x).f32), _ptr(dy), _ptr(dx), _ptr(ds), x.n,
                                            x.h, x.w, x.c, y.c, rec['group'], rt.stream()), 'tmx_mbstd_bwd')
            if adjoints is not None:
                adjoints[pos] = ds          # adjoint of the group statistic
            grads.add(x, ('f32', dx))
        elif kind == 'view':
            # y holds the first `pixels` pixels of x (or vice versa) in another [n,h,w] arrangement
            contribs = grads.pop(rec['y'])
            if not contribs:
                continue
            _, f32 = _combine(rt, rec['y'], contribs, want_planes=False, want_f32=True)
            x, m = rec['x'], rec['pixels']
            gx = torch.zeros(x.n * x.h * x.w, x.c, dtype=torch.float32, device=rt.device)
            gx[:m].copy_(f32.view(-1, x.c)[:m])
            grads.add(x, ('f32', gx.view(x.n, x.h, x.w, x.c)))
        elif kind == 'pool':
            contribs = grads.pop(rec['y'])
            if not contribs:
                continue
            _, f32 = _combine(rt, rec['y'], contribs, want_planes=False, want_f32=True)
            grads.add(rec['x'], ('pool', f32))
        elif kind == 'conv':
            y, x = rec['y'], rec['x']
            contribs = grads.pop(y)
            if not contribs:
                continue
            cin_g, cout, k, up2 = x.c, rec['cout'], rec['k'], rec['up2']
            has_res = rec['residual'] is not None
            mask = _mask_of(rt, y) if rec['act'] else None
            zero_pad = rec.get('halo') == 'zero'        # fused_scale layers: SAME (zero) padding, nothing to fold
            dz, dz_f32 = _combine(rt, y, contribs, want_planes=True, want_f32=has_res, mask_y=mask,
                                  dbias=gview(rec['b']) if (param_grads and rec['b'] is not None) else None,
                                  phase_pack=up2, alpha=rec.get('alpha'))
            if has_res:
                grads.add(rec['residual'], ('f32', dz_f32))         # y = conv(x) + residual (networks.py:437)
            if adjoints is not None:
                adjoints[pos] = dz
            w = net.vars[rec['w']]
            fwd = rec.get('w_planes') or net.prepared_weights(w, rec['wscale'], k, rec['cin'], cout, up2_phase=up2,
                                                              cin_pad=x.c)
            if up2:
                # sub-pixel form: low-res geometry, 4*Cout phase channels
                hs, ws_, ng = x.h, x.w, 4 * cout
                if param_grads:
                    dwp = torch.zeros(9, cin_g, ng, dtype=torch.float32, device=rt.device)
                    rt.conv_wgrad((x.hi, x.lo), dz, x.n, hs, ws_, cin_g, ng, 3, rec['wscale'], dwp)
                    if rec.get('transposed_var'):
                        # upscale2d_conv2d: the kernel above differentiates w_eq[u,v,ci,co] = var[2-u,2-v,co,ci]
                        dw_eq = torch.zeros(3, 3, cin_g, cout, dtype=torch.float32, device=rt.device)
                        _lib.check(rt.lib.tmx_conv_wgrad_unphase(rt.handle, _ptr(dwp), _ptr(dw_eq), cin_g, cout,
                                                                 rt.stream()), 'tmx_conv_wgrad_unphase')
                        dvar = dw_eq.flip(0, 1).permute(0, 1, 3, 2).contiguous()         # layout move only
                        gv = gview(rec['w'])
                        _lib.check(rt.lib.tmx_add_f32(rt.handle, _ptr(gv), _ptr(dvar), _ptr(gv), gv.numel(),
                                                      rt.stream()), 'tmx_add_f32')
                    else:
                        _lib.check(rt.lib.tmx_conv_wgrad_unphase(rt.handle, _ptr(dwp), _ptr(gview(rec['w'])), cin_g,
                                                                 cout, rt.stream()), 'tmx_conv_wgrad_unphase')
                wt = net.cached(('wt', rec['w'], True), lambda: rt.transpose_weights(fwd, ng, 9, cin_g))
                if not fused_dgrad(pos, rec, x, dz, wt, hs, ws_, cin_g, ng, 3, 2 if zero_pad else 1):
                    g = rt.conv_dgrad(dz, x.n, hs, ws_, cin_g, ng, 3, wt)
                    grads.add(x, ('grid', g, 2 if zero_pad else 1))
            else:
                if param_grads:
                    conv_wgrad_into(rt, net, rec, (x.hi, x.lo), dz, gview(rec['w']))
                wt = net.cached(('wt', rec['w'], False), lambda: rt.transpose_weights(fwd, cout, k * k, cin_g))
                fold = 0 if (k == 3 and not zero_pad) else 2
                if not fused_dgrad(pos, rec, x, dz, wt, x.h, x.w, cin_g, cout, k, fold):
                    g = rt.conv_dgrad(dz, x.n, x.h, x.w, cin_g, cout, k, wt)
                    grads.add(x, ('grid', g, fold))
        elif kind == 'fromrgb':
            y = rec['y']
            contribs = grads.pop(y)
            if not contribs:
                continue
            _, dz = _combine(rt, y, contribs, want_planes=False, want_f32=True, mask_y=y.f32,
                             dbias=gview(rec['b']) if param_grads else None)
            if adjoints is not None:
                adjoints[pos] = dz
            img = rec['img']
            n, cimg, h, w_ = img.shape
            dimg = rt.empty(n, cimg, h, w_) if want_input_grads else None
            wv = net.vars[rec['w']]
            _lib.check(rt.lib.tmx_fromrgb_bwd(rt.handle, _ptr(img), _ptr(dz), _ptr(wv.value), float(rec['wscale']),
                                              _ptr(gview(rec['w'])), _ptr(dimg), n, cimg, h, w_, rec['cout'],
                                              rt.stream()), 'tmx_fromrgb_bwd')
            if dimg is not None:
                _accumulate(rt, input_grads, ('img', id(img)), dimg)
        elif kind == 'concat':
            y = rec['y']
            contribs = grads.pop(y)
            if not contribs or not want_input_grads:
                continue
            _, f32 = _combine(rt, y, contribs, want_planes=False, want_f32=True)
            c = rec['c']
            input_grads['concat'] = [rt.nhwc_to_nchw(f32, c_off=i * c, c=c) for i in range(len(rec['inputs']))]
        else:
            raise NotImplementedError('backward of tape record %r' % kind)
    if 'concat' in input_grads:
        return input_grads['concat']
    imgs = [v for k, v in input_grads.items() if isinstance(k, tuple) and k[0] == 'img' and v is not None]
    return imgs if imgs else []
