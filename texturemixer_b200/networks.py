"""Build functions of the TextureMixer hot path with the reference's names and
keyword signatures (networks.py:194-211, 296-314, 388-409, 491-505), so that
`config.*.func = 'networks.G_res'` etc. resolve here (network.import_module).

They are written against `network.T` handles: in *template* mode they only
declare variables (same names, shapes and creation order as the reference's
`recursive` structure, tf.cond building both branches) and propagate shapes;
in *run* mode every layer is one libtmx launch on device activations
(`runtime.Act`).  No arithmetic happens in Python/torch.

Layer variants on the device: the reference defaults (use_wscale=True,
fused_scale=False, leaky ReLU, float32) plus `use_pixelnorm` in the generator;
the others raise NotImplementedError.  Progressive growing
(`lod` > 0, integer or fractional: the tf.cond trees of networks.py:276-282,
368-374, 473-479, 568-574) is evaluated and recorded on the tape like every
other layer."""
import functools
import os

import numpy as np
import torch

from .network import T
from .runtime import Act

SQRT2 = float(np.sqrt(2))


# ---------------------------------------------------------------------- primitives
def _wscale(shape, gain):
    """get_weight, networks.py:26-33 (use_wscale=True): float32(gain / sqrt(fan_in))."""
    return _wscale_cached(tuple(shape), float(gain))


@functools.lru_cache(maxsize=None)
def _wscale_cached(shape, gain):
    return float(np.float32(gain / np.sqrt(np.prod(shape[:-1]))))


def _check_variants(use_wscale, use_leakyrelu, fused_scale, dtype):
    if not use_wscale or not use_leakyrelu or dtype != 'float32':
        raise NotImplementedError('texturemixer_b200: only use_wscale=True, use_leakyrelu=True, dtype=float32 (the '
                                  'reference defaults; use_pixelnorm and fused_scale are free) are implemented on the '
                                  'device')


def _act_of(t):
    """Device activation of a run-mode handle (NCHW -> NHWC on first use)."""
    if t.act is None:
        x = t.nchw
        n, c, h, w = x.shape
        t.act = Act(n, h, w, c, f32=t.ctx.rt.nchw_to_nhwc(x))
    return t.act


def conv2d_layer(x, fmaps, kernel, gain=SQRT2, act=True, residual=None, up2=False, next_tc=False,
                 next_up2=False, keep_f32=False, torgb=None, pad='reflect', alpha=None, wscale=None, trainable=True,
                 next_pad=None):
    """act(apply_bias(conv2d(x))) [+ residual] under the current variable scope:
    networks.py:48-56 + 61-67 + 72-75 (+ :437).  `up2`: the logical input is
    upscale2d(x) (networks.py:448), never materialised.  `next_tc`/`next_up2` are
    layout hints: the consumer is a tensor-core conv (wants split-bf16 planes),
    reading through an upscale2d (wants a REPLICATE halo).  `torgb` =
    (scope, num_channels, tanh): fuse that 1x1 image head into the epilogue."""
    assert kernel >= 1 and kernel % 2 == 1                                # networks.py:49
    ctx = x.ctx
    cin = x.shape[1]
    w = ctx.get_variable('weight', (kernel, kernel, cin, fmaps), trainable=trainable)
    b = ctx.get_variable('bias', (fmaps,), init='zeros', trainable=trainable)
    f = 2 if up2 else 1
    shape = [x.shape[0], fmaps, _mul(x.shape[2], f), _mul(x.shape[3], f)]
    if ctx.mode == 'template':
        return T(shape, ctx)
    if kernel not in (1, 3):
        raise NotImplementedError('conv2d kernel=%d: only 1 and 3 occur on the path' % kernel)
    rt = ctx.rt
    xa = _act_of(x)
    # `pad='zero'` (SAME), `alpha` (0 = ReLU) and `wscale` (1 = plain filters) are the VGG-19 layer of
    # tensorflow_vgg's conv_layer; `next_pad`: halo kind the consumer wants on the written planes
    ws = _wscale(w.shape, gain) if wscale is None else float(wscale)
    strip = None
    if kernel == 1 and (xa.h < 2 or xa.w < 2) and ctx.tape is not None:
        # a 1x1 conv couples no pixels: run it on the pixels laid out as one 2-row strip so that the
        # tensor-core kernels (which need a 2x2 halo layout) and their backward apply (E_zg's zg_Conv3)
        strip = _pixel_strip(ctx, xa)
        xa = strip
    algo = rt.choose_algo(xa.c, fmaps, kernel, up2, (xa.h, xa.w))
    if xa.c != cin and algo == 1:
        raise NotImplementedError('conv2d: a channel-padded activation (%d -> %d) needs the tensor-core kernel'
                                  % (cin, xa.c))
    prepared = None
    if algo != 1:  # tensor-core path: cached bf16 hi/lo weight planes (zero rows for padded input channels)
        prepared = ctx.net.prepared_weights(w, ws, kernel, cin, fmaps, up2_phase=up2, cin_pad=xa.c,
                                            xmerge=rt.use_xmerge(xa.c, kernel, up2))
    res = None if residual is None else rt.split_unpack(_act_of(residual)).f32
    head = None
    pn = getattr(ctx, 'pixelnorm', None) if act else None      # PN(act(...)) of networks.py:216,320,415
    if pn is not None:
        torgb, next_tc, keep_f32 = None, False, True           # the normalised map is what every consumer reads
    rec = ctx.tape is not None
    if rec:
        # training forward: keep what the backward needs - the input planes (weight gradient), the fp32
        # output (activation mask, pooling) and the output planes (next layer) - and no fused image head
        if algo == 1:
            raise NotImplementedError('backward needs the tensor-core conv (channels multiple of 16), got %d -> %d'
                                      % (cin, fmaps))
        # (the fp32 map only where a consumer reads it: pooling, the residual stream, image heads; the leaky-ReLU mask
        # of the backward comes from the bf16 hi plane, backward._mask_of)
        torgb, keep_f32, next_tc = None, keep_f32 or not next_tc or bool(os.environ.get('TMX_KEEP_F32')), True
    if torgb is not None and algo != 1:
        scope, nch, tanh = torgb
        wr, br = ctx.net.vars[scope + '/weight'], ctx.net.vars[scope + '/bias']
        head = (wr.value, br.value, _wscale(wr.shape, 1.0), nch, tanh)
    out = rt.conv2d(xa, w.value, b.value, ws, kernel, fmaps, lrelu=act, residual=res, up2=up2,
                    want_f32=((not next_tc) and head is None) or keep_f32, want_split=next_tc,
                    halo_out=next_pad or ('replicate' if next_up2 else 'reflect'), algo=algo, prepared=prepared,
                    torgb=head, halo_in='zero' if pad == 'zero' else None, alpha=alpha)
    t = T(shape, ctx)
    if head is not None:
        t.act, images = out
        t.rgb = (torgb[0], torgb[2], images)
    else:
        t.act = out
    if rec:
        ctx.tape.append(dict(kind='conv', x=xa, y=t.act, w=w.name, b=b.name, wscale=ws, k=kernel, cin=cin, cout=fmaps,
                             act=act, up2=up2, residual=None if residual is None else _act_of(residual), alpha=alpha,
                             halo='zero' if pad == 'zero' else None))
    if pn is not None:
        t = _pixel_norm(ctx, t, pn)
    if strip is not None:
        orig = _act_of(x)
        m = orig.n * orig.h * orig.w
        flat = rt.split_unpack(t.act).f32.view(-1, fmaps)[:m]
        out = Act(orig.n, orig.h, orig.w, fmaps, f32=flat.view(orig.n, orig.h, orig.w, fmaps))
        ctx.tape.append(dict(kind='view', x=t.act, y=out, pixels=m))
        t.act = out
    return t


def _fused_conv_prologue(x, wshape, fmaps, out_hw_factor):
    """Variables and template shape shared by the two fused_scale layers."""
    ctx = x.ctx
    w = ctx.get_variable('weight', wshape)
    b = ctx.get_variable('bias', (fmaps,), init='zeros')
    num, den = out_hw_factor
    shape = [x.shape[0], fmaps, None if x.shape[2] is None else x.shape[2] * num // den,
             None if x.shape[3] is None else x.shape[3] * num // den]
    return ctx, w, b, shape


def upscale2d_conv2d_layer(x, fmaps, gain=SQRT2):
    """act(apply_bias(upscale2d_conv2d(x))) (networks.py:94-101, 444-446): conv2d_transpose, stride 2, SAME, with the
    3x3 variable [k,k,fmaps,Cin] fused to 4x4.  Identical to the 3x3 convolution of the nearest-neighbour upscaled
    input under ZERO padding with the kernel flipped and its channel axes swapped (checked against the reference
    code in tests/golden/networks_fused.npz), i.e. the sub-pixel upsample+conv kernel over zero-halo planes."""
    cin = x.shape[1]
    ctx, w, b, shape = _fused_conv_prologue(x, (3, 3, fmaps, cin), fmaps, (2, 1))
    if ctx.mode == 'template':
        return T(shape, ctx)
    rt = ctx.rt
    xa = _act_of(x)
    ws = float(np.float32(gain / np.sqrt(9 * cin)))                       # fan_in = k*k*Cin (networks.py:96)
    w_eq, planes = ctx.net.cached(('fused_up', w.name), lambda: _fused_up_weights(rt, w.value, ws, cin, fmaps))
    out = rt.conv2d(xa, w_eq, b.value, ws, 3, fmaps, lrelu=True, up2=True, want_f32=True, want_split=False, algo=2,
                    prepared=planes, halo_in='zero')
    if ctx.tape is not None:
        # the backward differentiates the equivalent sub-pixel conv: weight gradient w.r.t. w_eq, mapped back to
        # the [k,k,fmaps,Cin] variable by the adjoint of the flip / channel swap; ZERO padding has no fold
        ctx.tape.append(dict(kind='conv', x=xa, y=out, w=w.name, b=b.name, wscale=ws, k=3, cin=cin, cout=fmaps,
                             act=True, up2=True, residual=None, halo='zero', w_planes=planes, transposed_var=True))
    t = T(shape, ctx, act=out)
    return _pixel_norm(ctx, t, ctx.pixelnorm) if ctx.pixelnorm is not None else t


def _fused_up_weights(rt, w_var, ws, cin, fmaps):
    w_eq = w_var.flip(0, 1).permute(0, 1, 3, 2).contiguous()              # [3,3,Cin,fmaps], taps flipped (layout move)
    return w_eq, rt.prepare_weights(w_eq, ws, 3, cin, fmaps, up2_phase=True)


def conv2d_downscale2d_layer(x, fmaps, gain=SQRT2, act=True):
    """[act](apply_bias(conv2d_downscale2d(x))) (networks.py:142-148): stride-2 SAME conv with the 3x3 variable fused
    to 4x4 and scaled by 1/4 == the 2x2 average of the ZERO-padded 3x3 convolution; bias and activation follow the
    average (they precede it in the unfused form)."""
    cin = x.shape[1]
    ctx, w, b, shape = _fused_conv_prologue(x, (3, 3, cin, fmaps), fmaps, (1, 2))
    if ctx.mode == 'template':
        return T(shape, ctx)
    rt = ctx.rt
    xa = _act_of(x)
    ws = _wscale(w.shape, gain)
    planes = ctx.net.prepared_weights(w, ws, 3, cin, fmaps, cin_pad=xa.c)
    lin = rt.conv2d(xa, w.value, None, ws, 3, fmaps, lrelu=False, want_f32=True, want_split=False, algo=2,
                    prepared=planes, halo_in='zero')
    pooled = rt.avgpool2(lin)
    out = rt.bias_act(pooled, b.value, act)
    if ctx.tape is not None:
        ctx.tape.append(dict(kind='conv', x=xa, y=lin, w=w.name, b=None, wscale=ws, k=3, cin=cin, cout=fmaps,
                             act=False, up2=False, residual=None, halo='zero'))
        ctx.tape.append(dict(kind='pool', x=lin, y=pooled))
        ctx.tape.append(dict(kind='bias_act', x=pooled, y=out, b=b.name, act=act))
    t = T(shape, ctx, act=out)
    return _pixel_norm(ctx, t, ctx.pixelnorm) if (ctx.pixelnorm is not None and act) else t


def _pixel_strip(ctx, a):
    """[n,h,w,c] -> the same pixels as one image [1,2,ceil(m/2),c] (zero padded); records the view."""
    rt = ctx.rt
    rt.split_unpack(a)
    m = a.n * a.h * a.w
    m2 = max((m + 1) // 2, 2)          # at least 2x2 stored pixels (halo layout), also for batches of 1-2
    buf = torch.zeros(2 * m2, a.c, dtype=torch.float32, device=rt.device)
    buf[:m].copy_(a.f32.view(m, a.c))
    strip = Act(1, 2, m2, a.c, f32=buf.view(1, 2, m2, a.c))
    ctx.tape.append(dict(kind='view', x=a, y=strip, pixels=m))
    return strip


def _mul(d, f):
    return None if d is None else d * f


def downscale2d(x, factor=2):
    """networks.py:131-136."""
    assert isinstance(factor, int) and factor >= 1
    if factor == 1:
        return x
    ctx = x.ctx
    shape = [x.shape[0], x.shape[1], x.shape[2] // factor, x.shape[3] // factor]
    if ctx.mode == 'template':
        return T(shape, ctx)
    if x.act is None and x.nchw is not None:
        # the network's input image pooled for a lower level of detail (networks.py:278,281): one VALID average
        # pool of the whole factor on the NCHW planes (3 channels: not an NHWC/tensor-core layout)
        pooled = _pool_image(ctx.rt, x.nchw, factor)
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='imgpool', x=x.nchw, y=pooled, factor=factor))
        return T(shape, ctx, nchw=pooled)
    a = _act_of(x)
    f = factor
    while f > 1:
        assert f % 2 == 0
        # the last pooling step feeds the next block's 3x3 conv (networks.py:238-256, 340-350, 530-548): written as
        # split planes with the REFLECT halo by the pooling kernel itself when that conv runs on the tensor cores
        tc_next = f == 2 and a.c % 16 == 0 and a.h >= 4 and a.w >= 4
        b = ctx.rt.avgpool2(a, pack='reflect' if tc_next else None)
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='pool', x=a, y=b))
        a = b
        f //= 2
    return T(shape, ctx, act=a)


def _pool_image(rt, img, factor):
    import ctypes as C
    from . import _lib
    n, c, h, w = img.shape
    out = rt.empty(n, c, h // factor, w // factor)
    _lib.check(rt.lib.tmx_convert_output(rt.handle, C.c_void_p(img.data_ptr()), C.c_void_p(out.data_ptr()), n * c, h, w,
                                         1.0, 0.0, int(factor), 0, rt.stream()), 'tmx_convert_output')
    return out


def _lerp_lod(rt, a, b, t):
    """tfutil.lerp(a, b, t) = a + (b - a) * t (tfutil.py:41-43) on two equally shaped fp32 device tensors with the
    scalar t = lod_in - lod, in float32 like the graph."""
    from . import _lib
    n = a.shape[0]
    per = a[0].numel()
    tt = torch.full((n,), float(np.float32(t)), dtype=torch.float32, device=a.device)
    out = rt.latent_blend([a.contiguous().view(n, 1, 1, per), b.contiguous().view(n, 1, 1, per)], 1, per,
                          _lib.BLEND_LERP, t=tt)
    return out.view(a.shape)


def _upscale_image(rt, img, factor):
    """upscale2d (networks.py:80-88) of an NCHW image: nearest-neighbour gather y[i][j] = x[i // f][j // f]."""
    from . import _lib
    n, c, h, w = img.shape
    ih = torch.arange(h * factor, dtype=torch.int32, device=img.device).div_(factor, rounding_mode='floor')
    iw = torch.arange(w * factor, dtype=torch.int32, device=img.device).div_(factor, rounding_mode='floor')
    return rt.latent_blend([img.contiguous()], h * factor, w * factor, _lib.BLEND_COPY,
                           idx_h=[ih.repeat(n, 1).contiguous()], idx_w=[iw.repeat(n, 1).contiguous()])


def _window(t, win):
    """[n, oy:oy+h, ox:ox+w, :] of a run-mode handle's feature map as a new activation (recorded on the tape).  The
    offset may live in device memory (loss.Window.dev) so that the launch does not change from step to step."""
    ctx = t.ctx
    oy, ox, h, w = win
    a = ctx.rt.split_unpack(_act_of(t))
    if getattr(win, 'dev', None) is None and (oy < 0 or ox < 0 or oy + h > a.h or ox + w > a.w):
        raise ValueError('tail_window %r outside the %dx%d feature map' % (tuple(win), a.h, a.w))
    out = Act(a.n, h, w, a.c, f32=ctx.rt.window(a.f32, win, nhwc=True))
    if ctx.tape is not None:
        ctx.tape.append(dict(kind='window', x=a, y=out, win=win))
    return T([t.shape[0], t.shape[1], h, w], ctx, act=out)


def _tanh(rt, x):
    import ctypes as C
    from . import _lib
    out = rt.empty(*x.shape)
    _lib.check(rt.lib.tmx_tanh_f32(rt.handle, C.c_void_p(x.contiguous().data_ptr()), C.c_void_p(out.data_ptr()),
                                   x.numel(), rt.stream()), 'tmx_tanh_f32')
    return out


def _pixel_norm(ctx, t, epsilon):
    """pixel_norm (networks.py:170-172) of a run-mode handle: x * rsqrt(mean_c x^2 + eps) on the NHWC fp32 map."""
    import ctypes as C
    from . import _lib
    rt = ctx.rt
    a = rt.split_unpack(_act_of(t))
    out = rt.empty(a.n, a.h, a.w, a.c)
    _lib.check(rt.lib.tmx_pixel_norm(rt.handle, C.c_void_p(a.f32.data_ptr()), C.c_void_p(out.data_ptr()),
                                     a.n * a.h * a.w, a.c, float(epsilon), rt.stream()), 'tmx_pixel_norm')
    y = Act(a.n, a.h, a.w, a.c, f32=out)
    if ctx.tape is not None:
        ctx.tape.append(dict(kind='pixelnorm', x=a, y=y, eps=float(epsilon)))
    return T(t.shape, ctx, act=y)


def _fromrgb(x, fmaps, name):
    """act(apply_bias(conv2d(x, kernel=1))) straight from the NCHW image
    (networks.py:226-228, 330-332, 518-520)."""
    ctx = x.ctx
    with ctx.variable_scope(name):
        cin = x.shape[1]
        w = ctx.get_variable('weight', (1, 1, cin, fmaps))
        b = ctx.get_variable('bias', (fmaps,), init='zeros')
        shape = [x.shape[0], fmaps, x.shape[2], x.shape[3]]
        if ctx.mode == 'template':
            return T(shape, ctx)
        if x.nchw is None:
            a = ctx.rt.split_unpack(_act_of(x))
            x.nchw = ctx.rt.nhwc_to_nchw(a.f32)
        out = ctx.rt.fromrgb(x.nchw, w.value, b.value, _wscale(w.shape, SQRT2), fmaps, lrelu=True)
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='fromrgb', img=x.nchw, y=out, w=w.name, b=b.name, wscale=_wscale(w.shape, SQRT2),
                                 cout=fmaps))
        return T(shape, ctx, act=out)


def _slice_outputs(t, latent_channels, names):
    """x[:, :C], x[:, C:] as NCHW outputs (networks.py:289-290, 381-382)."""
    ctx = t.ctx
    outs = []
    for i, name in enumerate(names):
        shape = [t.shape[0], latent_channels, t.shape[2], t.shape[3]]
        if ctx.mode == 'template':
            outs.append(T(shape, ctx, name=name))
        else:
            a = ctx.rt.split_unpack(_act_of(t))
            nchw = ctx.rt.nhwc_to_nchw(a.f32, c_off=i * latent_channels, c=latent_channels)
            if ctx.tape is not None:
                ctx.tape.append(dict(kind='slice', x=a, out=nchw, c_off=i * latent_channels, c=latent_channels))
            outs.append(T(shape, ctx, nchw=nchw, name=name))
    return tuple(outs)


def _lod(ctx):
    ctx.get_variable('lod', (), init=0.0, trainable=False)                # networks.py:223,327,424,515
    return ctx.net.lod


def _encoder_grow(ctx, images_in, resolution_log2, min_res_log2, block, fromrgb, lod_in):
    """The recursive `grow` shared by E_zg / E_zl / D_patch
    (networks.py:276-282, 368-374, 568-574), tf.cond -> ctx.cond."""
    def lerp_lod(x, y, t):
        if ctx.mode == 'template':
            return x
        rt = ctx.rt
        a, b = rt.split_unpack(_act_of(x)), rt.split_unpack(_act_of(y))
        out = Act(a.n, a.h, a.w, a.c, f32=_lerp_lod(rt, a.f32, b.f32, t))
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='lerp', a=a, b=b, y=out, t=float(np.float32(t))))
        return T(x.shape, ctx, act=out)

    def grow(res, lod):
        def x_fn():
            return fromrgb(downscale2d(images_in, 2 ** lod), res)
        if lod > 0:
            x = ctx.cond(lod_in < lod, lambda: grow(res + 1, lod - 1), x_fn)
        else:
            x = x_fn()
        x = block(x, res)
        if res > min_res_log2:
            x = ctx.cond(lod_in > lod,
                         lambda: lerp_lod(x, fromrgb(downscale2d(images_in, 2 ** (lod + 1)), res - 1), lod_in - lod),
                         lambda: x)
        return x
    return grow(min_res_log2, resolution_log2 - min_res_log2)


def _tc(ctx, cin, cout, k, up2=False):
    """Will a conv with these dims run on the tensor-core kernel?  (layout hint)"""
    return ctx.mode == 'run' and ctx.rt.choose_algo(cin, cout, k, up2) != 1


# ---------------------------------------------------------------------- E_zg (networks.py:194-291)
def E_zg(images_in, num_channels=3, resolution=128, fmap_base=8192, fmap_decay=1.0, fmap_max=512,
         latent_channels=4, use_wscale=True, use_pixelnorm=False, pixelnorm_epsilon=1e-8, use_leakyrelu=True,
         tanh_at_end=False, dtype='float32', fused_scale=False, structure='recursive', is_template_graph=False,
         **kwargs):
    resolution_log2 = int(np.log2(resolution))
    assert resolution == 2 ** resolution_log2 and resolution >= 4         # networks.py:214
    _check_variants(use_wscale, use_leakyrelu, fused_scale, dtype)
    if tanh_at_end:
        raise NotImplementedError('E_zg tanh_at_end=True is not used by the reference config')

    def nf(stage):
        return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)
    if latent_channels is None:
        latent_channels = nf(0)
    ctx = images_in.ctx
    ctx.pixelnorm = pixelnorm_epsilon if use_pixelnorm else None
    images_in.set_shape([None, num_channels, resolution, resolution])
    lod_in = _lod(ctx)

    def fromrgb(x, res):
        return _fromrgb(x, nf(res - 1), 'FromRGB_lod%d' % (resolution_log2 - res))

    def block(x, res):
        with ctx.variable_scope('%dx%d' % (2 ** res, 2 ** res)):
            if res >= 3:
                with ctx.variable_scope('Conv0'):
                    x = conv2d_layer(x, nf(res - 1), 3, next_tc=_tc(ctx, nf(res - 1), nf(res - 2), 3))
                if fused_scale:
                    with ctx.variable_scope('Conv1_down'):
                        return conv2d_downscale2d_layer(x, nf(res - 2))
                with ctx.variable_scope('Conv1'):
                    x = conv2d_layer(x, nf(res - 2), 3)
                return downscale2d(x)
            with ctx.variable_scope('Conv0'):
                x = conv2d_layer(x, nf(res - 1), 3, next_tc=_tc(ctx, nf(res - 1), nf(res - 2), 3))
            if fused_scale:                                                # networks.py:245-249
                with ctx.variable_scope('zg_Conv1_down'):
                    x = conv2d_downscale2d_layer(x, nf(res - 2))
                with ctx.variable_scope('zg_Conv2_down'):
                    return conv2d_downscale2d_layer(x, latent_channels * 2, gain=1, act=False)
            with ctx.variable_scope('zg_Conv1'):
                x = conv2d_layer(x, nf(res - 2), 3)
            x = downscale2d(x)
            with ctx.variable_scope('zg_Conv2'):
                x = conv2d_layer(x, nf(res - 3), 3)
            x = downscale2d(x)
            with ctx.variable_scope('zg_Conv3'):
                x = conv2d_layer(x, latent_channels * 2, 1, gain=1, act=False)
            return x

    out = _encoder_grow(ctx, images_in, resolution_log2, 2, block, fromrgb, lod_in)
    return _slice_outputs(out, latent_channels, ('zg_mu', 'zg_log_sigma'))


# ---------------------------------------------------------------------- E_zl (networks.py:296-383)
def E_zl(images_in, num_channels=3, resolution=128, fmap_base=8192, fmap_decay=1.0, fmap_max=512, latent_res=4,
         latent_channels=512, use_wscale=True, use_pixelnorm=False, pixelnorm_epsilon=1e-8, use_leakyrelu=True,
         tanh_at_end=False, dtype='float32', fused_scale=False, structure='recursive', is_template_graph=False,
         **kwargs):
    resolution_log2 = int(np.log2(resolution))
    latent_res_log2 = int(np.log2(latent_res))
    assert resolution == 2 ** resolution_log2 and latent_res == 2 ** latent_res_log2 and resolution >= latent_res
    _check_variants(use_wscale, use_leakyrelu, fused_scale, dtype)
    if tanh_at_end:
        raise NotImplementedError('E_zl tanh_at_end=True is not used by the reference config')

    def nf(stage):
        return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)
    if latent_channels is None:
        latent_channels = nf(0)
    ctx = images_in.ctx
    ctx.pixelnorm = pixelnorm_epsilon if use_pixelnorm else None
    images_in.set_shape([None, num_channels, resolution, resolution])
    lod_in = _lod(ctx)

    def fromrgb(x, res):
        return _fromrgb(x, nf(res - 1), 'FromRGB_lod%d' % (resolution_log2 - res))

    def block(x, res):
        with ctx.variable_scope('%dx%d' % (2 ** res, 2 ** res)):
            if res > latent_res_log2:
                with ctx.variable_scope('Conv0'):
                    x = conv2d_layer(x, nf(res - 1), 3, next_tc=_tc(ctx, nf(res - 1), nf(res - 2), 3))
                if fused_scale:
                    with ctx.variable_scope('Conv1_down'):
                        return conv2d_downscale2d_layer(x, nf(res - 2))
                with ctx.variable_scope('Conv1'):
                    x = conv2d_layer(x, nf(res - 2), 3)
                return downscale2d(x)
            with ctx.variable_scope('Conv0'):
                x = conv2d_layer(x, nf(res - 1), 3, next_tc=_tc(ctx, nf(res - 1), latent_channels * 2, 1))
            with ctx.variable_scope('z_Conv1'):
                x = conv2d_layer(x, latent_channels * 2, 1, gain=1, act=False)
            return x

    out = _encoder_grow(ctx, images_in, resolution_log2, latent_res_log2, block, fromrgb, lod_in)
    return _slice_outputs(out, latent_channels, ('z_mu', 'z_log_sigma'))


# ---------------------------------------------------------------------- G_res (networks.py:388-486)
def G_res(zg_latents_in, zl_latents_in, num_channels=3, resolution=128, fmap_base=8192, fmap_decay=1.0,
          fmap_max=512, latent_res=4, latent_channels=512, use_wscale=True, use_pixelnorm=False,
          pixelnorm_epsilon=1e-8, use_leakyrelu=True, tanh_at_end=False, dtype='float32', fused_scale=False,
          structure='recursive', is_template_graph=False, scale_h=1, scale_w=1, tail_window=None, mid_window=None,
          **kwargs):
    """`tail_window` = (oy, ox, h, w) in latent pixels (not a reference argument): after the latent-resolution block
    only that window of the feature map goes on through the up-sampling blocks and the image heads - the output is
    the corresponding [4h, 4w] part of the image.  Used by the crop-aware train step (loss.tail_window): the layers
    above the trunk see 2 latent pixels of context around the crop instead of the trunk's 14.  `mid_window` does
    the same after the fourth residual block (loss.mid_window; `tail_window` is then relative to it)."""
    resolution_log2 = int(np.log2(resolution))
    latent_res_log2 = int(np.log2(latent_res))
    assert resolution == 2 ** resolution_log2 and latent_res == 2 ** latent_res_log2 and resolution >= latent_res
    _check_variants(use_wscale, use_leakyrelu, fused_scale, dtype)

    def nf(stage):
        return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)
    if latent_channels is None:
        latent_channels = nf(0)
    ctx = zg_latents_in.ctx
    ctx.pixelnorm = pixelnorm_epsilon if use_pixelnorm else None
    # inference convenience (not in the reference): a [N,C,1,1] global code is tiled over the canvas ON THE DEVICE -
    # what every caller does on the host first (run.py:375 np.tile, loss.py:130 tf.tile), at half the H2D bytes
    zg_bcast = ctx.mode == 'run' and ctx.tape is None and list(zg_latents_in.shape[2:]) == [1, 1] and \
        list(zl_latents_in.shape[2:]) != [1, 1]
    if not zg_bcast:
        zg_latents_in.set_shape([None, latent_channels, latent_res * scale_h, latent_res * scale_w])
    zl_latents_in.set_shape([None, latent_channels, latent_res * scale_h, latent_res * scale_w])
    c2 = latent_channels * 2

    # combo_in = concat([zg, zl], axis=1) (networks.py:423): two NCHW -> NHWC slice moves
    combo_shape = [zg_latents_in.shape[0], c2, latent_res * scale_h, latent_res * scale_w]
    if ctx.mode == 'template':
        combo_in = T(combo_shape, ctx)
    else:
        rt = ctx.rt
        n, _, h, w = zl_latents_in.nchw.shape
        if tuple(zg_latents_in.nchw.shape) != ((n, latent_channels, 1, 1) if zg_bcast else (n, latent_channels, h, w)):
            raise ValueError('G_res: zg %s and zl %s disagree' % (tuple(zg_latents_in.nchw.shape),
                                                                  tuple(zl_latents_in.nchw.shape)))
        buf = rt.empty(n, h, w, c2)
        rt.nchw_to_nhwc(zg_latents_in.nchw, out=buf, c_off=0, c_total=c2, bcast_hw=(h, w) if zg_bcast else None)
        rt.nchw_to_nhwc(zl_latents_in.nchw, out=buf, c_off=latent_channels, c_total=c2)
        combo_shape[0] = n
        combo_in = T(combo_shape, ctx, act=Act(n, h, w, c2, f32=buf))
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='concat', y=combo_in.act, inputs=[zg_latents_in.nchw, zl_latents_in.nchw],
                                 c=latent_channels))
    lod_in = _lod(ctx)

    def block(x, res):
        with ctx.variable_scope('%dx%d' % (2 ** res, 2 ** res)):
            if res == latent_res_log2:
                for count in range(5):                                     # networks.py:430-437
                    x0 = x
                    with ctx.variable_scope('Residual%d_0' % count):
                        x = conv2d_layer(x, c2, 3, next_tc=_tc(ctx, c2, c2, 3))
                    with ctx.variable_scope('Residual%d_1' % count):
                        nxt = c2 if count < 4 else nf(res - 1)
                        # the sum is also the next block's x0: keep the exact fp32 copy beside the planes
                        x = conv2d_layer(x, c2, 3, gain=1, act=False, residual=x0, next_tc=_tc(ctx, c2, nxt, 3),
                                         keep_f32=count < 4)
                    if count == 3 and mid_window is not None and ctx.mode == 'run':
                        x = _window(x, mid_window)      # crop-aware: 4 + 2 convs of context left -> a smaller window
                with ctx.variable_scope('Conv0'):
                    x = conv2d_layer(x, nf(res - 1), 3, gain=SQRT2 / 4, next_tc=_tc(ctx, nf(res - 1), nf(res - 1), 3))
                with ctx.variable_scope('Conv1'):
                    x = last_conv(x, res)
            elif fused_scale:
                with ctx.variable_scope('Conv0_up'):                       # networks.py:444-446
                    x = upscale2d_conv2d_layer(x, nf(res - 1))
                with ctx.variable_scope('Conv1'):
                    x = last_conv(x, res)
            else:
                with ctx.variable_scope('Conv0'):                          # upscale2d + conv2d (networks.py:448-450)
                    x = conv2d_layer(x, nf(res - 1), 3, up2=True, next_tc=_tc(ctx, nf(res - 1), nf(res - 1), 3))
                with ctx.variable_scope('Conv1'):
                    x = last_conv(x, res)
            return x

    def last_conv(x, res):
        """Conv1 of a block: its consumer is either the next block's upscale2d+Conv0
        (hand over REPLICATE-halo planes) or, at the output resolution with lod == 0,
        the ToRGB head + tanh (fused into the epilogue)."""
        if res < resolution_log2:
            ntc = _tc(ctx, nf(res - 1), nf(res), 3, up2=True)
            # a windowed output (crop-aware step) is cut from the exact fp32 map: re-splitting hi + lo is not
            # idempotent at rounding ties, and window / whole-canvas evaluations must stay bit-identical
            return conv2d_layer(x, nf(res - 1), 3, next_tc=ntc, next_up2=ntc,
                                keep_f32=res == latent_res_log2 and tail_window is not None)
        head = None
        if ctx.mode == 'run' and lod_in == 0 and not use_pixelnorm and nf(res - 1) in (16, 32) and \
                _tc(ctx, nf(res - 1), nf(res - 1), 3):
            head = ('ToRGB_lod0', num_channels, bool(tanh_at_end))
        return conv2d_layer(x, nf(res - 1), 3, torgb=head)

    def torgb(x, res, apply_tanh=False):
        lod = resolution_log2 - res
        with ctx.variable_scope('ToRGB_lod%d' % lod):
            cin = x.shape[1]
            w = ctx.get_variable('weight', (1, 1, cin, num_channels))
            b = ctx.get_variable('bias', (num_channels,), init='zeros')
            shape = [x.shape[0], num_channels, x.shape[2], x.shape[3]]
            if ctx.mode == 'template':
                return T(shape, ctx)
            if x.rgb is not None and x.rgb[0] == 'ToRGB_lod%d' % lod and x.rgb[1] == bool(apply_tanh):
                out = T(shape, ctx, nchw=x.rgb[2])                         # produced by the conv epilogue
            else:
                img = ctx.rt.torgb(_act_of(x), w.value, b.value, _wscale(w.shape, 1.0), num_channels, apply_tanh)
                if ctx.tape is not None:
                    ctx.tape.append(dict(kind='torgb', x=_act_of(x), img=img, w=w.name, b=b.name,
                                         wscale=_wscale(w.shape, 1.0), tanh=bool(apply_tanh)))
                out = T(shape, ctx, nchw=img)
            out.tanh_done = bool(apply_tanh)
            return out

    def up_img(t, factor):
        if factor == 1:
            return t
        shape = [t.shape[0], t.shape[1], _mul(t.shape[2], factor), _mul(t.shape[3], factor)]
        if ctx.mode == 'template':
            return T(shape, ctx)
        up = _upscale_image(ctx.rt, t.nchw, factor)
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='imgup', x=t.nchw, y=up, factor=factor))
        return T(shape, ctx, nchw=up)

    def lerp_lod(a, b, t):
        if ctx.mode == 'template':
            return a
        out = _lerp_lod(ctx.rt, a.nchw, b.nchw, t)
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='imglerp', a=a.nchw, b=b.nchw, y=out, t=float(np.float32(t))))
        return T(a.shape, ctx, nchw=out)

    def grow(x, res, lod):                                                 # networks.py:473-479
        y = block(x, res)
        if res == latent_res_log2 and tail_window is not None and ctx.mode == 'run':
            y = _window(y, tail_window)         # crop-aware evaluation: the blocks above only see what the crop needs

        def img_fn():
            return up_img(torgb(y, res, apply_tanh=tanh_at_end and lod == 0), 2 ** lod)
        img = img_fn
        if res > latent_res_log2:
            prev = img

            def img_fade():
                return ctx.cond(lod_in > lod,
                                lambda: up_img(lerp_lod(torgb(y, res), up_img(torgb(x, res - 1), 2), lod_in - lod),
                                               2 ** lod),
                                prev)
            img = img_fade
        if lod > 0:
            prev2 = img

            def img_deeper():
                return ctx.cond(lod_in < lod, lambda: grow(y, res + 1, lod - 1), prev2)
            img = img_deeper
        return img()

    images_out = grow(combo_in, latent_res_log2, resolution_log2 - latent_res_log2)
    if ctx.mode == 'run' and tanh_at_end and not getattr(images_out, 'tanh_done', False):
        # lod != 0: tf.nn.tanh follows the fade / upscale of the lower-resolution heads (networks.py:482-483)
        pre = images_out.nchw
        images_out = T(images_out.shape, ctx, nchw=_tanh(ctx.rt, pre))
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='tanh', x=pre, y=images_out.nchw))
    images_out.name = 'images_out'
    return images_out


# ---------------------------------------------------------------------- D_patch (networks.py:491-577)
def D_patch(images_in, num_channels=3, resolution=128, fmap_base=8192, fmap_decay=1.0, fmap_max=512,
            latent_res=4, use_wscale=True, mbstd_group_size=4, dtype='float32', fused_scale=False,
            structure='recursive', is_template_graph=False, **kwargs):
    resolution_log2 = int(np.log2(resolution))
    latent_res_log2 = 2 if latent_res == -1 else int(np.log2(latent_res))
    _check_variants(use_wscale, True, fused_scale, dtype)

    def nf(stage):
        return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)
    ctx = images_in.ctx
    images_in.set_shape([None, num_channels, resolution, resolution])
    lod_in = _lod(ctx)

    def fromrgb(x, res):
        return _fromrgb(x, nf(res - 1), 'FromRGB_lod%d' % (resolution_log2 - res))

    def block(x, res):
        with ctx.variable_scope('%dx%d' % (2 ** res, 2 ** res)):
            if res > latent_res_log2:
                with ctx.variable_scope('Conv0'):
                    x = conv2d_layer(x, nf(res - 1), 3, next_tc=_tc(ctx, nf(res - 1), nf(res - 2), 3))
                if fused_scale:
                    with ctx.variable_scope('Conv1_down'):
                        return conv2d_downscale2d_layer(x, nf(res - 2))
                with ctx.variable_scope('Conv1'):
                    x = conv2d_layer(x, nf(res - 2), 3)
                return downscale2d(x)
            if mbstd_group_size > 1:
                x = _minibatch_stddev_layer(x, mbstd_group_size)
            with ctx.variable_scope('Conv0'):
                x = conv2d_layer(x, nf(res - 1), 3)
            if latent_res == -1:
                with ctx.variable_scope('Dense1'):
                    x = _dense_layer(x, nf(res - 2), act=True)
                with ctx.variable_scope('Dense2'):
                    x = _dense_layer(x, 1, gain=1, act=False)
                if ctx.mode == 'run':
                    x.nchw = x.nchw.view(x.nchw.shape[0], 1, 1, 1)
                    if ctx.tape is not None:
                        ctx.tape[-1]['alias'] = x.nchw
                x.shape = [x.shape[0], 1, 1, 1]
            else:
                with ctx.variable_scope('Conv1'):
                    x = conv2d_layer(x, nf(res - 2), 1)
                with ctx.variable_scope('Conv2'):
                    x = conv2d_layer(x, 1, 1, gain=1, act=False)
            return x

    scores_out = _encoder_grow(ctx, images_in, resolution_log2, latent_res_log2, block, fromrgb, lod_in)
    if ctx.mode == 'run' and scores_out.nchw is None:
        a = ctx.rt.split_unpack(_act_of(scores_out))
        scores_out.nchw = ctx.rt.nhwc_to_nchw(a.f32)
    scores_out.name = 'scores_out'
    return scores_out


def _minibatch_stddev_layer(x, group_size):
    """networks.py:177-189: one extra channel holding the group-of-G stddev statistic."""
    ctx = x.ctx
    shape = [x.shape[0], x.shape[1] + 1, x.shape[2], x.shape[3]]
    if ctx.mode == 'template':
        return T(shape, ctx)
    xa = _act_of(x)
    out, _ = ctx.rt.mbstd(xa, group_size)
    if ctx.tape is not None:
        ctx.tape.append(dict(kind='mbstd', x=xa, y=out, group=group_size))
    return T(shape, ctx, act=out)


def _dense_layer(x, fmaps, gain=SQRT2, act=True):
    """act(apply_bias(dense(x))) (networks.py:38-43, 61-67): rows are the NCHW
    flattening of x, the variable is [in, out]."""
    ctx = x.ctx
    fan_in = int(np.prod(x.shape[1:]))
    w = ctx.get_variable('weight', (fan_in, fmaps))
    b = ctx.get_variable('bias', (fmaps,), init='zeros')
    shape = [x.shape[0], fmaps]
    if ctx.mode == 'template':
        return T(shape, ctx)
    rt = ctx.rt
    if x.nchw is None:
        a = rt.split_unpack(_act_of(x))
        x.nchw = rt.nhwc_to_nchw(a.f32)
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='flatten', x=a, y=x.nchw))
    flat = x.nchw.view(x.nchw.shape[0], -1)
    out = rt.dense(flat, w.value, b.value, _wscale(w.shape, gain), lrelu=act)
    if ctx.tape is not None:
        ctx.tape.append(dict(kind='dense', x=x.nchw, y=out, w=w.name, b=b.name, wscale=_wscale(w.shape, gain), act=act))
    return T(shape, ctx, nchw=out)


# ---------------------------------------------------------------------- VGG-19 features (custom_vgg19.py:20-66)
VGG19_LAYERS = (('conv1_1', 64), ('conv1_2', 64), ('pool1', 0), ('conv2_1', 128), ('conv2_2', 128), ('pool2', 0),
                ('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256), ('conv3_4', 256), ('pool3', 0),
                ('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512), ('conv4_4', 512), ('pool4', 0), ('conv5_1', 512))
VGG19_GRAM_LAYERS = ('conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1')                # loss.py:153


def Vgg19_features(images_in, num_channels=3, resolution=128, dtype='float32', is_template_graph=False, gram_sink=None,
                   **kwargs):
    """The feature extractor of the Gram loss: `custom_Vgg19` (custom_vgg19.py:20-66) up to conv5_1 - input scaling to
    [0,255] BGR minus the VGG mean, then tensorflow_vgg's conv_layer = relu(conv2d SAME + bias) and 2x2 average
    pooling - returning the five activations whose Gram matrices the loss compares (loss.py:153), NCHW.
    Variables `<layer>/weight` [3,3,Cin,Cout] and `<layer>/bias` are non-trainable constants (vgg19.npy, see
    texturemixer_b200.vgg.load_vgg19_npy).  (The reference's networks.Vgg19_gram_autocorrelation, networks.py:582-600,
    wraps the same extractor; its autocorrelation output is used by no loss in the tree.)
    `gram_sink` (a list; not a reference argument): receives the five activations in their internal form (split-bf16
    planes, NHWC) for the tensor-core Gram kernels and the NCHW outputs are not materialised (vgg.GramLoss)."""
    ctx = images_in.ctx
    ctx.pixelnorm = None
    images_in.set_shape([None, num_channels, resolution, resolution])
    assert num_channels == 3
    if ctx.mode == 'template':
        x = T([None, 3, resolution, resolution], ctx)
    else:
        import ctypes as C
        from . import _lib
        rt = ctx.rt
        img = images_in.nchw
        n, _, h, w = img.shape
        pre = Act(n, h, w, 16, f32=rt.empty(n, h, w, 16))
        _lib.check(rt.lib.tmx_vgg_preprocess(rt.handle, C.c_void_p(img.data_ptr()), C.c_void_p(pre.f32.data_ptr()), n, h,
                                             w, rt.stream()), 'tmx_vgg_preprocess')
        if ctx.tape is not None:
            ctx.tape.append(dict(kind='vggpre', img=img, y=pre))
        x = T([n, 3, h, w], ctx, act=pre)
    outs = []
    for i, (name, fmaps) in enumerate(VGG19_LAYERS):
        if name.startswith('pool'):
            x = downscale2d(x)                                             # avg_pool 2x2 (custom_vgg19.py:44)
            continue
        nxt = VGG19_LAYERS[i + 1][0] if i + 1 < len(VGG19_LAYERS) else 'end'
        is_gram = name in VGG19_GRAM_LAYERS
        internal = gram_sink is not None and ctx.mode != 'template'
        with ctx.variable_scope(name):
            x = conv2d_layer(x, fmaps, 3, act=True, pad='zero', alpha=0.0, wscale=1.0, trainable=False,
                             next_tc=nxt.startswith('conv') or (is_gram and internal), next_pad='zero',
                             keep_f32=is_gram and not internal)
        if is_gram and internal:
            gram_sink.append(x.act)
            outs.append(T(list(x.shape), ctx, name=name))
        elif is_gram:
            (o,) = _slice_outputs(x, fmaps, (name,))
            outs.append(o)
    return tuple(outs)


# ---------------------------------------------------------------------- north_star aliases
def build_encoder(images_in, kind='zl', **kwargs):
    """`build_encoder(kind='zg'|'zl')` == E_zg / E_zl."""
    return {'zg': E_zg, 'zl': E_zl}[kind](images_in, **kwargs)


def build_generator(zg_latents_in, zl_latents_in, **kwargs):
    """`build_generator` == G_res."""
    return G_res(zg_latents_in, zl_latents_in, **kwargs)
