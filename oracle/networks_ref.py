"""ORACLE (test infrastructure, never shipped or measured as the product).

CPU restatement, in torch fp32/fp64 eager mode, of the TextureMixer hot-path
networks.  Every function cites the reference lines it follows
(`/root/reference/networks.py` unless noted).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module.

Parity status: the reference ships no golden vectors and TensorFlow 1.12 is
not installable here, so the *op-level* semantics (cross-correlation conv,
REFLECT pad, VALID avg-pool) are restated from the TF documentation
("parity unpinned" at op level).  The *network structure* (layer order, gains,
channel counts, variable names, creation order) IS pinned against the
reference's own `networks.py`, executed unmodified on top of `oracle/tfshim`
(see `tests/golden/make_golden.py` and `tests/test_oracle_vs_reference.py`).

Layout: activations NCHW, conv weights HWIO `[k,k,Cin,Cout]`, dense weights
`[in,out]` — the reference's own layouts (networks.py:26-56).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = float(np.sqrt(2))


# ----------------------------------------------------------------------------
# Layer primitives (networks.py:19-189)

def lerp(a, b, t):
    """networks.py:19 / tfutil.py:41-43."""
    return a + (b - a) * t


def wscale_of(shape, gain=SQRT2, fan_in=None):
    """Runtime equalised-lr scale, networks.py:26-33 (use_wscale=True):
    std = gain / sqrt(fan_in), fan_in = prod(shape[:-1]); cast to float32."""
    if fan_in is None:
        fan_in = np.prod(shape[:-1])
    return np.float32(gain / np.sqrt(fan_in))


def conv2d(x, w, gain=SQRT2):
    """networks.py:48-56.  Cross-correlation, stride 1; k=1 VALID, k=3 REFLECT
    pad by k//2 then VALID.  `w` is the raw N(0,1) variable [k,k,Cin,Cout]; the
    wscale multiply happens here like get_weight (networks.py:30-31)."""
    k = w.shape[0]
    assert k >= 1 and k % 2 == 1
    ws = float(wscale_of(tuple(w.shape), gain))
    wt = (w * ws).permute(3, 2, 0, 1)  # HWIO -> OIHW (torch conv2d is cross-correlation too)
    if k > 1:
        x = F.pad(x, (k // 2, k // 2, k // 2, k // 2), mode='reflect')
    return F.conv2d(x, wt)


def dense(x, w, gain=SQRT2):
    """networks.py:38-43: flatten NCHW row-major, matmul with [in,out]."""
    if x.dim() > 2:
        x = x.reshape(x.shape[0], -1)
    ws = float(wscale_of(tuple(w.shape), gain))
    return x @ (w * ws)


def apply_bias(x, b):
    """networks.py:61-67."""
    if x.dim() == 2:
        return x + b
    return x + b.reshape(1, -1, 1, 1)


def leaky_relu(x, alpha=0.2):
    """networks.py:72-75: max(x*alpha, x)."""
    return torch.maximum(x * alpha, x)


def upscale2d(x, factor=2):
    """networks.py:80-88: nearest-neighbour, y[i,j] = x[i//f, j//f]."""
    if factor == 1:
        return x
    n, c, h, w = x.shape
    x = x.reshape(n, c, h, 1, w, 1).expand(n, c, h, factor, w, factor)
    return x.reshape(n, c, h * factor, w * factor)


def downscale2d(x, factor=2):
    """networks.py:131-136: VALID average pool, ksize=stride=factor."""
    if factor == 1:
        return x
    return F.avg_pool2d(x, factor, factor)


def _fuse4(w):
    """networks.py:97-98,145-146: zero-pad the 3x3 kernel to 5x5 and add its four one-pixel shifts -> 4x4."""
    w = F.pad(w, (0, 0, 0, 0, 1, 1, 1, 1))
    return w[1:, 1:] + w[:-1, 1:] + w[1:, :-1] + w[:-1, :-1]


def upscale2d_conv2d(x, w, gain=SQRT2):
    """networks.py:94-101 (fused_scale=True): conv2d_transpose, stride 2, SAME, with the fused 4x4 kernel.
    Variable [k, k, fmaps, Cin]; wscale from fan_in = k*k*Cin."""
    k, _, fmaps, cin = w.shape
    w = _fuse4(w * wscale_of(w.shape, gain, fan_in=k * k * cin))
    return F.conv_transpose2d(x, w.permute(3, 2, 0, 1), stride=2, padding=1)


def conv2d_downscale2d(x, w, gain=SQRT2):
    """networks.py:142-148 (fused_scale=True): conv2d, stride 2, SAME, with the fused 4x4 kernel * 0.25."""
    w = _fuse4(w * wscale_of(w.shape, gain)) * 0.25
    return F.conv2d(F.pad(x, (1, 1, 1, 1)), w.permute(3, 2, 0, 1), stride=2)


def pixel_norm(x, epsilon=1e-8):
    """networks.py:170-172 (off in config.py:77-79; kept for completeness)."""
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + epsilon)


def minibatch_stddev_layer(x, group_size=4):
    """networks.py:177-189.  reshape(x,[G,-1,C,H,W]) puts sample n = g*M + m in
    group slot g of group m."""
    n, c, h, w = x.shape
    g = min(group_size, n)
    y = x.reshape(g, -1, c, h, w)
    y = y - y.mean(dim=0, keepdim=True)
    y = (y * y).mean(dim=0)
    y = torch.sqrt(y + 1e-8)
    y = y.mean(dim=(1, 2, 3), keepdim=True)           # [M,1,1,1]
    y = y.repeat(g, 1, h, w)                           # tf.tile -> sample g*M+m gets y[m]
    return torch.cat([x, y], dim=1)


# ----------------------------------------------------------------------------
# Parameter tables.  Creation order == TF variable creation order under the
# 'recursive' structure with tf.cond building both branches (true_fn first):
# verified against the reference code in tests/test_oracle_vs_reference.py.

def nf_fn(fmap_base=1024, fmap_decay=1.0, fmap_max=512):
    def nf(stage):
        return min(int(fmap_base / (2.0 ** (stage * fmap_decay))), fmap_max)
    return nf


def _add_conv(spec, scope, k, cin, cout, gain):
    spec[scope + '/weight'] = dict(shape=(k, k, cin, cout), gain=gain)
    spec[scope + '/bias'] = dict(shape=(cout,))


def _add_dense(spec, scope, cin, cout, gain):
    spec[scope + '/weight'] = dict(shape=(cin, cout), gain=gain)
    spec[scope + '/bias'] = dict(shape=(cout,))


def spec_E_zg(num_channels=3, resolution=128, fmap_base=1024, fmap_decay=1.0, fmap_max=512,
              latent_channels=128, fused_scale=False, **_):
    """Variable table of E_zg (networks.py:194-291)."""
    c1 = 'Conv1_down' if fused_scale else 'Conv1'
    nf = nf_fn(fmap_base, fmap_decay, fmap_max)
    rl2 = int(np.log2(resolution))
    spec = OrderedDict(lod=dict(shape=()))
    for res in range(rl2, 2, -1):
        _add_conv(spec, 'FromRGB_lod%d' % (rl2 - res), 1, num_channels, nf(res - 1), SQRT2)
        if res == rl2:
            pass
        s = '%dx%d' % (2 ** res, 2 ** res)
        _add_conv(spec, s + '/Conv0', 3, nf(res - 1), nf(res - 1), SQRT2)
        _add_conv(spec, s + '/' + c1, 3, nf(res - 1), nf(res - 2), SQRT2)
    # the deepest recursion level creates FromRGB for res=2 last-but-structure:
    _add_conv(spec, 'FromRGB_lod%d' % (rl2 - 2), 1, num_channels, nf(1), SQRT2)
    _add_conv(spec, '4x4/Conv0', 3, nf(1), nf(1), SQRT2)
    if fused_scale:                                             # networks.py:245-249
        _add_conv(spec, '4x4/zg_Conv1_down', 3, nf(1), nf(0), SQRT2)
        _add_conv(spec, '4x4/zg_Conv2_down', 3, nf(0), latent_channels * 2, 1.0)
        return spec
    _add_conv(spec, '4x4/zg_Conv1', 3, nf(1), nf(0), SQRT2)
    _add_conv(spec, '4x4/zg_Conv2', 3, nf(0), nf(-1), SQRT2)
    _add_conv(spec, '4x4/zg_Conv3', 1, nf(-1), latent_channels * 2, 1.0)
    return spec


def spec_E_zl(num_channels=3, resolution=128, fmap_base=1024, fmap_decay=1.0, fmap_max=512,
              latent_res=32, latent_channels=128, fused_scale=False, **_):
    """Variable table of E_zl (networks.py:296-383)."""
    c1 = 'Conv1_down' if fused_scale else 'Conv1'
    nf = nf_fn(fmap_base, fmap_decay, fmap_max)
    rl2 = int(np.log2(resolution))
    ll2 = int(np.log2(latent_res))
    spec = OrderedDict(lod=dict(shape=()))
    for res in range(rl2, ll2, -1):
        _add_conv(spec, 'FromRGB_lod%d' % (rl2 - res), 1, num_channels, nf(res - 1), SQRT2)
        s = '%dx%d' % (2 ** res, 2 ** res)
        _add_conv(spec, s + '/Conv0', 3, nf(res - 1), nf(res - 1), SQRT2)
        _add_conv(spec, s + '/' + c1, 3, nf(res - 1), nf(res - 2), SQRT2)
    _add_conv(spec, 'FromRGB_lod%d' % (rl2 - ll2), 1, num_channels, nf(ll2 - 1), SQRT2)
    s = '%dx%d' % (2 ** ll2, 2 ** ll2)
    _add_conv(spec, s + '/Conv0', 3, nf(ll2 - 1), nf(ll2 - 1), SQRT2)
    _add_conv(spec, s + '/z_Conv1', 1, nf(ll2 - 1), latent_channels * 2, 1.0)
    return spec


def spec_G_res(num_channels=3, resolution=128, fmap_base=1024, fmap_decay=1.0, fmap_max=512,
               latent_res=32, latent_channels=128, fused_scale=False, **_):
    """Variable table of G_res (networks.py:388-486)."""
    nf = nf_fn(fmap_base, fmap_decay, fmap_max)
    rl2 = int(np.log2(resolution))
    ll2 = int(np.log2(latent_res))
    spec = OrderedDict(lod=dict(shape=()))
    c = latent_channels * 2
    s = '%dx%d' % (2 ** ll2, 2 ** ll2)
    for i in range(5):
        _add_conv(spec, s + '/Residual%d_0' % i, 3, c, c, SQRT2)
        _add_conv(spec, s + '/Residual%d_1' % i, 3, c, c, 1.0)
    _add_conv(spec, s + '/Conv0', 3, c, nf(ll2 - 1), SQRT2 / 4)
    _add_conv(spec, s + '/Conv1', 3, nf(ll2 - 1), nf(ll2 - 1), SQRT2)
    for res in range(ll2 + 1, rl2 + 1):
        s = '%dx%d' % (2 ** res, 2 ** res)
        if fused_scale:                  # networks.py:95: the transposed-conv variable is [k, k, fmaps, Cin]
            spec[s + '/Conv0_up/weight'] = dict(shape=(3, 3, nf(res - 1), nf(res - 2)), gain=SQRT2)
            spec[s + '/Conv0_up/bias'] = dict(shape=(nf(res - 1),))
        else:
            _add_conv(spec, s + '/Conv0', 3, nf(res - 2), nf(res - 1), SQRT2)
        _add_conv(spec, s + '/Conv1', 3, nf(res - 1), nf(res - 1), SQRT2)
    # ToRGB heads are created on the way back out of the recursion: lod0 first.
    for res in range(rl2, ll2 - 1, -1):
        _add_conv(spec, 'ToRGB_lod%d' % (rl2 - res), 1, nf(res - 1), num_channels, 1.0)
    return spec


def spec_D_patch(num_channels=3, resolution=128, fmap_base=1024, fmap_decay=1.0, fmap_max=512,
                 latent_res=-1, mbstd_group_size=4, fused_scale=False, **_):
    """Variable table of D_patch (networks.py:491-577)."""
    c1 = 'Conv1_down' if fused_scale else 'Conv1'
    nf = nf_fn(fmap_base, fmap_decay, fmap_max)
    rl2 = int(np.log2(resolution))
    ll2 = 2 if latent_res == -1 else int(np.log2(latent_res))
    spec = OrderedDict(lod=dict(shape=()))
    for res in range(rl2, ll2, -1):
        _add_conv(spec, 'FromRGB_lod%d' % (rl2 - res), 1, num_channels, nf(res - 1), SQRT2)
        s = '%dx%d' % (2 ** res, 2 ** res)
        _add_conv(spec, s + '/Conv0', 3, nf(res - 1), nf(res - 1), SQRT2)
        _add_conv(spec, s + '/' + c1, 3, nf(res - 1), nf(res - 2), SQRT2)
    _add_conv(spec, 'FromRGB_lod%d' % (rl2 - ll2), 1, num_channels, nf(ll2 - 1), SQRT2)
    s = '%dx%d' % (2 ** ll2, 2 ** ll2)
    cin = nf(ll2 - 1) + (1 if mbstd_group_size > 1 else 0)
    _add_conv(spec, s + '/Conv0', 3, cin, nf(ll2 - 1), SQRT2)
    if latent_res == -1:
        _add_dense(spec, s + '/Dense1', nf(ll2 - 1) * (2 ** ll2) ** 2, nf(ll2 - 2), SQRT2)
        _add_dense(spec, s + '/Dense2', nf(ll2 - 2), 1, 1.0)
    else:
        _add_conv(spec, s + '/Conv1', 1, nf(ll2 - 1), nf(ll2 - 2), SQRT2)
        _add_conv(spec, s + '/Conv2', 1, nf(ll2 - 2), 1, 1.0)
    return spec


SPECS = dict(E_zg=spec_E_zg, E_zl=spec_E_zl, G_res=spec_G_res, D_patch=spec_D_patch)


def init_params(func, rng, bias_scale=0.1, **cfg):
    """Random parameters in creation order.  Weights ~N(0,1) like the reference
    initialiser (networks.py:31); biases 0.1*N(0,1) instead of the reference's
    zeros (networks.py:62) so that bias bugs cannot hide."""
    out = OrderedDict()
    for name, d in SPECS[func](**cfg).items():
        if name == 'lod':
            out[name] = np.float32(0.0)
        elif name.endswith('/bias'):
            out[name] = (bias_scale * rng.randn(*d['shape'])).astype(np.float32)
        else:
            out[name] = rng.randn(*d['shape']).astype(np.float32)
    return out


def to_torch(params, dtype=torch.float32, requires_grad=False):
    out = OrderedDict()
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v)).to(dtype)
        if requires_grad and k != 'lod':
            t.requires_grad_(True)
        out[k] = t
    return out


# ----------------------------------------------------------------------------
# Networks.  `P` maps reference variable names to torch tensors; `taps`, when a
# dict, receives every named intermediate (pre-activation too) for per-layer
# parity checks.

def _layer(P, scope, x, gain=SQRT2, act=True, taps=None, pn=None, op=None):
    """[PN](act(apply_bias(conv2d(x)))): `pn` = pixel_norm epsilon of the blocks' activated convs when
    use_pixelnorm (networks.py:216,320,415), None otherwise; `op` = the fused up/down-scaling conv instead."""
    x = apply_bias((op or conv2d)(x, P[scope + '/weight'], gain), P[scope + '/bias'])
    if taps is not None:
        taps[scope + ':pre'] = x
    if act:
        x = leaky_relu(x)
        if pn is not None:
            x = pixel_norm(x, pn)
    return x


def _encoder_trunk(images_in, P, resolution, min_res_log2, block, fromrgb, lod_in):
    """The shared recursive `grow` of E_zg / E_zl / D_patch
    (networks.py:276-282, 368-374, 568-574)."""
    rl2 = int(np.log2(resolution))

    def grow(res, lod):
        if lod > 0 and lod_in < lod:
            x = grow(res + 1, lod - 1)
        else:
            x = fromrgb(downscale2d(images_in, 2 ** lod), res)
        x = block(x, res)
        if res > min_res_log2 and lod_in > lod:
            x = lerp(x, fromrgb(downscale2d(images_in, 2 ** (lod + 1)), res - 1), lod_in - lod)
        return x
    return grow(min_res_log2, rl2 - min_res_log2)


def E_zg(images_in, P, num_channels=3, resolution=128, fmap_base=1024, fmap_decay=1.0, fmap_max=512,
         latent_channels=128, tanh_at_end=False, taps=None, use_pixelnorm=False, pixelnorm_epsilon=1e-8,
         fused_scale=False, **_):
    """networks.py:194-291 -> (zg_mu, zg_log_sigma), each [N,latent_channels,1,1]."""
    pn = pixelnorm_epsilon if use_pixelnorm else None
    rl2 = int(np.log2(resolution))
    assert resolution == 2 ** rl2 and resolution >= 4
    assert tuple(images_in.shape[1:]) == (num_channels, resolution, resolution)
    lod_in = float(P['lod'])

    def fromrgb(x, res):
        return _layer(P, 'FromRGB_lod%d' % (rl2 - res), x, taps=taps)

    def block(x, res):
        s = '%dx%d' % (2 ** res, 2 ** res)
        if res >= 3:
            x = _layer(P, s + '/Conv0', x, taps=taps, pn=pn)
            if fused_scale:
                return _layer(P, s + '/Conv1_down', x, taps=taps, pn=pn, op=conv2d_downscale2d)
            x = _layer(P, s + '/Conv1', x, taps=taps, pn=pn)
            return downscale2d(x)
        x = _layer(P, s + '/Conv0', x, taps=taps, pn=pn)
        if fused_scale:                                                 # networks.py:245-249
            x = _layer(P, s + '/zg_Conv1_down', x, taps=taps, pn=pn, op=conv2d_downscale2d)
            return _layer(P, s + '/zg_Conv2_down', x, gain=1.0, act=False, taps=taps, op=conv2d_downscale2d)
        x = downscale2d(_layer(P, s + '/zg_Conv1', x, taps=taps, pn=pn))
        x = downscale2d(_layer(P, s + '/zg_Conv2', x, taps=taps, pn=pn))
        return _layer(P, s + '/zg_Conv3', x, gain=1.0, act=False, taps=taps, pn=pn)

    out = _encoder_trunk(images_in, P, resolution, 2, block, fromrgb, lod_in)
    if tanh_at_end:
        out = torch.tanh(out)
    return out[:, :latent_channels], out[:, latent_channels:]


def E_zl(images_in, P, num_channels=3, resolution=128, fmap_base=1024, fmap_decay=1.0, fmap_max=512,
         latent_res=32, latent_channels=128, tanh_at_end=False, taps=None, use_pixelnorm=False,
         pixelnorm_epsilon=1e-8, fused_scale=False, **_):
    """networks.py:296-383 -> (z_mu, z_log_sigma), each [N,latent_channels,latent_res,latent_res]."""
    pn = pixelnorm_epsilon if use_pixelnorm else None
    rl2 = int(np.log2(resolution))
    ll2 = int(np.log2(latent_res))
    assert resolution == 2 ** rl2 and latent_res == 2 ** ll2 and resolution >= latent_res
    assert tuple(images_in.shape[1:]) == (num_channels, resolution, resolution)
    lod_in = float(P['lod'])

    def fromrgb(x, res):
        return _layer(P, 'FromRGB_lod%d' % (rl2 - res), x, taps=taps)

    def block(x, res):
        s = '%dx%d' % (2 ** res, 2 ** res)
        if res > ll2:
            x = _layer(P, s + '/Conv0', x, taps=taps, pn=pn)
            if fused_scale:
                return _layer(P, s + '/Conv1_down', x, taps=taps, pn=pn, op=conv2d_downscale2d)
            x = _layer(P, s + '/Conv1', x, taps=taps, pn=pn)
            return downscale2d(x)
        x = _layer(P, s + '/Conv0', x, taps=taps, pn=pn)
        return _layer(P, s + '/z_Conv1', x, gain=1.0, act=False, taps=taps, pn=pn)

    out = _encoder_trunk(images_in, P, resolution, ll2, block, fromrgb, lod_in)
    if tanh_at_end:
        out = torch.tanh(out)
    return out[:, :latent_channels], out[:, latent_channels:]


def G_res(zg_latents_in, zl_latents_in, P, num_channels=3, resolution=128, fmap_base=1024,
          fmap_decay=1.0, fmap_max=512, latent_res=32, latent_channels=128, tanh_at_end=True,
          scale_h=1, scale_w=1, taps=None, use_pixelnorm=False, pixelnorm_epsilon=1e-8, fused_scale=False,
          tail_window=None, mid_window=None, **_):
    """networks.py:388-486 -> images [N,num_channels,resolution*scale_h,resolution*scale_w]."""
    pn = pixelnorm_epsilon if use_pixelnorm else None
    rl2 = int(np.log2(resolution))
    ll2 = int(np.log2(latent_res))
    assert resolution == 2 ** rl2 and latent_res == 2 ** ll2 and resolution >= latent_res
    want = (latent_channels, latent_res * scale_h, latent_res * scale_w)
    assert tuple(zg_latents_in.shape[1:]) == want and tuple(zl_latents_in.shape[1:]) == want
    lod_in = float(P['lod'])
    combo_in = torch.cat([zg_latents_in, zl_latents_in], dim=1)       # networks.py:423

    def block(x, res):
        s = '%dx%d' % (2 ** res, 2 ** res)
        if res == ll2:
            for count in range(5):                                     # networks.py:431-437
                x0 = x
                x = _layer(P, s + '/Residual%d_0' % count, x, taps=taps, pn=pn)
                x = _layer(P, s + '/Residual%d_1' % count, x, gain=1.0, act=False, taps=taps, pn=pn)
                x = x0 + x
                if taps is not None:
                    taps[s + '/Residual%d:sum' % count] = x
                if count == 3 and mid_window is not None:   # not in the reference: test hook (crop-aware windows)
                    oy, ox, th, tw = mid_window
                    x = x[:, :, oy:oy + th, ox:ox + tw]
            x = _layer(P, s + '/Conv0', x, gain=SQRT2 / 4, taps=taps, pn=pn)  # networks.py:440
            x = _layer(P, s + '/Conv1', x, taps=taps, pn=pn)
        else:
            if fused_scale:
                x = _layer(P, s + '/Conv0_up', x, taps=taps, pn=pn, op=upscale2d_conv2d)
            else:
                x = upscale2d(x)
                x = _layer(P, s + '/Conv0', x, taps=taps, pn=pn)
            x = _layer(P, s + '/Conv1', x, taps=taps, pn=pn)
        return x

    def torgb(x, res):
        return _layer(P, 'ToRGB_lod%d' % (rl2 - res), x, gain=1.0, act=False, taps=taps)

    def grow(x, res, lod):                                             # networks.py:473-479
        y = block(x, res)
        if res == ll2 and tail_window is not None:       # not in the reference: test hook for the crop-aware windows
            oy, ox, th, tw = tail_window
            y = y[:, :, oy:oy + th, ox:ox + tw]
        if lod > 0 and lod_in < lod:
            return grow(y, res + 1, lod - 1)
        if res > ll2 and lod_in > lod:
            return upscale2d(lerp(torgb(y, res), upscale2d(torgb(x, res - 1)), lod_in - lod), 2 ** lod)
        return upscale2d(torgb(y, res), 2 ** lod)

    images_out = grow(combo_in, ll2, rl2 - ll2)
    if taps is not None:
        taps['images_out:pre_tanh'] = images_out
    if tanh_at_end:
        images_out = torch.tanh(images_out)
    return images_out


def D_patch(images_in, P, num_channels=3, resolution=128, fmap_base=1024, fmap_decay=1.0, fmap_max=512,
            latent_res=-1, mbstd_group_size=4, taps=None, fused_scale=False, **_):
    """networks.py:491-577 -> scores [N,1,1,1] (latent_res=-1: FC head)."""
    rl2 = int(np.log2(resolution))
    ll2 = 2 if latent_res == -1 else int(np.log2(latent_res))
    assert tuple(images_in.shape[1:]) == (num_channels, resolution, resolution)
    lod_in = float(P['lod'])

    def fromrgb(x, res):
        return _layer(P, 'FromRGB_lod%d' % (rl2 - res), x, taps=taps)

    def block(x, res):
        s = '%dx%d' % (2 ** res, 2 ** res)
        if res > ll2:
            x = _layer(P, s + '/Conv0', x, taps=taps)
            if fused_scale:
                return _layer(P, s + '/Conv1_down', x, taps=taps, op=conv2d_downscale2d)
            x = _layer(P, s + '/Conv1', x, taps=taps)
            return downscale2d(x)
        if mbstd_group_size > 1:
            x = minibatch_stddev_layer(x, mbstd_group_size)
        x = _layer(P, s + '/Conv0', x, taps=taps)
        if latent_res == -1:
            x = leaky_relu(apply_bias(dense(x, P[s + '/Dense1/weight']), P[s + '/Dense1/bias']))
            x = apply_bias(dense(x, P[s + '/Dense2/weight'], gain=1.0), P[s + '/Dense2/bias'])
            return x[:, :, None, None]
        x = _layer(P, s + '/Conv1', x, taps=taps)
        return _layer(P, s + '/Conv2', x, gain=1.0, act=False, taps=taps)

    return _encoder_trunk(images_in, P, resolution, ll2, block, fromrgb, lod_in)


NETWORKS = dict(E_zg=E_zg, E_zl=E_zl, G_res=G_res, D_patch=D_patch)

# Hot-path config (config.py:39-59, 77-82): the kwargs the reference passes.
CONFIG = dict(
    E_zg=dict(fmap_base=1024, fmap_max=512, latent_channels=128, use_pixelnorm=False, tanh_at_end=False),
    E_zl=dict(fmap_base=1024, fmap_max=512, latent_res=32, latent_channels=128, use_pixelnorm=False,
              tanh_at_end=False),
    G_res=dict(fmap_base=1024, fmap_max=512, latent_res=32, latent_channels=128, use_pixelnorm=False,
               tanh_at_end=True),
    D_patch=dict(fmap_base=1024, fmap_max=512, latent_res=-1),
)

# Forward FLOPs per image: 2*k*k*Cin*Cout*H*W summed over convs.
GFLOP_PER_IMAGE = dict(E_zl=0.5636, E_zg=1.3042, G_res=12.9116, D_patch=1.2181)
