"""Generate the committed golden fixtures by running the REFERENCE's own code
from /root/reference in this container (it cannot travel to the GPU box).

    python tests/golden/make_golden.py

* perm_sampler.npz  - run.py:107-182 `my_swap_h/w` + `block_permutation` driven
  exactly like run.py:436-507 / util_scripts.py:402-442 under np.random.seed(..),
  reduced to argmax index vectors (the integer tile-index grid, bit-exact).
* mattes.npz        - util_scripts.py:85-102 `linkern_for_weight_*`.
* networks.npz      - networks.py `E_zg/E_zl/G_res/D_patch` executed unmodified
  on oracle/tfshim (torch-backed TF stand-in) with the seeded parameters of
  oracle.networks_ref.init_params; outputs stored (strided subsample for the
  large ones) with the variable creation order.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refload  # noqa: E402
from oracle import networks_ref as R  # noqa: E402

SUBSAMPLE = 37  # stride for large outputs


def gen_perm():
    f = refload.reference_functions('run.py', ['my_swap_h', 'my_swap_w', 'block_permutation'])
    out = {}
    # (tag, length, levels, count, seed): training config (run.py:440: int(log2(32)) = 5 levels, 96 long),
    # inference quirk (util_scripts.py:405: int(np.log(32)) = 3 levels, 128 long), and small/edge sizes.
    cases = [('train96', 96, 5, 3, 1000), ('interp128', 128, 3, 2, 1000), ('small8', 8, 3, 4, 7),
             ('tiny2', 2, 1, 4, 3), ('wide256', 256, 5, 1, 11)]
    for tag, length, levels, count, seed in cases:
        np.random.seed(seed)
        draws = 0
        hs, ws = [], []
        for _ in range(count):                       # h matrices first, like run.py:437-452
            p = np.eye(length)
            for idx in range(levels):
                bs = int(2 ** idx)
                perm = f['my_swap_h'](np.eye(length // bs))
                p = np.matmul(p, f['block_permutation'](perm, bs))
            assert (p.sum(0) == 1).all() and (p.sum(1) == 1).all()
            hs.append(np.argmax(p, axis=1).astype(np.int32))
        for _ in range(count):                       # then w matrices, run.py:453-469
            p = np.eye(length)
            for idx in range(levels):
                bs = int(2 ** idx)
                perm = f['my_swap_w'](np.eye(length // bs))
                p = np.matmul(f['block_permutation'](perm, bs), p)
            ws.append(np.argmax(p, axis=0).astype(np.int32))
        out[tag + '_h'] = np.stack(hs)
        out[tag + '_w'] = np.stack(ws)
        out[tag + '_meta'] = np.array([length, levels, count, seed], np.int64)
        out[tag + '_next_uniform'] = np.array([np.random.uniform()])   # pins the number of draws consumed
    np.savez_compressed(os.path.join(HERE, 'perm_sampler.npz'), **out)
    print('perm_sampler.npz', {k: v.shape for k, v in out.items() if not k.endswith('meta')})


def gen_perm_options():
    """The config-off branches of run.py:436-507: config.block_size > 0 and config.perm (np.random.permutation),
    driven like the reference's loops (h matrices first, then w), reduced to index vectors."""
    f = refload.reference_functions('run.py', ['my_swap_h', 'my_swap_w', 'block_permutation'])
    out = {}
    for tag, length, levels, count, seed, block_size, perm in [('bs4', 96, 5, 2, 5, 4, False), ('perm', 96, 5, 2, 6, 0, True),
                                                               ('bs8perm', 96, 5, 2, 7, 8, True), ('bs32', 64, 5, 3, 8, 32, False)]:
        np.random.seed(seed)
        hs, ws = [], []
        for axis, acc in (('h', hs), ('w', ws)):
            swap = f['my_swap_h'] if axis == 'h' else f['my_swap_w']
            for _ in range(count):
                if block_size == 0:
                    p = np.eye(length)
                    for idx in range(levels):
                        bs = int(2 ** idx)
                        pm = np.random.permutation(np.eye(length // bs)) if perm else swap(np.eye(length // bs))
                        blk = f['block_permutation'](pm, bs)
                        p = np.matmul(p, blk) if axis == 'h' else np.matmul(blk, p)
                else:
                    pm = np.random.permutation(np.eye(length // block_size)) if perm else swap(np.eye(length // block_size))
                    p = f['block_permutation'](pm, block_size)
                assert (p.sum(0) == 1).all() and (p.sum(1) == 1).all()
                acc.append(np.argmax(p, axis=1 if axis == 'h' else 0).astype(np.int32))
        out[tag + '_h'] = np.stack(hs)
        out[tag + '_w'] = np.stack(ws)
        out[tag + '_meta'] = np.array([length, levels, count, seed, block_size, int(perm)], np.int64)
        out[tag + '_next_uniform'] = np.array([np.random.uniform()])
    np.savez_compressed(os.path.join(HERE, 'perm_sampler_options.npz'), **out)
    print('perm_sampler_options.npz', {k: v.shape for k, v in out.items() if not k.endswith('meta')})


def gen_mattes():
    f = refload.reference_functions('util_scripts.py',
                                    ['linkern_for_weight_horizontal', 'linkern_for_weight_arbitrary_shape'])
    out = {}
    for (h, w, r) in [(128, 128, 32), (96, 256, 32), (72, 80, 8)]:
        ul, ur, bl, br = f['linkern_for_weight_arbitrary_shape'](h, w, r)
        out['arb_%d_%d_%d' % (h, w, r)] = np.stack([ul, ur, bl, br])           # float64
    out['hor_1_2_4_256_32'] = f['linkern_for_weight_horizontal']([1, 2, 4, 256], 32)
    np.savez_compressed(os.path.join(HERE, 'mattes.npz'), **out)
    print('mattes.npz', {k: (v.shape, v.dtype) for k, v in out.items()})


APP_MATTE_CASES = dict(
    gkern_for_weight_arbitrary_shape=[(24, 40, 7, 5, 6.0), (32, 32, 31.5, 0, 12.0), (20, 48, -3, 25, 3.0)],
    gkern_for_weight_arbitrary_shape_hybridization=[(24, 40, 7, 5, 6.0), (16, 16, 8, 8, 2.5)],
    gkern_for_weight_grid_shape_hybridization=[(24, 40, 3.25, 4.5, 8.0, 6.0), (32, 24, -2.0, 20.0, 8.0, 4.0),
                                               (16, 16, 0.0, 0.0, 32.0, 8.0)],
    linkern_for_weight_square=[(24, 4), (40, 8)],
    gkern_for_scale_horizontal=[([2, 3, 4, 40], 8)],
)


def gen_app_mattes():
    """util_scripts.py:53-62, 104-182: the brush / hybridization weight kernels, executed from the reference file."""
    f = refload.reference_functions('util_scripts.py', list(APP_MATTE_CASES) + ['l2', 'dist2square'])
    out = {}
    for name, cases in APP_MATTE_CASES.items():
        for i, args in enumerate(cases):
            r = f[name](*args)
            out['%s_%d' % (name, i)] = np.stack(r) if isinstance(r, tuple) else np.asarray(r)
    np.savez_compressed(os.path.join(HERE, 'mattes_apps.npz'), **out)
    print('mattes_apps.npz', {k: (v.shape, v.dtype) for k, v in out.items()})


def network_inputs(func, rng, n):
    if func == 'G_res':
        return [rng.randn(n, 128, 32, 32).astype(np.float32), rng.randn(n, 128, 32, 32).astype(np.float32)]
    return [rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)]


def gen_networks():
    net, tf = refload.reference_networks()
    out = {}
    for func, n in [('E_zg', 2), ('E_zl', 2), ('G_res', 2), ('D_patch', 8)]:
        rng = np.random.RandomState(1000)                       # config.py:75 random_seed
        cfg = R.CONFIG[func]
        params = R.init_params(func, rng, **cfg)
        ins = network_inputs(func, rng, n)
        tf.reset_default_graph(values={func + '/' + k: v for k, v in params.items()})
        with tf.variable_scope(func):
            res = getattr(net, func)(*[tf.convert_to_tensor(a) for a in ins], num_channels=3, resolution=128, **cfg)
        res = res if isinstance(res, tuple) else (res,)
        names = [k[len(func) + 1:] for k in tf.STORE.vars]
        assert names == list(params.keys())
        out[func + '_varnames'] = np.array(names)
        for i, r in enumerate(res):
            a = r.numpy()
            out['%s_out%d_shape' % (func, i)] = np.array(a.shape, np.int64)
            flat = a.reshape(-1)
            out['%s_out%d' % (func, i)] = flat[::SUBSAMPLE] if flat.size > 4096 else flat
            out['%s_out%d_absmax' % (func, i)] = np.array([np.abs(a).max()], np.float32)
    np.savez_compressed(os.path.join(HERE, 'networks.npz'), **out)
    print('networks.npz', {k: v.shape for k, v in out.items() if 'out' in k and 'shape' not in k})


# (tag, func, lod, use_pixelnorm): progressive-growing levels of detail (integer and fractional: every tf.cond
# branch of networks.py:276-282, 368-374, 473-479, 568-574) and the pixel-norm variant (networks.py:170-172)
VARIANTS = [('E_zg', 1.0, False), ('E_zg', 0.5, False), ('E_zg', 3.25, False), ('E_zg', 5.0, False),
            ('E_zl', 1.0, False), ('E_zl', 1.5, False), ('E_zl', 2.0, False),
            ('G_res', 1.0, False), ('G_res', 0.25, False), ('G_res', 1.75, False), ('G_res', 2.0, False),
            ('D_patch', 1.0, False), ('D_patch', 2.5, False), ('D_patch', 5.0, False),
            ('E_zg', 0.0, True), ('E_zl', 0.0, True), ('G_res', 0.0, True), ('G_res', 1.5, True)]


def variant_tag(func, lod, pn):
    return '%s_lod%s%s' % (func, ('%g' % lod).replace('.', 'p'), '_pn' if pn else '')


# fused_scale=True (networks.py:94-101 upscale2d_conv2d, :142-148 conv2d_downscale2d): other variables, other ops
FUSED_VARIANTS = [('E_zg', 0.0), ('E_zl', 0.0), ('G_res', 0.0), ('D_patch', 0.0), ('G_res', 1.5), ('E_zl', 0.5)]


def gen_network_fused():
    net, tf = refload.reference_networks()
    out = {}
    for func, lod in FUSED_VARIANTS:
        n = 4 if func == 'D_patch' else 2
        rng = np.random.RandomState(1000)
        cfg = dict(R.CONFIG[func], fused_scale=True)
        params = R.init_params(func, rng, **cfg)
        params['lod'] = np.float32(lod)
        ins = network_inputs(func, rng, n)
        tf.reset_default_graph(values={func + '/' + k: v for k, v in params.items()})
        with tf.variable_scope(func):
            res = getattr(net, func)(*[tf.convert_to_tensor(a) for a in ins], num_channels=3, resolution=128, **cfg)
        res = res if isinstance(res, tuple) else (res,)
        names = [k[len(func) + 1:] for k in tf.STORE.vars]
        assert names == list(params.keys()), (names, list(params.keys()))
        tag = variant_tag(func, lod, False) + '_fused'
        out[tag + '_varnames'] = np.array(names)
        for i, r in enumerate(res):
            a = r.numpy()
            flat = a.reshape(-1)
            out['%s_out%d_shape' % (tag, i)] = np.array(a.shape, np.int64)
            out['%s_out%d' % (tag, i)] = (flat[::SUBSAMPLE] if flat.size > 4096 else flat).astype(np.float32)
            out['%s_out%d_absmax' % (tag, i)] = np.array([np.abs(a).max()], np.float32)
    np.savez_compressed(os.path.join(HERE, 'networks_fused.npz'), **out)
    print('networks_fused.npz', sorted(k for k in out if k.endswith('_out0')))


def gen_network_variants():
    net, tf = refload.reference_networks()
    out = {}
    for func, lod, pn in VARIANTS:
        n = 4 if func == 'D_patch' else 2
        rng = np.random.RandomState(1000)
        cfg = dict(R.CONFIG[func])
        if pn:
            cfg['use_pixelnorm'] = True
        params = R.init_params(func, rng, **cfg)
        params['lod'] = np.float32(lod)
        ins = network_inputs(func, rng, n)
        tf.reset_default_graph(values={func + '/' + k: v for k, v in params.items()})
        with tf.variable_scope(func):
            res = getattr(net, func)(*[tf.convert_to_tensor(a) for a in ins], num_channels=3, resolution=128, **cfg)
        res = res if isinstance(res, tuple) else (res,)
        tag = variant_tag(func, lod, pn)
        for i, r in enumerate(res):
            a = r.numpy()
            flat = a.reshape(-1)
            out['%s_out%d_shape' % (tag, i)] = np.array(a.shape, np.int64)
            out['%s_out%d' % (tag, i)] = (flat[::SUBSAMPLE] if flat.size > 4096 else flat).astype(np.float32)
            out['%s_out%d_absmax' % (tag, i)] = np.array([np.abs(a).max()], np.float32)
    np.savez_compressed(os.path.join(HERE, 'networks_variants.npz'), **out)
    print('networks_variants.npz', sorted(k for k in out if k.endswith('_out0')))


# (num_gpus, schedule kwargs): the reference's own config (config.py:96 + run.py:649-658 presets) and the defaults
SCHEDULES = [
    (1, dict(lod_initial_resolution=32, lod_training_kimg=1000, lod_transition_kimg=3000, minibatch_base=4,
             lrate_dict={128: 0.0015, 256: 0.002, 512: 0.003, 1024: 0.003})),
    (8, dict(lod_initial_resolution=32, lod_training_kimg=1000, lod_transition_kimg=3000, minibatch_base=32,
             lrate_dict={128: 0.0015, 256: 0.002, 512: 0.003, 1024: 0.003}, max_minibatch_per_gpu={128: 3})),
    (2, dict()),
]
SCHEDULE_NIMG = [0, 1, 999999, 1000000, 1000001, 2500000, 3999999, 4000000, 4700000, 7999000, 8000000, 8000001,
                 12000000, 500000000]


def gen_schedule():
    """run.py:187-226 TrainingSchedule executed from the reference file -> (lod, resolution, minibatch, lrate, tick)."""
    import types
    out = {}
    for i, (gpus, kw) in enumerate(SCHEDULES):
        cfg = types.SimpleNamespace(num_gpus=gpus)
        cls = refload.reference_functions('run.py', ['TrainingSchedule'], extra_globals={'config': cfg})['TrainingSchedule']
        rows = []
        for nimg in SCHEDULE_NIMG:
            s = cls(nimg, types.SimpleNamespace(resolution_log2=7), **kw)
            rows.append([s.lod, s.resolution, s.minibatch, s.lrate, s.tick_kimg])
        out['sched%d' % i] = np.array(rows, np.float64)
    np.savez_compressed(os.path.join(HERE, 'schedule.npz'), **out)
    print('schedule.npz', {k: v.shape for k, v in out.items()})


# ---------------------------------------------------------------- losses (loss.py executed unmodified)
sys.path.insert(0, os.path.dirname(HERE))
from loss_case import (LOSS_N, LOSS_GRAD_STRIDE, LOSS_NETS, GRAM_WEIGHT, MODE_CASES, MODES_N, mode_noise,  # noqa: E402
                       loss_case_inputs, vgg_standin_weights,
                       gram_alpha)                                                                      # (tests/loss_case.py)


def gen_losses():
    """loss.py:105-259 (EG_wgan) and :303-521 (D_rec/D_interp/D_blend_wgangp) executed UNMODIFIED on oracle/tfshim
    with the reference's own config.py values (config.py:50-68, 89-92) except gram_weight = 0 (VGG weights are not
    in the tree).  The networks are the reference's networks.py behind `refload.ReferenceNetwork`; the graph's random
    ops (random_crop offsets, mixing factors) are pinned through the shim's RANDOM hooks and stored.
    Stored: per-sample loss vectors, every autosummary'd term, and d mean(loss) / d variable (tfutil.py:299 over the
    optimizer's own variables, run.py:321-324) - strided subsample + L2 norm per variable."""
    from oracle import interp_ref as I
    lossmod, net, tf = refload.reference_loss()
    import config as refcfg                                             # the reference's config.py (pure Python)
    import tfutil as shim_tfutil
    assert refcfg.__file__.startswith(refload.REFERENCE_ROOT) and lossmod.__file__.startswith(refload.REFERENCE_ROOT)
    n, sh, sw = LOSS_N, refcfg.scale_h, refcfg.scale_w
    params, reals, idx, crops, mixes = loss_case_inputs(n, sh, sw)
    values = {}
    scope_of = dict(E_zg='E_zg', E_zl='E_zl', G='G', D_rec='D_rec', D_interp='D_interp', D_blend='D_blend')
    for k in LOSS_NETS:
        values.update({scope_of[k] + '/' + vn: v for vn, v in params[k].items()})
    lat, res = refcfg.latent_res_EG, refcfg.train_size
    C = refcfg.latent_channels

    def kw(d):
        d = dict(d)
        d.pop('func')
        return d
    nets = dict(
        E_zg=refload.ReferenceNetwork(net, tf, 'E_zg', 'E_zg', [[None, 3, res, res]], [[None, C, 1, 1]] * 2,
                                      num_channels=3, resolution=res, **kw(refcfg.E_zg)),
        E_zl=refload.ReferenceNetwork(net, tf, 'E_zl', 'E_zl', [[None, 3, res, res]], [[None, C, lat, lat]] * 2,
                                      num_channels=3, resolution=res, **kw(refcfg.E_zl)),
        G=refload.ReferenceNetwork(net, tf, 'G', 'G_res', [[None, C, lat, lat]] * 2, [[None, 3, res, res]],
                                   num_channels=3, resolution=res, **kw(refcfg.G)),
        G_fcn=refload.ReferenceNetwork(net, tf, 'G', 'G_res', [[None, C, lat * sh, lat * sw]] * 2,
                                       [[None, 3, res * sh, res * sw]], num_channels=3, resolution=res, scale_h=sh,
                                       scale_w=sw, **kw(refcfg.G)),                     # run.py:273
    )
    for k in ('D_rec', 'D_interp', 'D_blend'):
        nets[k] = refload.ReferenceNetwork(net, tf, k, 'D_patch', [[None, 3, res, res]], [[None, 1, 1, 1]],
                                           num_channels=3, resolution=res, **kw(getattr(refcfg, k)))
    ph = np.stack([I.index_to_matrix_h(r) for r in idx['h_forward']])[:, None].astype(np.float32)
    pw = np.stack([I.index_to_matrix_w(c) for c in idx['w_forward']])[:, None].astype(np.float32)
    phb = np.stack([I.index_to_matrix_h(r) for r in idx['h_backward']])[:, None].astype(np.float32)
    pwb = np.stack([I.index_to_matrix_w(c) for c in idx['w_backward']])[:, None].astype(np.float32)
    opt = refload.ReferenceOptimizerStub()
    eg_kw = kw(refcfg.EG_loss)
    eg_kw['gram_weight'] = 0.0
    out = {'meta_n_sh_sw_stride': np.array([n, sh, sw, LOSS_GRAD_STRIDE], np.int64)}
    for k, v in crops.items():
        out['draw_' + k] = np.array(v, np.int64)
    for k, v in mixes.items():
        out['draw_' + k] = v

    out_all = out

    def run(tag, fn, trainable_scopes, int_draws, float_draws, out=None, x_np=None, normal_draws=()):
        out = out_all if out is None else out
        tf.reset_default_graph(values=values, requires_grad=True)
        terms = {}

        def record(name, value):
            terms[name] = value
            return value
        shim_tfutil.autosummary = record
        ints, floats = list(int_draws), list(float_draws)
        tf.RANDOM['uniform_int'] = lambda shp, lo, hi: np.full(shp, ints.pop(0), np.int64)
        tf.RANDOM['uniform'] = lambda shp, lo, hi: floats.pop(0).reshape(shp)
        normals = list(normal_draws)

        def next_normal(shp):
            a = normals.pop(0)
            assert tuple(a.shape) == tuple(int(v) for v in shp), (tag, 'tf.random_normal shape', a.shape, shp)
            return a
        tf.RANDOM['normal'] = next_normal
        x = tf.convert_to_tensor(reals if x_np is None else x_np)
        loss = fn(x)
        assert not ints and not floats and not normals, (tag, 'unused random draws', ints, floats, len(normals))
        tf.RANDOM['uniform_int'] = tf.RANDOM['uniform'] = tf.RANDOM['normal'] = None
        loss.t.mean().backward()                                        # tf.reduce_mean(loss), run.py:321-324
        out[tag + '_loss'] = loss.numpy().astype(np.float32)
        for name, val in terms.items():
            out[tag + '_term_' + name.replace('/', '_')] = val.numpy().astype(np.float32)
        for full, var in tf.STORE.vars.items():
            scope, vn = full.split('/', 1)
            if scope not in trainable_scopes or not var.trainable:
                continue
            g = var.t.grad
            g = np.zeros(tuple(var.t.shape), np.float32) if g is None else g.numpy()   # tfutil.py:298
            flat = g.reshape(-1)
            key = '%s_grad_%s_%s' % (tag, scope, vn.replace('/', '.'))
            out[key] = (flat if flat.size <= 4096 else flat[::LOSS_GRAD_STRIDE]).astype(np.float32)
            out[key + '_norm'] = np.array([np.linalg.norm(flat.astype(np.float64))])
        print(tag, 'loss', out[tag + '_loss'], {k: float(np.mean(v.numpy())) for k, v in terms.items()})

    ci, cb = crops['eg_crop_interp'], crops['eg_crop_blend']
    run('EG', lambda x: lossmod.EG_wgan(
        nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], nets['G_fcn'], nets['D_interp'], nets['D_blend'], n, x, x,
        None, tf.convert_to_tensor(ph), tf.convert_to_tensor(pw), tf.convert_to_tensor(phb), tf.convert_to_tensor(pwb),
        **eg_kw), ('E_zg', 'E_zl', 'G'), [ci[0], ci[1], cb[0], cb[1]], [mixes['eg_mix']])
    run('D_rec', lambda x: lossmod.D_rec_wgangp(nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], opt, n, x, x,
                                                **kw(refcfg.D_rec_loss)), ('D_rec',), [], [mixes['d_rec_gp']])
    c = crops['d_interp_crop']
    run('D_interp', lambda x: lossmod.D_interp_wgangp(
        nets['E_zg'], nets['E_zl'], nets['G_fcn'], nets['D_interp'], opt, n, x, x, tf.convert_to_tensor(ph),
        tf.convert_to_tensor(pw), **kw(refcfg.D_interp_loss)), ('D_interp',), [c[0], c[1]], [mixes['d_interp_gp']])
    c = crops['d_blend_crop']
    run('D_blend', lambda x: lossmod.D_blend_wgangp(
        nets['E_zg'], nets['E_zl'], nets['G_fcn'], nets['D_blend'], opt, n, x, x, tf.convert_to_tensor(ph),
        tf.convert_to_tensor(pw), tf.convert_to_tensor(phb), tf.convert_to_tensor(pwb), **kw(refcfg.D_blend_loss)),
        ('D_blend',), [c[0], c[1]], [mixes['d_blend_mix'], mixes['d_blend_gp']])
    np.savez_compressed(os.path.join(HERE, 'losses.npz'), **out)
    print('losses.npz: %d arrays, %.1f MB' % (len(out), os.path.getsize(os.path.join(HERE, 'losses.npz')) / 1e6))

    # ---- the same E/G loss WITH the VGG-19 Gram terms (config.py:64 gram_weight = 0.002): custom_vgg19.py and the
    # Gram branches of loss.py run unmodified; `tensorflow_vgg.vgg19.Vgg19` is the shim's restatement of the
    # un-vendored base class and the weights are the seeded stand-in (vgg19.npy is not redistributable)
    data_dict = vgg_standin_weights()
    assert refcfg.gram_weight == GRAM_WEIGHT
    lossmod.loadWeightsData = lambda path=None: data_dict           # np.load of tensorflow_vgg/vgg19.npy
    gram = {'meta_n_sh_sw_stride': out['meta_n_sh_sw_stride'], 'draw_eg_gram_alpha': gram_alpha(n)}
    eg_kw_gram = kw(refcfg.EG_loss)
    run('EGgram', lambda x: lossmod.EG_wgan(
        nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], nets['G_fcn'], nets['D_interp'], nets['D_blend'], n, x, x,
        None, tf.convert_to_tensor(ph), tf.convert_to_tensor(pw), tf.convert_to_tensor(phb), tf.convert_to_tensor(pwb),
        **eg_kw_gram), ('E_zg', 'E_zl', 'G'), [ci[0], ci[1], cb[0], cb[1]], [mixes['eg_mix'], gram_alpha(n)], out=gram)
    np.savez_compressed(os.path.join(HERE, 'losses_gram.npz'), **gram)
    print('losses_gram.npz: %d arrays, %.1f MB' % (len(gram), os.path.getsize(os.path.join(HERE, 'losses_gram.npz')) / 1e6))

    # ---- EG_wgan with the config-off interpolation modes (zg_/zl_interp_variational, loss.py:176-193, 218-235); the
    # graph's tf.random_normal draws are fed in the reference's own call order from canvas-shaped noise tensors
    n2 = MODES_N
    _, reals2, idx2, crops2, mixes2 = loss_case_inputs(n2, sh, sw)
    mats2 = [tf.convert_to_tensor(np.stack([f(r) for r in idx2[k]])[:, None].astype(np.float32))
             for k, f in (('h_forward', I.index_to_matrix_h), ('w_forward', I.index_to_matrix_w),
                          ('h_backward', I.index_to_matrix_h), ('w_backward', I.index_to_matrix_w))]
    noise = mode_noise(n2, C, lat, sh, sw)
    H2, W2 = lat * sh, lat * sw
    modes = {'meta_n_sh_sw_stride': np.array([n2, sh, sw, LOSS_GRAD_STRIDE], np.int64)}

    def normals_for(zg, zl, which):
        out_ = []
        if zg == 'variational':
            out_.append(noise['zg_' + which])
        e = noise['zl_' + which]
        if zl == 'variational':
            out_.append(e)
        elif zl == 'random':                                             # loss.py:189-191: three separately drawn blocks
            out_ += [e[:, :, :lat, lat:W2 - lat], e[:, :, lat:H2 - lat, :], e[:, :, H2 - lat:, lat:W2 - lat]]
        return out_
    ci2, cb2 = crops2['eg_crop_interp'], crops2['eg_crop_blend']
    for tag, zg, zl in MODE_CASES:
        kw_m = kw(refcfg.EG_loss)
        kw_m['gram_weight'] = 0.0
        kw_m['zg_interp_variational'], kw_m['zl_interp_variational'] = zg, zl
        run('EG' + tag, lambda x: lossmod.EG_wgan(
            nets['E_zg'], nets['E_zl'], nets['G'], nets['D_rec'], nets['G_fcn'], nets['D_interp'], nets['D_blend'], n2,
            x, x, None, *mats2, **kw_m), ('E_zg', 'E_zl', 'G'), [ci2[0], ci2[1], cb2[0], cb2[1]], [mixes2['eg_mix']],
            out=modes, x_np=reals2, normal_draws=normals_for(zg, zl, 'f') + normals_for(zg, zl, 'b'))
    np.savez_compressed(os.path.join(HERE, 'losses_modes.npz'), **modes)
    print('losses_modes.npz: %d arrays, %.1f MB' % (len(modes), os.path.getsize(os.path.join(HERE, 'losses_modes.npz')) / 1e6))


if __name__ == '__main__':
    assert refload.reference_available(), 'needs /root/reference'
    if 'losses' in sys.argv[1:]:
        gen_losses()
        sys.exit(0)
    if 'variants' in sys.argv[1:]:
        gen_network_variants()
        sys.exit(0)
    if 'schedule' in sys.argv[1:]:
        gen_schedule()
        sys.exit(0)
    if 'apps' in sys.argv[1:]:
        gen_app_mattes()
        sys.exit(0)
    if 'perm_options' in sys.argv[1:]:
        gen_perm_options()
        sys.exit(0)
    if 'fused' in sys.argv[1:]:
        gen_network_fused()
        sys.exit(0)
    gen_perm()
    gen_perm_options()
    gen_mattes()
    gen_networks()
    gen_network_variants()
    gen_schedule()
    gen_app_mattes()
    gen_network_fused()
    gen_losses()
