"""Seeded inputs of the loss goldens (tests/golden/losses.npz): ONE recipe shared by the generator
(tests/golden/make_golden.py gen_losses, which runs the reference's loss.py) and by the CPU / GPU tests that
compare the oracle restatement and the device path against the fixture."""
import os

import numpy as np

from oracle import interp_ref as I
from oracle import networks_ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'losses.npz')
LOSS_N = 4                 # per-tower minibatch: one full minibatch-stddev group, two reverse pairs for the blend
LOSS_GRAD_STRIDE = 97      # subsample of every variable gradient kept in the fixture (whole variable if <= 4096)
LOSS_NETS = ('E_zg', 'E_zl', 'G', 'D_rec', 'D_interp', 'D_blend')
LOSS_FUNCS = dict(E_zg='E_zg', E_zl='E_zl', G='G_res', D_rec='D_patch', D_interp='D_patch', D_blend='D_patch')
CROP_KEYS = ('eg_crop_interp', 'eg_crop_blend', 'd_interp_crop', 'd_blend_crop')
MIX_KEYS = ('eg_mix', 'd_rec_gp', 'd_interp_gp', 'd_blend_mix', 'd_blend_gp')


def loss_case_inputs(n=LOSS_N, scale_h=3, scale_w=3):
    rng = np.random.RandomState(1000)                                   # config.py:75
    params = {k: R.init_params(LOSS_FUNCS[k], rng, **R.CONFIG[LOSS_FUNCS[k]]) for k in LOSS_NETS}
    reals = rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)
    np.random.seed(1000)
    idx = I.sample_schedule_indices(n, latent_res=32, scale_h=scale_h, scale_w=scale_w)
    hy, hx = 128 * scale_h - 128, 128 * scale_w - 128
    crops = {k: (int(rng.randint(0, hy)), int(rng.randint(0, hx))) for k in CROP_KEYS}
    mixes = {k: rng.uniform(0, 1, (n, 1, 1, 1)).astype(np.float32) for k in MIX_KEYS}
    return params, reals, idx, crops, mixes


VGG_LAYERS = (('conv1_1', 3, 64), ('conv1_2', 64, 64), ('conv2_1', 64, 128), ('conv2_2', 128, 128),
              ('conv3_1', 128, 256), ('conv3_2', 256, 256), ('conv3_3', 256, 256), ('conv3_4', 256, 256),
              ('conv4_1', 256, 512), ('conv4_2', 512, 512), ('conv4_3', 512, 512), ('conv4_4', 512, 512),
              ('conv5_1', 512, 512), ('conv5_2', 512, 512), ('conv5_3', 512, 512), ('conv5_4', 512, 512))
GRAM_WEIGHT = 0.002        # config.py:64


def vgg_standin_weights(seed=19):
    """Stand-in for tensorflow_vgg/vgg19.npy (not redistributable): the same dict layout
    {layer: [filter [3,3,Cin,Cout], bias [Cout]]} with seeded He-scaled random filters, so that the Gram-loss path
    (custom_vgg19.py, loss.py:29-35,68-75,148-160,206-213,248-257) can be built and parity-tested; the real file
    drops in through the same loader."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, cin, cout in VGG_LAYERS:
        out[name] = [(rng.randn(3, 3, cin, cout) * np.sqrt(2.0 / (9 * cin))).astype(np.float32),
                     (0.05 * rng.randn(cout)).astype(np.float32)]
    return out


def gram_alpha(n=LOSS_N):
    """The extra uniform draw of the blend Gram term (loss.py:253)."""
    return np.random.RandomState(77).uniform(0, 1, (n, 1, 1, 1)).astype(np.float32)


def golden_gradient(g, tag, scope, name):
    """(subsample or whole gradient, its L2 norm over the WHOLE variable) as stored by gen_losses."""
    key = '%s_grad_%s_%s' % (tag, scope, name.replace('/', '.'))
    return g[key], float(g[key + '_norm'][0])


def subsample(flat):
    flat = np.asarray(flat).reshape(-1)
    return flat if flat.size <= 4096 else flat[::LOSS_GRAD_STRIDE]


# config-off interpolation modes (loss.py:176-193, 218-235): (tag, zg_interp_variational, zl_interp_variational)
MODE_CASES = (('var', 'variational', 'variational'), ('rnd', 'hard', 'random'), ('hard', 'hard', 'hard'))
MODES_N = 2                # (two samples: one reverse pair for the blend)


def mode_noise(n=MODES_N, c=128, lat=32, scale_h=3, scale_w=3, seed=4242):
    """The tf.random_normal draws of the sampling modes as canvas-shaped tensors (see oracle.loss_ref.zl_canvas)."""
    rng = np.random.RandomState(seed)
    H, W = lat * scale_h, lat * scale_w
    return {'zg_f': rng.standard_normal((n, c, 1, 1)).astype(np.float32),
            'zl_f': rng.standard_normal((n, c, H, W)).astype(np.float32),
            'zg_b': rng.standard_normal((n, c, 1, 1)).astype(np.float32),
            'zl_b': rng.standard_normal((n, c, H, W)).astype(np.float32)}
