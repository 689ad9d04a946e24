"""VGG-19 Gram-matrix loss of the reference's E/G objective on the device:
`custom_vgg19.custom_Vgg19` (custom_vgg19.py:20-66) -> `gram_matrix` (loss.py:29-35) -> `multi_layer_diff`
(loss.py:68-75), used three times by `EG_wgan` (loss.py:148-160 reconstruction, :206-213 interpolated crop,
:248-257 blended crop) with `gram_weight = 0.002` (config.py:64).

The feature extractor is the build function `networks.Vgg19_features` behind an ordinary `Network` (tensor-core
convs with SAME zero padding, ReLU, 2x2 average pooling); its variables are constants loaded from the
tensorflow_vgg weight file `vgg19.npy` - which is not redistributable and not in the reference tree, so tests and
benchmarks use a seeded stand-in with the same layout (tests/loss_case.vgg_standin_weights) and the real file drops
in through `load_vgg19_npy`.  Gradients w.r.t. the images come from the same hand-written reverse pass as the other
networks (`backward.backward` with `param_grads=False`)."""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .backward import backward
from .network import Network
from .networks import VGG19_LAYERS


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def load_vgg19_npy(path):
    """custom_vgg19.loadWeightsData (custom_vgg19.py:10-17): the pickled dict {layer: [filter, bias]} of
    tensorflow_vgg/vgg19.npy (only load files you trust: it is a pickle)."""
    return np.load(path, encoding='latin1', allow_pickle=True).item()


VGG19_SHAPES = (('conv1_1', 3, 64), ('conv1_2', 64, 64), ('conv2_1', 64, 128), ('conv2_2', 128, 128),
                ('conv3_1', 128, 256), ('conv3_2', 256, 256), ('conv3_3', 256, 256), ('conv3_4', 256, 256),
                ('conv4_1', 256, 512), ('conv4_2', 512, 512), ('conv4_3', 512, 512), ('conv4_4', 512, 512),
                ('conv5_1', 512, 512))


def standin_weights(seed=19):
    """A seeded stand-in with vgg19.npy's layout ({layer: [filter [3,3,Cin,Cout], bias [Cout]]}, He-scaled random
    filters) for benchmarks and tests: the real file is not redistributable.  Same arithmetic and byte
    traffic as the real weights; the loss VALUES are of course not those of a trained VGG."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, cin, cout in VGG19_SHAPES:
        out[name] = [(rng.randn(3, 3, cin, cout) * np.sqrt(2.0 / (9 * cin))).astype(np.float32),
                     (0.05 * rng.randn(cout)).astype(np.float32)]
    return out


class GramLoss:
    """Owns the VGG-19 feature network and evaluates Gram terms + their image gradients.

    Tensor-core path (the default wherever the shapes allow): the feature extractor hands over its activations as
    split-bf16 planes (`gram_sink`), the Gram matrices are the weight-gradient kernel with the sample as the tap
    (tmx_gram_fwd_tc), and the gradient dF[n] = F[n] (S[n] + S[n]^T) / (h w) is a 1x1 convolution with one weight set
    per sample (TMX_CONV_W_PER_SAMPLE) whose NHWC output seeds the reverse pass directly.  Shapes those kernels do not
    serve (conv5_1: 64 pixels per sample) use the fp32 CUDA-core kernels on NCHW copies."""

    def __init__(self, data_dict, resolution=128, device=None):
        self.net = Network('Vgg19', func='networks.Vgg19_features', seed=0, num_channels=3, resolution=resolution,
                           device=device)
        for name, fmaps in VGG19_LAYERS:
            if name.startswith('pool'):
                continue
            filt, bias = data_dict[name][0], data_dict[name][1]
            self.net.set_var(name + '/weight', np.asarray(filt, np.float32))
            self.net.set_var(name + '/bias', np.asarray(bias, np.float32))
        self.rt = self.net.rt
        self.use_tc = not os.environ.get('TMX_GRAM_FFMA')

    # ------------------------------------------------------------------ Gram matrices
    def _gram_of(self, act):
        rt = self.rt
        n, c, h, w = act.n, act.c, act.h, act.w
        g = rt.empty(n, c, c)
        nbytes = C.c_size_t(0)
        if self.use_tc and act.hi is not None:
            _lib.check(rt.lib.tmx_gram_fwd_tc_workspace_bytes(rt.handle, n, c, h, w, C.byref(nbytes)),
                       'tmx_gram_fwd_tc_workspace_bytes')
        if nbytes.value:
            ws = rt.empty((nbytes.value + 3) // 4)
            _lib.check(rt.lib.tmx_gram_fwd_tc(rt.handle, _ptr(act.hi), _ptr(act.lo), _ptr(g), _ptr(ws), n, c, h, w,
                                              rt.stream()), 'tmx_gram_fwd_tc')
        else:
            f = rt.nhwc_to_nchw(rt.split_unpack(act).f32)
            _lib.check(rt.lib.tmx_gram_fwd(rt.handle, _ptr(f), _ptr(g), n, c, h, w, rt.stream()), 'tmx_gram_fwd')
        return g

    def features(self, images, tape=None):
        """images [N,3,R,R] in [-1,1] -> the five gram-layer activations in internal form (runtime.Act)."""
        acts = []
        self.net.get_output_for(images, return_as_list=True, tape=tape, gram_sink=acts)
        return acts

    def grams(self, images, tape=None):
        """images [N,3,R,R] in [-1,1] -> (activations of the five layers, their Gram matrices [N,C,C])."""
        acts = self.features(images, tape=tape)
        return acts, [self._gram_of(a) for a in acts]

    # ------------------------------------------------------------------ gradient w.r.t. one layer's features
    def _feature_gradient(self, act, S):
        """dL/dF (NHWC fp32) from S = dL/dG: dF[n][p][i] = sum_j (S[n][i][j] + S[n][j][i]) F[n][p][j] / h / w."""
        rt = self.rt
        n, c, h, w = act.n, act.c, act.h, act.w
        tiles = (h * w) // 128
        tc_ok = self.use_tc and act.hi is not None and c % 16 == 0 and c >= 64 and (h * w) % 128 == 0 and \
            tiles >= 1 and (c % 256 != 0 or tiles % 2 == 0)
        if tc_ok:
            w_hi = rt.empty(n * c, c, dtype=torch.bfloat16)
            w_lo = rt.empty(n * c, c, dtype=torch.bfloat16)
            _lib.check(rt.lib.tmx_gram_sym_split(rt.handle, _ptr(S), _ptr(w_hi), _ptr(w_lo), n, c, 1.0 / (h * w),
                                                 rt.stream()), 'tmx_gram_sym_split')
            out = rt.conv2d(act, None, None, 1.0, 1, c, lrelu=False, want_f32=True, want_split=False,
                            algo=_lib.ALGO_TC, prepared=(w_hi, w_lo), halo_in=act.halo, per_sample_weights=True)
            return out.f32
        f = rt.nhwc_to_nchw(rt.split_unpack(act).f32)
        df = rt.empty(n, c, h, w)
        _lib.check(rt.lib.tmx_gram_bwd(rt.handle, _ptr(S), _ptr(f), _ptr(df), n, c, h, w, rt.stream()), 'tmx_gram_bwd')
        return rt.nchw_to_nhwc(df)

    def term(self, images, targets, gram_weight):
        """One Gram term of the E/G loss for the fake batch `images` and its gradient w.r.t. them.
        targets: list of (target Gram matrices per layer, reverse_batch, wdev, wmode): the term is
            gram_weight * sum_targets w * mean_n sum_layers mean_{C x C} |G_l[n] - T_l[n']|
        with w = 1 / *wdev / 1 - *wdev (wmode 0 / 1 / 2) - one target for loss.py:157,210; two, weighted by the batch
        mean of the alpha draw, for loss.py:254 as written (its [N,1,1,1] x [N] product broadcasts to [N,1,1,N]).
        -> (batch mean of the term [device scalar], d term / d images [N,3,R,R])."""
        rt = self.rt
        n = images.shape[0]
        tape = []
        acts, gs = self.grams(images, tape=tape)
        sums = torch.zeros(n, dtype=torch.float32, device=rt.device)
        seeds = []
        for li, (a, g) in enumerate(zip(acts, gs)):
            c = g.shape[1]
            S = rt.empty(n, c, c)
            for ti, (tg, rev, wdev, wmode) in enumerate(targets):
                _lib.check(rt.lib.tmx_gram_l1(rt.handle, _ptr(g), _ptr(tg[li]), _ptr(S), _ptr(sums), n, c, int(rev),
                                              float(gram_weight) / (n * c * c), float(gram_weight) / (c * c),
                                              int(ti > 0), _ptr(wdev), int(wmode), rt.stream()), 'tmx_gram_l1')
            seeds.append((a, self._feature_gradient(a, S)))
        (dimg,) = backward(self.net, tape, [None] * len(acts), None, param_grads=False, seeds=seeds)
        value = rt.empty(1)
        _lib.check(rt.lib.tmx_row_sum(rt.handle, _ptr(sums), _ptr(value), 1, n, 1.0 / n, 0, 0, rt.stream()),
                   'tmx_row_sum')
        return value, dimg
