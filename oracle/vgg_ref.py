"""ORACLE (test infrastructure): torch restatement of the VGG-19 Gram-matrix loss of the reference's E/G objective -
`custom_vgg19.custom_Vgg19` (/root/reference/custom_vgg19.py:20-66 over machrisaa/tensorflow-vgg's conv_layer /
avg_pool, see oracle/tfshim/tensorflow_vgg/vgg19.py), `gram_matrix` (loss.py:29-35), `multi_layer_diff`
(loss.py:68-75) and the three Gram terms of `EG_wgan` (loss.py:148-160, 206-213, 248-257).

Pinned to the reference's own code: tests/golden/losses_gram.npz is minted by running loss.py + custom_vgg19.py
unmodified on oracle/tfshim with the seeded stand-in weights of tests/loss_case.vgg_standin_weights (the real
vgg19.npy is not redistributable); tests/test_loss_golden.py compares this restatement with it."""
import torch
import torch.nn.functional as F

VGG_MEAN = (103.939, 116.779, 123.68)                      # custom_vgg19.py:8 (BGR)
VGG_ORDER = ('conv1_1', 'conv1_2', 'pool1', 'conv2_1', 'conv2_2', 'pool2', 'conv3_1', 'conv3_2', 'conv3_3', 'conv3_4',
             'pool3', 'conv4_1', 'conv4_2', 'conv4_3', 'conv4_4', 'pool4', 'conv5_1')      # custom_vgg19.py:42-65
GRAM_LAYERS = ('conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1')                     # loss.py:153


def relu(x):
    """tensorflow_vgg's conv_layer activation; a module-level hook so that a test can feed the device's branch masks
    (tests/test_gpu_gram.py), like networks_ref.leaky_relu."""
    return torch.relu(x)


def vgg_features(images, data_dict, upto='conv5_1'):
    """images: [N,3,H,W] RGB in [-1,1] (NCHW).  -> {layer: NCHW activation} for the layers of GRAM_LAYERS.
    custom_vgg19.py:31-40: x = (rgb + 1) / 2 * 255, BGR order, minus VGG_MEAN; then conv (SAME zero pad) + bias +
    ReLU blocks with 2x2 average pooling in between."""
    x = (images + 1.0) / 2.0 * 255.0
    r, g, b = x[:, 0:1], x[:, 1:2], x[:, 2:3]
    x = torch.cat([b - VGG_MEAN[0], g - VGG_MEAN[1], r - VGG_MEAN[2]], dim=1)
    out = {}
    for name in VGG_ORDER:
        if name.startswith('pool'):
            x = F.avg_pool2d(x, 2, 2)
        else:
            w, bias = data_dict[name]
            w = torch.as_tensor(w, dtype=x.dtype)
            bias = torch.as_tensor(bias, dtype=x.dtype)
            x = relu(F.conv2d(x, w.permute(3, 2, 0, 1), padding=1) + bias.reshape(1, -1, 1, 1))
            if name in GRAM_LAYERS:
                out[name] = x
        if name == upto:
            break
    return out


def gram_matrix(x):
    """loss.py:29-35 on an NCHW activation: F^T F / h / w with F = [h*w, C] per sample -> [N, C, C]."""
    n, c, h, w = x.shape
    f = x.reshape(n, c, h * w)
    return torch.matmul(f, f.transpose(1, 2)) / float(h) / float(w)


def multi_layer_diff(feature, feature_):
    """loss.py:68-75: sum over layers of mean |f - f_| per sample -> [N]."""
    total = 0
    for f, f_ in zip(feature, feature_):
        total = total + (f - f_).abs().mean(dim=tuple(range(1, f.dim())))
    return total


def grams(images, data_dict):
    feats = vgg_features(images, data_dict)
    return [gram_matrix(feats[k]) for k in GRAM_LAYERS]


def blend_gram_term(blend_gram, real_gram, alpha, gram_weight):
    """loss.py:252-255 AS WRITTEN: alpha is [N,1,1,1] and multi_layer_diff returns [N], so the product broadcasts to
    [N,1,1,N] (sample i's alpha with sample j's difference) - the reference's own behaviour, kept (SURVEY Appendix D:
    do not fix quirks silently).  Its mean equals mean_i(1 - alpha_i) * mean_j A_j + mean_i(alpha_i) * mean_j B_j."""
    real2 = [torch.flip(m, dims=[0]) for m in real_gram]
    return ((1.0 - alpha) * multi_layer_diff(blend_gram, real2) + alpha * multi_layer_diff(blend_gram, real_gram)) \
        * gram_weight
