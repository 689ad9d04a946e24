"""Stub for the reference's `custom_vgg19` (needs the un-vendored
`tensorflow_vgg` package + vgg19.npy, README.md:39).  The VGG Gram loss is out
of scope; this stub only lets `networks.py:13` import."""


def loadWeightsData(*a, **k):
    raise NotImplementedError('VGG19 weights are not available (gram_weight must be 0)')


class custom_Vgg19:
    def __init__(self, *a, **k):
        raise NotImplementedError('VGG19 weights are not available (gram_weight must be 0)')
