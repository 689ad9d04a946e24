"""Host logic of the crop-aware G_fcn evaluation (texturemixer_b200.loss.crop_window, SURVEY Appendix C note): decoding
only the latent window a random_crop depends on must give the SAME crop pixels and the SAME gradients as decoding the
whole canvas.  Checked on the CPU with the oracle's G_res (thin channels: the cone of dependence is a property of the
layer sequence networks.py:427-457, not of the widths)."""
import numpy as np
import pytest
import torch

from oracle import networks_ref as R
from texturemixer_b200 import loss as dev_loss
from texturemixer_b200.loss import G_CONTEXT, compose_window, crop_window, image_offset, mid_window, tail_window

CFG = dict(num_channels=3, resolution=128, fmap_base=64, fmap_max=8, latent_res=32, latent_channels=2,
           use_pixelnorm=False, tanh_at_end=True)
RES, LAT, S = 128, 32, 3


def test_window_geometry():
    H = W = LAT * S
    for y in range(0, RES * S - RES):
        oy, ox, wh, ww = crop_window((y, 5), RES, LAT, H, W)
        assert wh == ww == 64 and 0 <= oy <= H - wh and ox == 0
        lo, hi = y // 4 - G_CONTEXT, (y + RES - 1) // 4 + G_CONTEXT          # latent rows the crop's cone touches
        assert oy <= max(lo, 0) and oy + wh > min(hi, H - 1)
        # an edge of the window that is not an edge of the canvas keeps G_CONTEXT latent pixels of distance
        assert oy == 0 or y // 4 - oy >= G_CONTEXT
        assert oy + wh == H or oy + wh - 1 - (y + RES - 1) // 4 >= G_CONTEXT
    assert crop_window((3, 9), RES, LAT, 64, 64) is None                      # canvas no larger than the window
    assert crop_window((3, 9), RES, LAT, 64, 96) == (0, 0, 64, 64) or crop_window((3, 9), RES, LAT, 64, 96)[2] == 64
    assert crop_window((0, 0), 64, 32, 96, 96) is None                        # other up-sampling depths: whole canvas


@pytest.mark.parametrize('yx', [(0, 0), (255, 255), (57, 131), (128, 3), (56, 200), (199, 60), (1, 254), (130, 129)])
def test_windowed_decode_equals_full_decode(yx):
    rng = np.random.RandomState(7)
    P = R.to_torch(R.init_params('G_res', rng, **CFG), requires_grad=True)
    H = W = LAT * S
    zg = torch.from_numpy(rng.randn(1, 2, 1, 1).astype(np.float32)).repeat(1, 1, H, W)
    zl = torch.from_numpy(rng.randn(1, 2, H, W).astype(np.float32)).requires_grad_(True)
    seed = torch.from_numpy(rng.randn(1, 3, RES, RES).astype(np.float32))
    cfg = dict(CFG, scale_h=S, scale_w=S)

    full = R.G_res(zg, zl, P, **cfg)[:, :, yx[0]:yx[0] + RES, yx[1]:yx[1] + RES]
    names = [k for k in P if k != 'lod' and P[k].requires_grad]
    g_full = torch.autograd.grad((full * seed).sum(), [zl] + [P[k] for k in names], allow_unused=True)

    oy, ox, wh, ww = crop_window(yx, RES, LAT, H, W)
    zl_w = zl[:, :, oy:oy + wh, ox:ox + ww]
    img = R.G_res(zg[:, :, :wh, :ww], zl_w, P, **dict(CFG, scale_h=wh // LAT, scale_w=ww // LAT))
    y0, x0 = yx[0] - 4 * oy, yx[1] - 4 * ox
    part = img[:, :, y0:y0 + RES, x0:x0 + RES]
    g_part = torch.autograd.grad((part * seed).sum(), [zl] + [P[k] for k in names], allow_unused=True)

    assert part.shape == full.shape
    assert float((part - full).detach().abs().max()) <= 1e-5      # fp32 noise of the CPU conv blocking; a cone error is O(0.1)
    for name, a, b in zip(['zl'] + names, g_full, g_part):
        if a is None and b is None:
            continue
        scale = float(a.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= 1e-4 * scale, name


@pytest.mark.parametrize('yx', [(0, 0), (255, 255), (57, 131), (128, 3), (56, 200), (199, 60), (1, 254), (130, 129),
                                (7, 249), (64, 191)])
def test_tail_window_equals_full_decode(yx, monkeypatch):
    """Second level (loss.tail_window): the up-sampling blocks only see the crop's footprint + TAIL_CONTEXT latent
    pixels.  Crop pixels and gradients must still equal the whole-canvas decode; with one pixel less context they
    must not (the test would be blind otherwise)."""
    rng = np.random.RandomState(9)
    P = R.to_torch(R.init_params('G_res', rng, **CFG), requires_grad=True)
    H = W = LAT * S
    zg = torch.from_numpy(rng.randn(1, 2, 1, 1).astype(np.float32)).repeat(1, 1, H, W)
    zl = torch.from_numpy(rng.randn(1, 2, H, W).astype(np.float32)).requires_grad_(True)
    seed = torch.from_numpy(rng.randn(1, 3, RES, RES).astype(np.float32))
    full = R.G_res(zg, zl, P, **dict(CFG, scale_h=S, scale_w=S))[:, :, yx[0]:yx[0] + RES, yx[1]:yx[1] + RES]
    names = [k for k in P if k != 'lod']
    g_full = torch.autograd.grad((full * seed).sum(), [zl] + [P[k] for k in names], allow_unused=True)

    def windowed():
        win = crop_window(yx, RES, LAT, H, W)
        mid = mid_window(yx, RES, LAT, win, H, W)
        win_abs = compose_window(win, mid, H, W)
        tail = tail_window(yx, RES, LAT, win_abs, H, W)
        assert mid[2] == mid[3] == dev_loss.MID_SIZE and tail[2] == tail[3] == dev_loss.TAIL_SIZE
        oy, ox, wh, ww = win
        img = R.G_res(zg[:, :, :wh, :ww], zl[:, :, oy:oy + wh, ox:ox + ww], P,
                      **dict(CFG, scale_h=wh // LAT, scale_w=ww // LAT, mid_window=mid, tail_window=tail))
        assert tuple(img.shape[2:]) == (4 * tail[2], 4 * tail[3])
        y0, x0 = image_offset(yx, 4, win_abs, tail)
        return img[:, :, y0:y0 + RES, x0:x0 + RES]
    part = windowed()
    assert part.shape == full.shape and float((part - full).detach().abs().max()) <= 1e-5
    g_part = torch.autograd.grad((part * seed).sum(), [zl] + [P[k] for k in names], allow_unused=True)
    for name, a, b in zip(['zl'] + names, g_full, g_part):
        if a is None and b is None:
            continue
        assert float((a - b).abs().max()) <= 1e-4 * (float(a.abs().max()) + 1e-12), name
    if 8 <= yx[0] <= 240 and 8 <= yx[1] <= 240 and (yx[0] % 4 or yx[1] % 4):   # an interior, unaligned crop
        monkeypatch.setattr(dev_loss, 'TAIL_CONTEXT', 1)
        monkeypatch.setattr(dev_loss, 'TAIL_SIZE', 36)
        bad = windowed()
        assert float((bad - full).detach().abs().max()) > 1e-4
        monkeypatch.setattr(dev_loss, 'TAIL_CONTEXT', 2)
        monkeypatch.setattr(dev_loss, 'TAIL_SIZE', 40)
        monkeypatch.setattr(dev_loss, 'MID_CONTEXT', 4)          # the middle level alone with too little context
        monkeypatch.setattr(dev_loss, 'MID_SIZE', 43)
        bad = windowed()
        assert float((bad - full).detach().abs().max()) > 1e-4


def test_nested_window_geometry_all_offsets():
    """Every crop offset of the 3x3 canvas: the three nested windows contain what their remaining layers need, and
    an edge that is not an edge of the canvas keeps the required context (G_CONTEXT / MID_CONTEXT / TAIL_CONTEXT)."""
    H = W = LAT * S
    for y in range(0, RES * S - RES):
        win = crop_window((y, 77), RES, LAT, H, W)
        mid = mid_window((y, 77), RES, LAT, win, H, W)
        win_abs = compose_window(win, mid, H, W)
        tail = tail_window((y, 77), RES, LAT, win_abs, H, W)
        k0, k1 = y // 4, (y + RES - 1) // 4                         # latent rows of the crop's footprint
        lo = {}
        for name, (o, size), ctx in (('trunk', (win[0], win[2]), G_CONTEXT),
                                     ('mid', (win_abs[0], win_abs[2]), dev_loss.MID_CONTEXT),
                                     ('tail', (win_abs[0] + tail[0], tail[2]), dev_loss.TAIL_CONTEXT)):
            assert 0 <= o and o + size <= H, (name, y)
            assert o == 0 or k0 - o >= ctx, (name, y, o)
            assert o + size == H or (o + size - 1) - k1 >= ctx, (name, y, o, size)
            lo[name] = (o, o + size)
        assert lo['trunk'][0] <= lo['mid'][0] <= lo['tail'][0] and lo['tail'][1] <= lo['mid'][1] <= lo['trunk'][1]
        y_img, _ = image_offset((y, 77), 4, win_abs, tail)
        assert 0 <= y_img and y_img + RES <= 4 * tail[2]            # the crop lies inside the decoded image
