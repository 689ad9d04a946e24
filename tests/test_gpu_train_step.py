"""GPU check of the whole train step (texturemixer_b200.train.Trainer.step == run.py:510-514): critics with pre-step
E/G, then E/G with the post-step critics, then EMA - against the oracle (autograd losses + TF1 Adam restatement)."""
import numpy as np
import pytest
import torch

from oracle import loss_ref as L
from oracle import networks_ref as R
from oracle import optim_ref as O

pytestmark = pytest.mark.gpu


def _flat(params):
    return np.concatenate([np.asarray(v, np.float32).reshape(-1) for k, v in params.items() if k != 'lod'])


def _unflat(params, flat):
    out, off = type(params)(), 0
    for k, v in params.items():
        if k == 'lod':
            out[k] = v
            continue
        n = int(np.prod(np.shape(v)))
        out[k] = flat[off:off + n].reshape(np.shape(v)).astype(np.float32)
        off += n
    return out


def _grads(P):
    return np.concatenate([(t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape))).astype(np.float32).reshape(-1)
                           for k, t in P.items() if k != 'lod'])


@pytest.mark.parametrize('lod', [0.0, 1.5])
def test_train_step_vs_oracle(lod):
    """lod = 1.5: the same step in the middle of a progressive-growing fade (run.py:310: every network at lod 1.5 -
    64x64 heads blended with the 32x32 ones; critics evaluated without CUDA graphs since the fade is baked in)."""
    from texturemixer_b200.train import Trainer, default_config, NET_FUNCS
    cfg = default_config(scale_h=2, scale_w=2)
    tr = Trainer(cfg, seed=1000)
    n = 4
    rng = np.random.RandomState(3)
    reals = rng.uniform(-1, 1, (n, 3, 128, 128)).astype(np.float32)
    np.random.seed(1000)
    draws = tr.sample_draws(n, rng)
    names = list(NET_FUNCS)
    ofunc = dict(E_zg='E_zg', E_zl='E_zl', G='G_res', D_rec='D_patch', D_interp='D_patch', D_blend='D_patch')
    # non-zero biases on both sides; oracle parameters = the trainer's initial variables
    params = {}
    for k in names:
        net = tr.nets[k]
        for vn, v in net.trainables.items():
            if vn.endswith('/bias'):
                net.set_var(vn, 0.1 * rng.randn(*v.shape).astype(np.float32))
        params[k] = type(R.init_params(ofunc[k], np.random.RandomState(0), **R.CONFIG[ofunc[k]]))(
            (vn, net.get_var(vn)) for vn in net.vars)
        params[k]['lod'] = np.float32(lod)
    for src, dst in (('E_zg', 'Es_zg'), ('E_zl', 'Es_zl'), ('G', 'Gs')):
        tr.nets[dst].copy_vars_from(tr.nets[src])
    w0 = {k: _flat(params[k]) for k in names}
    cfgs = dict(R.CONFIG)
    res = 128
    mixes = {k: draws[k].cpu() for k in ('eg_mix', 'd_rec_gp', 'd_interp_gp', 'd_blend_mix', 'd_blend_gp')}

    # ---------------- oracle step
    x = torch.from_numpy(reals)
    with torch.no_grad():
        P0 = {k: R.to_torch(params[k]) for k in names}
        zg, _ = R.E_zg(x, P0['E_zg'], **cfgs['E_zg'])
        zl, _ = R.E_zl(x, P0['E_zl'], **cfgs['E_zl'])
        rec = R.G_res(zg.repeat(1, 1, 32, 32), zl, P0['G'], **cfgs['G_res'])
        gcfg = dict(cfgs['G_res'], scale_h=2, scale_w=2)
        zg_c = zg.repeat(1, 1, 64, 64)
        zl_c = L.tiling_permutation(zl, 2, 2, draws['idx']['h_forward'], draws['idx']['w_forward'])
        y0, x0 = draws['d_interp_crop']
        fake_i = R.G_res(zg_c, zl_c, P0['G'], **gcfg)[:, :, y0:y0 + res, x0:x0 + res]
        zg_r = torch.flip(zg, dims=[0]).repeat(1, 1, 64, 64)
        zl_r = L.tiling_permutation(torch.flip(zl, dims=[0]), 2, 2, draws['idx']['h_backward'], draws['idx']['w_backward'])
        t = mixes['d_blend_mix']
        y0, x0 = draws['d_blend_crop']
        fake_b = R.G_res(zg_r + (zg_c - zg_r) * t, zl_r + (zl_c - zl_r) * t, P0['G'], **gcfg)[:, :, y0:y0 + res, x0:x0 + res]
    new = {}
    for k, fake, gp in (('D_rec', rec, 'd_rec_gp'), ('D_interp', fake_i, 'd_interp_gp'), ('D_blend', fake_b, 'd_blend_gp')):
        P = R.to_torch(params[k], requires_grad=True)
        loss, _ = L.D_wgangp(P, fake, x, mixes[gp])
        loss.mean().backward()
        w = w0[k].copy()
        assert O.optimizer_step(w, [_grads(P)], O.AdamState(w.size, 0.0, 0.99), 0.0015)
        new[k] = w
    P = {k: R.to_torch(params[k], requires_grad=True) for k in ('E_zg', 'E_zl', 'G')}
    for k in ('D_rec', 'D_interp', 'D_blend'):
        P[k] = R.to_torch(_unflat(params[k], new[k]))
    loss, _ = L.EG_wgan(P, x, draws['idx'], draws['eg_crop_interp'], draws['eg_crop_blend'], mixes['eg_mix'], scale_h=2,
                        scale_w=2)
    loss.mean().backward()
    g_all = np.concatenate([_grads(P[k]) for k in ('E_zg', 'E_zl', 'G')])
    w_all = np.concatenate([w0[k] for k in ('E_zg', 'E_zl', 'G')])
    assert O.optimizer_step(w_all, [g_all], O.AdamState(w_all.size, 0.0, 0.99), 0.0015)
    off = 0
    for k in ('E_zg', 'E_zl', 'G'):
        new[k] = w_all[off:off + w0[k].size]
        off += w0[k].size

    # ---------------- device step
    rep = tr.step(torch.from_numpy(reals).cuda(), draws, lod=lod)
    torch.cuda.synchronize()
    assert all(int(rep[k + '/skipped'].item()) == 0 for k in ('D_rec', 'D_interp', 'D_blend', 'EG'))
    step = 0.0015 * np.sqrt(1 - 0.99)      # first Adam step moves every weight by ~ lr_t * 10 * sign(g) = 1.5e-3
    for k in names:
        got = np.concatenate([tr.nets[k].get_var(vn).reshape(-1) for vn in tr.nets[k].vars if vn != 'lod'])
        moved = np.abs(new[k] - w0[k]) > 1e-6
        bad = np.abs(got - new[k]) > 1e-4
        # the first Adam step is +-1.5e-3 * sign(g): only gradients within the (leaky-ReLU-flip limited) device
        # error of zero may land on the other side
        assert bad.mean() <= 0.02, (k, float(bad.mean()))
        # (at lod 1.5 the 128x128 blocks and lod-0 heads take no part: their gradients are zero on both sides)
        assert moved.mean() > (0.9 if lod == 0 else 0.75) and np.abs(got - w0[k]).max() <= 1.05 * step * 10 + 1e-6
    # EMA: Gs = lerp(G, Gs, 0.999) with Gs initialised to the pre-step G
    g_new = np.concatenate([tr.nets['G'].get_var(vn).reshape(-1) for vn in tr.nets['G'].vars if vn != 'lod'])
    gs = np.concatenate([tr.nets['Gs'].get_var(vn).reshape(-1) for vn in tr.nets['Gs'].vars if vn != 'lod'])
    assert np.abs(gs - (g_new + (w0['G'] - g_new) * np.float32(0.999))).max() <= 1e-6
