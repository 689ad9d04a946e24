"""Latent-tile interpolation / blend: the host-side mirror of
  loss.py:92-100        tiling_permutation
  run.py:107-182,436-507 permutation sampler (index-vector form, same RNG stream)
  tfutil.py:41-43       lerp
  util_scripts.py:85-102 linear mattes, :496,525,738,766 4-corner weighted sums
over the one device kernel `tmx_latent_blend` (include/tmx.h).  Permutations
travel as int32 index vectors (the "integer tile-index grid"): row map r with
P_h[i, r[i]] = 1 and column map c with P_w[c[j], j] = 1, so that
P_h @ tile(X) @ P_w == tile(X)[r][:, c] bit for bit."""
import numpy as np
import torch

from . import _lib
from .runtime import Runtime, perm_indices_from_uniforms, uniforms_per_matrix


# ---------------------------------------------------------------------- sampler (host)
def sample_permutation_indices(count, length, levels, uniform=None):
    """`count` index vectors of `length` from the hierarchical swap sampler
    (my_swap_h/w + block_permutation over block sizes 2^0..2^(levels-1)).
    Consumes exactly the uniforms the reference's Python loops would draw from
    `np.random` (2*(length>>l) per level with length>>l > 1, in order), so the
    result is bit-identical to argmax of the reference's matrices.
    `uniform(size=n)` defaults to the global legacy MT19937 stream (run.py:599)."""
    if uniform is None:
        uniform = np.random.uniform
    n = uniforms_per_matrix(length, levels) * count
    u = uniform(size=n) if n > 0 else np.zeros(0)
    idx, used = perm_indices_from_uniforms(u, length, levels, count)
    assert used == n
    return idx


def _block_expand(r, block_size):
    """Row map of block_permutation(perm, block_size) (run.py:176-182) from the row map of `perm`."""
    r = np.asarray(r, np.int64)
    return (r[:, None] * block_size + np.arange(block_size)[None, :]).reshape(-1)


def sample_permutation_indices_general(count, length, levels, axis, block_size=0, perm=False, uniform=None):
    """The config-off branches of the scheduler loop (run.py:436-507) in index form:
    `block_size` > 0 (config.block_size: ONE swap / permutation of length/block_size blocks instead of the hierarchy)
    and `perm` (config.perm: np.random.permutation of the identity instead of the local swaps my_swap_h/w).
    axis 'h': row map r with P[i, r[i]] = 1 (matrices composed as temp @ block, run.py:447);
    axis 'w': column map c with P[c[j], j] = 1 (composed as block @ temp, run.py:464).
    Draws from the same legacy np.random stream, in the same order, as the reference's loops."""
    assert axis in ('h', 'w')
    if not perm and block_size == 0:
        return sample_permutation_indices(count, length, levels, uniform)
    out = np.empty((count, length), np.int32)

    def one_level(n):
        """Row map of `perm` for one level over n blocks."""
        if perm:
            # np.random.permutation(np.eye(n)) shuffles arange(n) and gathers rows: mat[i] = eye[p[i]]
            return np.random.permutation(n).astype(np.int64)
        idx = sample_permutation_indices(1, n, 1, uniform)[0].astype(np.int64)   # my_swap_h / my_swap_w of eye(n)
        # sample_permutation_indices returns the map its axis convention reads; with one level the h and w forms
        # describe the same matrix family: for 'w' it is the column map, whose inverse is the row map
        return idx

    for k in range(count):
        if block_size > 0:
            m = one_level(length // block_size)
            if perm or axis == 'h':
                r = _block_expand(m, block_size)
                out[k] = r if axis == 'h' else np.argsort(r)
            else:
                out[k] = _block_expand(m, block_size)      # already the column map of my_swap_w's matrix
            continue
        r = np.arange(length, dtype=np.int64)               # row map of temp_perm
        for lvl in range(levels):
            bs = 1 << lvl
            m = one_level(length // bs)
            rb = _block_expand(m, bs)
            r = rb[r] if axis == 'h' else r[rb]             # temp @ blk : blk @ temp
        out[k] = r if axis == 'h' else np.argsort(r)
    return out


def sample_schedule_indices(minibatch, latent_res=32, scale_h=3, scale_w=3, levels=None, uniform=None, block_size=0,
                            perm=False):
    """One scheduler iteration (run.py:436-507): h_forward, w_forward, h_backward,
    w_backward, each int32 [minibatch, latent_res*scale]; levels = int(log2(latent_res))
    as run.py:440 (the inference apps use int(np.log(latent_res)), util_scripts.py:405).
    `block_size` / `perm`: config.block_size / config.perm (0 / False in the reference config)."""
    if levels is None:
        levels = int(np.log2(latent_res))
    lh, lw = latent_res * scale_h, latent_res * scale_w
    out = {}
    for name, ln in (('h_forward', lh), ('w_forward', lw), ('h_backward', lh), ('w_backward', lw)):
        if block_size or perm:
            out[name] = sample_permutation_indices_general(minibatch, ln, levels, name[0], block_size, perm, uniform)
        else:
            out[name] = sample_permutation_indices(minibatch, ln, levels, uniform)
    return out


def indices_from_matrices(p_h, p_w):
    """Reference-format permutation matrices [N,1,H,H], [N,1,W,W] (0/1 floats fed at
    run.py:289-292) -> (row map [N,H], column map [N,W]) int32."""
    p_h = np.asarray(p_h)
    p_w = np.asarray(p_w)
    r = np.argmax(p_h.reshape(p_h.shape[0], p_h.shape[-2], p_h.shape[-1]), axis=2).astype(np.int32)
    c = np.argmax(p_w.reshape(p_w.shape[0], p_w.shape[-2], p_w.shape[-1]), axis=1).astype(np.int32)
    return r, c


# ---------------------------------------------------------------------- mattes (host, float64 like numpy)
def _ramp(length, latent_res):
    mid = np.linspace(start=0.0, stop=1.0, num=length - 2 * latent_res)[::-1]
    return np.concatenate((np.ones(latent_res), mid, np.zeros(latent_res)))


def linkern_for_weight_horizontal(out_shape, latent_res):
    """util_scripts.py:85-89: float32 [N,C,H,W], 1 -> 0 left to right."""
    k = _ramp(out_shape[3], latent_res).reshape(1, 1, 1, -1)
    return np.tile(k, list(out_shape[:3]) + [1]).astype(np.float32)


def linkern_ramps(out_h, out_w, latent_res):
    """The separable form of util_scripts.py:91-102: per corner (UL, UR, BL, BR) the
    float64 row ramp and column ramp whose outer product is the corner's matte."""
    kh, kw = _ramp(out_h, latent_res), _ramp(out_w, latent_res)
    return [kh, kh, 1.0 - kh, 1.0 - kh], [kw, 1.0 - kw, kw, 1.0 - kw]


def linkern_for_weight_arbitrary_shape(out_h, out_w, latent_res):
    """util_scripts.py:91-102 -> (weight_ul, weight_ur, weight_bl, weight_br) float64 [out_h,out_w]."""
    rh, rw = linkern_ramps(out_h, out_w, latent_res)
    return tuple(a[:, None] * b[None, :] for a, b in zip(rh, rw))


# ---------------------------------------------------------------------- app-level mattes (host, float64 like numpy)
def gkern_for_weight_arbitrary_shape(out_h, out_w, x, y, sig_div):
    """util_scripts.py:104-114 (texture brush strokes, :844-868): a horizontally stretched Gaussian around (x, y),
    sigma = out_w / sig_div, min-max normalised to [0, 1].  float64 [out_h, out_w]."""
    cx, cy = float(x), float(y)
    sig = float(out_w) / sig_div
    xx, yy = np.meshgrid(np.arange(0.0, float(out_w)), np.arange(0.0, float(out_h)))
    kernel = np.exp(-(np.maximum(np.absolute(xx - cx) - sig, 0.0) ** 2 + (yy - cy) ** 2) / (2. * sig ** 2))
    return (kernel - np.amin(kernel)) / (np.amax(kernel) - np.amin(kernel))


def gkern_for_weight_arbitrary_shape_hybridization(out_h, out_w, x, y, sig_div):
    """util_scripts.py:116-125: plain Gaussian around (x, y), not normalised."""
    cx, cy = float(x), float(y)
    sig = float(out_w) / sig_div
    xx, yy = np.meshgrid(np.arange(0.0, float(out_w)), np.arange(0.0, float(out_h)))
    return np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2. * sig ** 2))


def gkern_for_weight_grid_shape_hybridization(out_h, out_w, cx, cy, size, sig_div):
    """util_scripts.py:127-168 (hybridization RBF weights, :1145): exp(-d^2 / 2 sigma^2) with d the distance of
    latent pixel (j, i) to the square [cx, cx+size] x [cy, cy+size], sigma = min(out_h, out_w) / sig_div.
    The reference fills it with a Python double loop over `dist2square`; this is the same arithmetic on arrays."""
    cx, cy = float(cx), float(cy)
    sig = min([float(out_h), float(out_w)]) / sig_div
    x = np.arange(out_w, dtype=np.float64)[None, :]
    y = np.arange(out_h, dtype=np.float64)[:, None]
    x_max, y_max = cx + size, cy + size
    dx = np.where(x < cx, x - cx, np.where(x <= x_max, 0.0, x - x_max))
    dy = np.where(y < cy, y - cy, np.where(y <= y_max, 0.0, y - y_max))
    return np.exp(-(dx ** 2 + dy ** 2) / (2. * sig ** 2))


def linkern_for_weight_square(out_length, latent_res):
    """util_scripts.py:53-62 -> (weight_ul, weight_ur, weight_bl, weight_br) float64 [L, L]."""
    step = 1.0 / (out_length - 2.0 * latent_res - 1.0)
    ax = np.arange(start=0.0, stop=1.0 + step, step=step)
    ax = np.concatenate((np.zeros(latent_res), ax, np.ones(latent_res)))
    X, Y = np.meshgrid(ax, ax)
    weight_br = X * Y
    weight_ur = np.rot90(weight_br)
    weight_ul = np.rot90(weight_ur)
    weight_bl = np.rot90(weight_ul)
    return weight_ul, weight_ur, weight_bl, weight_br


def gkern_for_scale_horizontal(out_shape, latent_res):
    """util_scripts.py:170-182: float32 [N,C,H,W], 0 over the first latent_res columns, 1 elsewhere."""
    kernel = np.concatenate((np.zeros((1, 1, 1, latent_res)), np.ones((1, 1, 1, out_shape[3] - latent_res))), axis=3)
    return kernel.astype(np.float32)


def weighted_sum(sources, weights, math_f32=False):
    """sum_k sources[k] * weights[k] on the device: the matte compositing of the apps (util_scripts.py:1262,1268
    hybridization `np.sum(latents * weights, axis=0)`; :1337,1342 horizontal mattes; :844-868 brush strokes).
      sources: K device tensors [N,C,H,W] (or [N,C,1,1] = a global code tiled over the canvas)
      weights: [K,H,W] array (float64 like the reference's numpy products, rounded once to float32 at the end;
               math_f32=True multiplies and adds in float32 like a float32 matte does)
    -> [N,C,H,W] float32."""
    import ctypes as C
    rt = Runtime.get(sources[0].device)
    k = len(sources)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    assert weights.ndim == 3 and weights.shape[0] == k
    H, W = weights.shape[1:]
    n, c = sources[0].shape[:2]
    bcast = []
    srcs = []
    for s in sources:
        assert s.dtype == torch.float32 and s.is_cuda and s.shape[:2] == (n, c)
        if tuple(s.shape[2:]) == (1, 1) and (H, W) != (1, 1):
            bcast.append(1)
        else:
            assert tuple(s.shape[2:]) == (H, W)
            bcast.append(0)
        srcs.append(s.contiguous())
    ptrs = torch.tensor([s.data_ptr() for s in srcs], dtype=torch.int64).to(rt.device)
    flags = torch.tensor(bcast, dtype=torch.int32).to(rt.device)
    wd = torch.from_numpy(weights).to(rt.device)
    out = rt.empty(n, c, H, W)
    _lib.check(rt.lib.tmx_weighted_sum(rt.handle, C.c_void_p(ptrs.data_ptr()), C.c_void_p(flags.data_ptr()),
                                       C.c_void_p(wd.data_ptr()), C.c_void_p(out.data_ptr()), k, n, c, H, W,
                                       int(math_f32), rt.stream()), 'tmx_weighted_sum')
    return out


# ---------------------------------------------------------------------- device ops
def _dev_idx(rt, idx, n, length):
    if idx is None:
        return None
    if isinstance(idx, torch.Tensor):
        t = idx.to(device=rt.device, dtype=torch.int32).contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to(rt.device)
    if tuple(t.shape) != (n, length):
        raise ValueError('index vectors must be [%d,%d], got %s' % (n, length, tuple(t.shape)))
    return t


def _corner_pins(scale_h, scale_w):
    return (1 | (1 << (scale_h - 1))), (1 | (1 << (scale_w - 1)))


def tiling_permutation(tensor, scale_h, scale_w, permutation_h, permutation_w, pin_corners=True):
    """loss.py:92-100 on a device tensor [N,C,h,w] -> [N,C,h*scale_h,w*scale_w].
    `permutation_h/w`: int32 index vectors [N,h*scale_h] / [N,w*scale_w], or the
    reference's 0/1 matrices [N,1,H,H] / [N,1,W,W] (converted on the host)."""
    rt = Runtime.get(tensor.device)
    n, c, h, w = tensor.shape
    H, W = h * scale_h, w * scale_w
    if getattr(permutation_h, 'ndim', 2) == 4:
        ph = permutation_h.cpu().numpy() if isinstance(permutation_h, torch.Tensor) else permutation_h
        pw = permutation_w.cpu().numpy() if isinstance(permutation_w, torch.Tensor) else permutation_w
        permutation_h, permutation_w = indices_from_matrices(ph, pw)
    ih, iw = _dev_idx(rt, permutation_h, n, H), _dev_idx(rt, permutation_w, n, W)
    pr, pc = _corner_pins(scale_h, scale_w) if pin_corners else (0, 0)
    return rt.latent_blend([tensor.contiguous()], H, W, _lib.BLEND_COPY, idx_h=[ih], idx_w=[iw], pin_rows=pr,
                           pin_cols=pc)


def lerp(a, b, t):
    """tfutil.py:41-43 with a per-sample factor t [N,1,1,1] (loss.py:237-239): a + (b - a) * t."""
    rt = Runtime.get(a.device)
    n, c, h, w = a.shape
    tt = t.reshape(-1).to(device=rt.device, dtype=torch.float32).contiguous()
    return rt.latent_blend([a.contiguous(), b.contiguous()], h, w, _lib.BLEND_LERP, t=tt)


def blend_corners(sources, out_h, out_w, latent_res, idx_h=None, idx_w=None, pin_rows=0, pin_cols=0):
    """sum_k matte_k * canvas_k over the four corner sources UL, UR, BL, BR
    (util_scripts.py:496,525 / 738,766): canvas_k = tiled (+ permuted, re-pinned)
    source k, matte_k = linkern_for_weight_arbitrary_shape, float64 products
    and sums rounded once to float32 like numpy's promotion in the reference.
    sources: four [N,C,h,w] device tensors ([N,C,1,1] = tiled global code)."""
    assert len(sources) == 4
    rt = Runtime.get(sources[0].device)
    n = sources[0].shape[0]
    rh, rw = linkern_ramps(out_h, out_w, latent_res)
    rh = [torch.from_numpy(np.ascontiguousarray(r)).to(rt.device) for r in rh]
    rw = [torch.from_numpy(np.ascontiguousarray(r)).to(rt.device) for r in rw]
    ih = None if idx_h is None else [_dev_idx(rt, i, n, out_h) for i in idx_h]
    iw = None if idx_w is None else [_dev_idx(rt, i, n, out_w) for i in idx_w]
    return rt.latent_blend([s.contiguous() for s in sources], out_h, out_w, _lib.BLEND_MATTE, idx_h=ih, idx_w=iw,
                           ramps_h=rh, ramps_w=rw, pin_rows=pin_rows, pin_cols=pin_cols)


def interpolate(zg_sources, zl_sources, scale_h, scale_w, idx_h=None, idx_w=None, latent_res=None,
                pin_corners=True):
    """north_star `interpolate`: the 4-corner latent canvas pair (zg, zl) that
    `G_res(scale_h, scale_w)` decodes into an interpolated texture
    (util_scripts.py:722-787 pattern).
      zg_sources: four [N,C,1,1] global codes  (UL, UR, BL, BR)
      zl_sources: four [N,C,h,w] local codes
      idx_h/idx_w: per-source lists of int32 [N,h*scale_h] / [N,w*scale_w] (None = identity tiling)
    -> (zg_canvas, zl_canvas), each [N,C,h*scale_h,w*scale_w] float32 on the device."""
    h, w = zl_sources[0].shape[2:]
    latent_res = h if latent_res is None else latent_res
    H, W = h * scale_h, w * scale_w
    pr, pc = _corner_pins(scale_h, scale_w) if pin_corners else (0, 0)
    zg = blend_corners(zg_sources, H, W, latent_res)
    zl = blend_corners(zl_sources, H, W, latent_res, idx_h=idx_h, idx_w=idx_w, pin_rows=pr, pin_cols=pc)
    return zg, zl
