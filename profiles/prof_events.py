"""In-situ per-entry-point GPU time of one train step: CUDA events around every libtmx call (real clocks, warm caches;
ncu's launch list is cold-cache and lets the clocks drop between its serialised kernels)."""
import collections, sys, os
os.environ['TMX_NO_GRAPH'] = '1'
import numpy as np, torch
sys.path.insert(0, '.')
from texturemixer_b200 import _lib
from texturemixer_b200.train import Trainer
tr = Trainer(seed=1000, device=0)
rng = np.random.RandomState(0); np.random.seed(0)
reals = torch.from_numpy(rng.uniform(-1, 1, (32, 3, 128, 128)).astype(np.float32)).cuda()
for _ in range(4):
    tr.step(reals, tr.sample_draws(32, rng))
torch.cuda.synchronize()
lib = tr.rt.lib
records = []
orig = {}
def wrap(name, fn):
    def f(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*a)
        e1.record()
        tag = name
        if name == 'tmx_conv2d_fwd':
            d = a[1]._obj
            tag = 'conv2d_fwd %dx%d %d->%d k%d%s' % (d.H, d.W, d.Cin, d.Cout, d.k, ' up2' if d.flags & 4 else '')
        elif name in ('tmx_conv2d_dgrad', 'tmx_conv2d_wgrad'):
            tag = '%s %dx%d %d->%d k%d' % (name[4:], a[2], a[3], a[4], a[5], a[6])
        elif name == 'tmx_grad_prepare':
            d = a[1]._obj
            tag = 'grad_prepare %dx%d c%d' % (d.H, d.W, d.C)
        records.append((tag, name, e0, e1))
        return rc
    return f
for name in _lib.EXPORTED_SYMBOLS:
    if name in ('tmx_abi_version', 'tmx_last_error', 'tmx_create', 'tmx_destroy', 'tmx_device_info', 'tmx_launch_count',
                'tmx_dense_workspace_bytes', 'tmx_conv2d_wgrad_workspace_bytes', 'tmx_perm_indices_from_uniforms'):
        continue
    orig[name] = getattr(lib, name)
    setattr(lib, name, wrap(name, orig[name]))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
tr.step(reals, tr.sample_draws(32, rng))
s1.record()
torch.cuda.synchronize()
tot = s0.elapsed_time(s1)
by_name, by_tag = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(lambda: [0, 0.0])
for tag, name, e0, e1 in records:
    ms = e0.elapsed_time(e1)
    by_name[name][0] += 1; by_name[name][1] += ms
    by_tag[tag][0] += 1; by_tag[tag][1] += ms
inside = sum(v[1] for v in by_name.values())
print('step %.2f ms (with event overhead); inside libtmx calls %.2f ms; %d calls' % (tot, inside, len(records)))
for k, (c, v) in sorted(by_name.items(), key=lambda x: -x[1][1])[:25]:
    print('%8.3f ms %5.1f%% %5d  %s' % (v, 100 * v / inside, c, k))
print()
for k, (c, v) in sorted(by_tag.items(), key=lambda x: -x[1][1])[:60]:
    print('%8.3f ms %5.1f%% %5d  avg %7.1f us  %s' % (v, 100 * v / inside, c, 1e3 * v / c, k))
