"""CPU-only checks of the host side: the C-ABI library loads and exports every
symbol include/tmx.h declares, the host permutation sampler (C) is bit-exact
against the reference-generated goldens, mattes match, the `Network` boundary
mirrors the reference's template-graph bookkeeping, and the product fails
loudly without a GPU (no CPU fallback)."""
import os
import pickle
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

CFG = dict(
    E_zg=dict(fmap_base=1024, fmap_max=512, latent_channels=128, use_pixelnorm=False, tanh_at_end=False),
    E_zl=dict(fmap_base=1024, fmap_max=512, latent_res=32, latent_channels=128, use_pixelnorm=False,
              tanh_at_end=False),
    G_res=dict(fmap_base=1024, fmap_max=512, latent_res=32, latent_channels=128, use_pixelnorm=False,
               tanh_at_end=True),
    D_patch=dict(fmap_base=1024, fmap_max=512, latent_res=-1),
)


@pytest.fixture(scope='module')
def built():
    from texturemixer_b200 import build
    return build.build_library()


def test_library_exports_every_declared_symbol(built):
    import ctypes
    from texturemixer_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'tmx.h')).read()
    declared = set(re.findall(r'\b(tmx_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    lib = ctypes.CDLL(built)
    for name in sorted(declared):
        assert hasattr(lib, name), 'libtmx.so does not export %s' % name
    assert declared == set(_lib.EXPORTED_SYMBOLS), (declared ^ set(_lib.EXPORTED_SYMBOLS))
    assert _lib.load().tmx_abi_version() == _lib.TMX_ABI_VERSION


@pytest.mark.parametrize('tag', ['train96', 'interp128', 'small8', 'tiny2', 'wide256'])
def test_c_sampler_bit_exact_vs_reference_golden(built, tag):
    from texturemixer_b200 import interp
    g = np.load(os.path.join(GOLDEN, 'perm_sampler.npz'))
    length, levels, count, seed = (int(v) for v in g[tag + '_meta'])
    np.random.seed(seed)
    hs = interp.sample_permutation_indices(count, length, levels)
    ws = interp.sample_permutation_indices(count, length, levels)
    assert hs.dtype == np.int32 and np.array_equal(hs, g[tag + '_h'])
    assert np.array_equal(ws, g[tag + '_w'])
    assert np.random.uniform() == g[tag + '_next_uniform'][0]      # same number of draws consumed


@pytest.mark.parametrize('tag', ['bs4', 'perm', 'bs8perm', 'bs32'])
def test_sampler_block_size_and_perm_options_vs_reference_golden(built, tag):
    """config.block_size > 0 / config.perm (run.py:436-507, off in the reference config): index vectors and the number
    of np.random draws equal to the reference's loops (tests/golden/make_golden.py gen_perm_options)."""
    from texturemixer_b200 import interp
    g = np.load(os.path.join(GOLDEN, 'perm_sampler_options.npz'))
    length, levels, count, seed, block_size, perm = (int(v) for v in g[tag + '_meta'])
    np.random.seed(seed)
    hs = interp.sample_permutation_indices_general(count, length, levels, 'h', block_size, bool(perm))
    ws = interp.sample_permutation_indices_general(count, length, levels, 'w', block_size, bool(perm))
    assert hs.dtype == np.int32 and np.array_equal(hs, g[tag + '_h'])
    assert np.array_equal(ws, g[tag + '_w'])
    assert np.random.uniform() == g[tag + '_next_uniform'][0]


def test_c_sampler_schedule_order_and_errors(built):
    from texturemixer_b200 import interp
    from texturemixer_b200.runtime import perm_indices_from_uniforms, uniforms_per_matrix
    assert uniforms_per_matrix(96, 5) == 372
    np.random.seed(1000)
    s = interp.sample_schedule_indices(2, latent_res=32, scale_h=3, scale_w=3)
    assert list(s.keys()) == ['h_forward', 'w_forward', 'h_backward', 'w_backward']
    g = np.load(os.path.join(GOLDEN, 'perm_sampler.npz'))
    assert np.array_equal(s['h_forward'], g['train96_h'][:2])
    with pytest.raises(RuntimeError, match='uniforms needed'):
        perm_indices_from_uniforms(np.zeros(10), 96, 5, 1)
    with pytest.raises(RuntimeError, match='divisible'):
        perm_indices_from_uniforms(np.zeros(1000), 6, 3, 1)
    idx, used = perm_indices_from_uniforms(np.zeros(0), 1, 1, 2)     # degenerate length: no draws
    assert used == 0 and idx.tolist() == [[0], [0]]


def test_indices_from_matrices():
    from texturemixer_b200 import interp
    r = np.array([[2, 0, 1]], np.int32)
    c = np.array([[1, 2, 0]], np.int32)
    ph = np.zeros((1, 1, 3, 3), np.float32)
    pw = np.zeros((1, 1, 3, 3), np.float32)
    ph[0, 0, np.arange(3), r[0]] = 1
    pw[0, 0, c[0], np.arange(3)] = 1
    r2, c2 = interp.indices_from_matrices(ph, pw)
    assert np.array_equal(r, r2) and np.array_equal(c, c2)


def test_mattes_bit_exact_vs_reference_golden():
    from texturemixer_b200 import interp
    g = np.load(os.path.join(GOLDEN, 'mattes.npz'))
    for key in g.files:
        parts = key.split('_')
        if parts[0] == 'arb':
            h, w, r = (int(p) for p in parts[1:])
            mine = np.stack(interp.linkern_for_weight_arbitrary_shape(h, w, r))
            assert mine.dtype == np.float64 and np.array_equal(mine, g[key])
        else:
            shape = [int(p) for p in parts[1:5]]
            mine = interp.linkern_for_weight_horizontal(shape, int(parts[5]))
            assert mine.dtype == np.float32 and np.array_equal(mine, g[key])


@pytest.mark.parametrize('func', ['E_zg', 'E_zl', 'G_res', 'D_patch'])
def test_template_graph_matches_reference_variables(func):
    """Variable names + creation order equal those of the reference's networks.py
    (golden minted by running it, tests/golden/make_golden.py)."""
    from texturemixer_b200.network import Network
    g = np.load(os.path.join(GOLDEN, 'networks.npz'))
    net = Network(func, func='networks.' + func, device='cpu', seed=0, num_channels=3, resolution=128, **CFG[func])
    assert list(net.vars.keys()) == [str(s) for s in g[func + '_varnames']]
    assert 'lod' not in net.trainables and len(net.trainables) == len(net.vars) - 1
    shapes = {'E_zg': ([[None, 3, 128, 128]], [[None, 128, 1, 1]] * 2, ['zg_mu', 'zg_log_sigma']),
              'E_zl': ([[None, 3, 128, 128]], [[None, 128, 32, 32]] * 2, ['z_mu', 'z_log_sigma']),
              'G_res': ([[None, 128, 32, 32]] * 2, [[None, 3, 128, 128]], ['images_out']),
              'D_patch': ([[None, 3, 128, 128]], [[None, 1, 1, 1]], ['scores_out'])}[func]
    assert net.input_shapes == shapes[0] and net.output_shapes == shapes[1] and net.output_names == shapes[2]


def test_network_boundary_bookkeeping():
    from texturemixer_b200.network import Network
    G = Network('G', func='networks.G_res', device='cpu', seed=1, num_channels=3, resolution=128, **CFG['G_res'])
    assert G.input_names == ['zg_latents_in', 'zl_latents_in'] and G.num_outputs == 1
    n_train = sum(v.size for v in G.trainables.values())
    assert n_train == 6119955 + (32 * 3 + 3) + (64 * 3 + 3)            # SURVEY a6 + unused lod heads
    # fully-convolutional second view over the same scope (run.py:273)
    G_fcn = Network('G', func='networks.G_res', reuse=True, share_vars_with=G, device='cpu', num_channels=3,
                    resolution=128, scale_h=3, scale_w=3, **CFG['G_res'])
    assert G_fcn.vars is G.vars and G_fcn.input_shape == [None, 128, 96, 96]
    assert G_fcn.output_shape == [None, 3, 384, 384]
    # clone + EMA (tfutil.py:579-589, 611-621)
    Gs = G.clone('Gs')
    assert Gs.name == 'Gs' and np.array_equal(Gs.get_var('32x32/Conv0/weight'), G.get_var('32x32/Conv0/weight'))
    G.set_var('32x32/Conv0/bias', np.ones(64, np.float32))
    Gs.setup_as_moving_average_of(G, beta=0.75)()
    assert np.allclose(Gs.get_var('32x32/Conv0/bias'), 0.25)
    # version-2 pickle state (tfutil.py:543-550) round trip
    state = G.__getstate__()
    assert state['version'] == 2 and state['build_func_name'] == 'G_res'
    assert [k for k, _ in state['variables']] == list(G.vars.keys())
    G2 = pickle.loads(pickle.dumps(G))
    assert np.array_equal(G2.get_var('ToRGB_lod0/weight'), G.get_var('ToRGB_lod0/weight'))
    with pytest.raises(AssertionError):
        Network('bad', func='networks.G_res', device='cpu', num_channels=3, resolution=100, **CFG['G_res'])
    with pytest.raises(NotImplementedError):                     # variants the device path does not implement are loud
        Network('bad', func='networks.G_res', device='cpu', num_channels=3, resolution=128, use_wscale=False,
                **CFG['G_res'])
    fused = Network('G', func='networks.G_res', device='cpu', num_channels=3, resolution=128, fused_scale=True,
                    **CFG['G_res'])                              # fused_scale: the transposed-conv variable layout
    assert fused.vars['64x64/Conv0_up/weight'].shape == (3, 3, 32, 64) and '64x64/Conv0/weight' not in fused.vars


def test_aliases_resolve():
    from texturemixer_b200 import network
    from texturemixer_b200 import networks
    assert network.import_obj('networks.G_res') is networks.G_res
    assert network.import_obj('networks.build_generator') is networks.build_generator
    E = network.Network('E', func='networks.build_encoder', device='cpu', kind='zl', num_channels=3, resolution=128,
                        **CFG['E_zl'])
    assert E.output_names == ['z_mu', 'z_log_sigma']


def test_no_cpu_fallback():
    """Without CUDA the product path must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from texturemixer_b200.network import Network
    from texturemixer_b200.runtime import Runtime
    with pytest.raises(RuntimeError, match='no CPU path'):
        Runtime.get()
    G = Network('G', func='networks.G_res', device='cpu', seed=1, num_channels=3, resolution=128, **CFG['G_res'])
    z = np.zeros((1, 128, 32, 32), np.float32)
    with pytest.raises(RuntimeError):
        G.run(z, z)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'texturemixer_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f


def test_reference_checkpoint_round_trip(tmp_path):
    """misc.load_pkl / save_pkl (misc.py:27-33): a 9-tuple pickled the way the reference does it - instances of
    `tfutil.Network` whose state is the version-2 dict of tfutil.py:543-550 - loads into this package's Network by
    variable name, without executing the stored module source; saving writes the same class reference back."""
    import pickle
    import pickletools
    import sys
    import types
    from oracle import networks_ref as R
    from texturemixer_b200 import misc
    from texturemixer_b200.network import Network

    funcs = dict(E_zg='E_zg', E_zl='E_zl', G='G_res', D_rec='D_patch', D_interp='D_patch', D_blend='D_patch',
                 Es_zg='E_zg', Es_zl='E_zl', Gs='G_res')
    rng = np.random.RandomState(5)
    params = {k: R.init_params(f, rng, **R.CONFIG[f]) for k, f in funcs.items()}
    params['Gs']['lod'] = np.float32(1.5)

    mod = types.ModuleType('tfutil')

    class RefNetwork:                                   # stand-in for the reference class: only its pickled form
        def __init__(self, name, func, variables, kwargs):
            self.state = {'version': 2, 'name': name, 'static_kwargs': dict(kwargs, num_channels=3, resolution=128),
                          'build_module_src': 'raise SystemExit("module source must never be executed")',
                          'build_func_name': func, 'variables': list(variables.items())}

        def __getstate__(self):
            return self.state
    RefNetwork.__module__, RefNetwork.__qualname__, RefNetwork.__name__ = 'tfutil', 'Network', 'Network'
    mod.Network = RefNetwork
    sys.modules['tfutil'] = mod
    try:
        path = str(tmp_path / 'network-snapshot-000000.pkl')
        with open(path, 'wb') as f:
            pickle.dump(tuple(RefNetwork(k, f_, params[k], R.CONFIG[f_]) for k, f_ in funcs.items()), f,
                        protocol=pickle.HIGHEST_PROTOCOL)
    finally:
        del sys.modules['tfutil']
    nets = misc.load_pkl(path)
    assert len(nets) == 9 and all(isinstance(n, Network) for n in nets)
    for net, (k, f_) in zip(nets, funcs.items()):
        assert net.name == k and net._build_func_name == f_
        assert list(net.vars.keys()) == list(params[k].keys())
        for vn, want in params[k].items():
            assert np.array_equal(net.get_var(vn), np.asarray(want, np.float32)), (k, vn)
    assert nets[8].lod == 1.5 and nets[2].lod == 0.0
    # save: same class reference, same state keys, the file's own module source carried through
    out = str(tmp_path / 'network-final.pkl')
    misc.save_pkl(nets, out)
    assert 'tfutil' not in sys.modules
    ops = [(op.name, arg) for op, arg, _ in pickletools.genops(open(out, 'rb').read())]
    globals_named = {arg for name, arg in ops if name in ('GLOBAL', 'STACK_GLOBAL', 'SHORT_BINUNICODE', 'BINUNICODE')
                     and isinstance(arg, str)}
    assert 'tfutil' in globals_named and 'Network' in globals_named
    assert not any('texturemixer_b200' in a for a in globals_named)
    again = misc.load_pkl(out)
    for a, b in zip(again, nets):
        st_a, st_b = a.__getstate__(), b.__getstate__()
        assert set(st_a) == {'version', 'name', 'static_kwargs', 'build_module_src', 'build_func_name', 'variables'}
        assert st_a['build_module_src'].startswith('raise SystemExit') and st_a['name'] == st_b['name']
        for (na, va), (nb, vb) in zip(st_a['variables'], st_b['variables']):
            assert na == nb and np.array_equal(va, vb)


def test_training_schedule_matches_reference_class():
    """train.TrainingSchedule vs the reference's own class (run.py:187-226) executed from the reference file by
    tests/golden/make_golden.py -> schedule.npz: lod, resolution, minibatch, lrate, tick at phase boundaries."""
    import types
    from texturemixer_b200.train import TrainingSchedule
    src = open(os.path.join(GOLDEN, 'make_golden.py')).read()
    ns = {}
    exec(src[src.index('SCHEDULES = ['):src.index('def gen_schedule')], ns)
    g = np.load(os.path.join(GOLDEN, 'schedule.npz'))
    for i, (gpus, kw) in enumerate(ns['SCHEDULES']):
        for nimg, want in zip(ns['SCHEDULE_NIMG'], g['sched%d' % i]):
            s = TrainingSchedule(nimg, 7, num_gpus=gpus, **kw)
            got = [s.lod, s.resolution, s.minibatch, s.lrate, s.tick_kimg]
            assert np.array_equal(np.array(got, np.float64), want), (i, nimg, got, want)


def test_app_mattes_match_reference_functions():
    """interp.gkern_* / linkern_for_weight_square / gkern_for_scale_horizontal vs the reference's own functions
    (util_scripts.py:53-62, 104-182) executed by make_golden.py -> mattes_apps.npz.  The hybridization RBF kernel is
    a Python double loop of scalar math in the reference and array math here: equal to the last ulp of exp()."""
    from texturemixer_b200 import interp
    src = open(os.path.join(GOLDEN, 'make_golden.py')).read()
    ns = {}
    exec(src[src.index('APP_MATTE_CASES = dict('):src.index('def gen_app_mattes')], ns)
    g = np.load(os.path.join(GOLDEN, 'mattes_apps.npz'))
    for name, cases in ns['APP_MATTE_CASES'].items():
        for i, args in enumerate(cases):
            got = getattr(interp, name)(*args)
            got = np.stack(got) if isinstance(got, tuple) else np.asarray(got)
            want = g['%s_%d' % (name, i)]
            assert got.shape == want.shape and got.dtype == want.dtype, (name, i)
            if 'gkern_for_weight' in name:
                assert np.allclose(got, want, rtol=1e-14, atol=1e-300), (name, i)
            else:
                assert np.array_equal(got, want), (name, i)


def test_checkpoint_unpickler_refuses_foreign_globals(tmp_path):
    """misc.load_pkl resolves only what a reference checkpoint can name; a reducer naming os.system (or any other
    global) raises instead of being imported and called."""
    import pickle
    from texturemixer_b200 import misc

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ('echo pwned > %s' % (tmp_path / 'pwned'),))
    p = tmp_path / 'evil.pkl'
    with open(p, 'wb') as f:
        pickle.dump((Evil(),), f)
    with pytest.raises(pickle.UnpicklingError):
        misc.load_pkl(str(p))
    assert not (tmp_path / 'pwned').exists()


def test_broadcast_inputs_collapse_to_one_pixel():
    """Network.run uploads a host-tiled global code (np.broadcast_to of [N,C,1,1]) as [N,C,1,1]."""
    from texturemixer_b200.network import _collapse_broadcast
    zg = np.random.RandomState(0).randn(3, 8, 1, 1).astype(np.float32)
    view = np.broadcast_to(zg, (3, 8, 32, 32))
    assert _collapse_broadcast(view).shape == (3, 8, 1, 1) and np.array_equal(_collapse_broadcast(view), zg)
    tiled = np.tile(zg, (1, 1, 4, 4))
    assert _collapse_broadcast(tiled) is tiled and _collapse_broadcast(zg) is zg


def test_fused_dgrad_tape_analysis():
    """backward.gp_fusion_maps: the conv that may prepare an activation's gradient inside its data gradient is the
    activation's EARLIEST consumer (every later consumer has already been differentiated when the reverse pass gets
    there), and only activations produced by conv / FromRGB / pool / window / concat records qualify."""
    from texturemixer_b200.backward import gp_fusion_maps
    from texturemixer_b200.runtime import Act
    a0, a1, a2, a3, p3, m4 = (Act(2, 8, 8, 32) for _ in range(6))
    img = object()
    tape = [
        dict(kind='fromrgb', img=img, y=a0),                          # 0
        dict(kind='conv', x=a0, y=a1, residual=None, up2=False),      # 1  Residual_0
        dict(kind='conv', x=a1, y=a2, residual=a0, up2=False),        # 2  Residual_1: + a0
        dict(kind='conv', x=a2, y=a3, residual=None, up2=False),      # 3
        dict(kind='pool', x=a3, y=p3),                                # 4
        dict(kind='mbstd', x=p3, y=m4, group=4),                      # 5
        dict(kind='outputs', tensors=[]),
    ]
    first_use, producer = gp_fusion_maps(tape)
    assert first_use[id(a0)] == 1          # conv 1 reads it before conv 2 adds it back as the residual
    assert first_use[id(a1)] == 2 and first_use[id(a2)] == 3 and first_use[id(a3)] == 4 and first_use[id(p3)] == 5
    assert producer[id(a0)]['kind'] == 'fromrgb' and producer[id(a1)] is tape[1] and producer[id(a2)] is tape[2]
    assert producer[id(p3)]['kind'] == 'pool'
    assert id(m4) not in producer          # minibatch stddev: not a producer the fused kernel stands in for
    assert id(img) not in first_use        # plain tensors (images) are not activations
