"""ORACLE helper (generation-time only): execute pieces of the UNMODIFIED
reference from /root/reference inside this container.

* `reference_functions(file, names)` pulls plain-numpy functions out of a
  reference file by AST (run.py / util_scripts.py import TF and a dozen absent
  packages at module level, so they cannot be imported whole).
* `reference_networks()` imports the reference's `networks.py` on top of
  `oracle/tfshim` (a torch-backed stand-in for the TF ops it calls).

/root/reference does not exist on the GPU box: nothing here may be used by
`-m gpu` tests, smoke() or bench.py.  Golden fixtures are generated once by
tests/golden/make_golden.py and committed.
"""
import ast
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get('TMX_REFERENCE_ROOT', '/root/reference')
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tfshim')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'networks.py'))


def reference_functions(filename, names, extra_globals=None):
    import numpy as np
    src = open(os.path.join(REFERENCE_ROOT, filename)).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    assert not missing, 'not found in %s: %s' % (filename, missing)
    ns = {'np': np, '__name__': 'reference_' + filename}
    ns.update(extra_globals or {})
    exec(compile(ast.Module(body=wanted, type_ignores=[]), os.path.join(REFERENCE_ROOT, filename), 'exec'), ns)
    return {n: ns[n] for n in names}


def _import_reference(modname):
    for p in (REFERENCE_ROOT, SHIM_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, SHIM_DIR)           # shim 'tensorflow' + 'custom_vgg19' win
    try:
        return importlib.import_module(modname)
    finally:
        sys.path.remove(SHIM_DIR)
        sys.path.remove(REFERENCE_ROOT)


def reference_networks():
    """-> (reference `networks` module, shim `tensorflow` module)."""
    net = _import_reference('networks')
    return net, sys.modules['tensorflow']


def reference_loss():
    """-> (reference `loss` module, reference `networks` module, shim `tensorflow` module).  loss.py imports
    `tfutil` (shim: lerp + autosummary), `config` (the reference's own, pure Python) and `networks`."""
    net = _import_reference('networks')
    los = _import_reference('loss')
    return los, net, sys.modules['tensorflow']


class ReferenceNetwork:
    """The slice of tfutil.Network that loss.py uses (tfutil.py:416-516): `get_output_for` re-runs the reference's
    build function under the network's variable scope (variables are shared by name through the shim's store, like
    tfutil.py:471-474 does for G_fcn over G), plus the static shape attributes."""

    def __init__(self, net_module, tf, scope, func, input_shapes, output_shapes, **static_kwargs):
        self.net_module, self.tf, self.scope, self.func = net_module, tf, scope, func
        self.static_kwargs = dict(static_kwargs)
        self.input_shapes, self.output_shapes = [list(s) for s in input_shapes], [list(s) for s in output_shapes]
        self.input_shape, self.output_shape = self.input_shapes[0], self.output_shapes[0]

    def get_output_for(self, *ins):
        with self.tf.variable_scope(self.scope):
            return getattr(self.net_module, self.func)(*ins, **self.static_kwargs)


class ReferenceOptimizerStub:
    """tfutil.Optimizer.apply_loss_scaling / undo_loss_scaling with use_loss_scaling=False (tfutil.py:378-391):
    both return their argument."""

    @staticmethod
    def apply_loss_scaling(value):
        return value

    @staticmethod
    def undo_loss_scaling(value):
        return value
