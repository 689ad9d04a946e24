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
