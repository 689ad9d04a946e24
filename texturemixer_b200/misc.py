"""Checkpoint files of the reference (`misc.load_pkl` / `misc.save_pkl`, misc.py:27-33; written at run.py:583,589 as
the 9-tuple (E_zg, E_zl, G, D_rec, D_interp, D_blend, Es_zg, Es_zl, Gs) and read back at run.py:259 and
util_scripts.py:354,594,1138,1305).

A reference pickle stores every network as an instance of `tfutil.Network` (or `network.Network` for the legacy
Theano files, legacy.py:19-26) whose state is the version-2 dict of tfutil.py:543-550.  `load_pkl` resolves those
class names to `texturemixer_b200.network.Network`, whose `__setstate__` rebuilds the variables by NAME and never
exec's the module source stored in the file.  `save_pkl` writes the same class reference and
state layout, and carries a loaded network's original `build_module_src` through unchanged, so a file that came
from the reference goes back to it intact (a network created here stores this package's own build source, which
the reference's TensorFlow code cannot execute - stated limitation).

`ReferenceUnpickler.find_class` is a WHITELIST: the Network stubs, numpy array reconstruction, OrderedDict and the
plain-object reconstructor are everything a reference checkpoint names; any other global (os.system, builtins.eval,
...) raises UnpicklingError instead of being imported."""
import copyreg
import pickle
import sys
import types

from .network import Network

_REFERENCE_CLASSES = {('tfutil', 'Network'), ('network', 'Network')}


# every global a checkpoint written by run.py:583 / misc.py:31-33 (or by save_pkl below) can name
_ALLOWED_GLOBALS = {
    ('numpy.core.multiarray', '_reconstruct'), ('numpy._core.multiarray', '_reconstruct'),
    ('numpy.core.multiarray', 'scalar'), ('numpy._core.multiarray', 'scalar'),
    ('numpy', 'ndarray'), ('numpy', 'dtype'),
    ('numpy.core.numeric', '_frombuffer'), ('numpy._core.numeric', '_frombuffer'),      # protocol-5 array payloads
    ('collections', 'OrderedDict'), ('copyreg', '_reconstructor'), ('copy_reg', '_reconstructor'),
    ('builtins', 'object'), ('__builtin__', 'object'),
}


class ReferenceUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _REFERENCE_CLASSES:
            return Network
        if (module, name) in _ALLOWED_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError('checkpoint names the global %s.%s, which a TextureMixer network pickle has no '
                                     'business naming - refusing to import it' % (module, name))


def load_pkl(filename):
    """misc.py:27-29."""
    with open(filename, 'rb') as f:
        return ReferenceUnpickler(f, encoding='latin1').load()


def save_pkl(obj, filename):
    """misc.py:31-33: HIGHEST_PROTOCOL pickle in which every Network is recorded as `tfutil.Network`."""
    stub_mod = types.ModuleType('tfutil')
    stub_cls = type('Network', (), {})
    stub_cls.__module__ = 'tfutil'
    stub_cls.__qualname__ = 'Network'
    stub_mod.Network = stub_cls
    prev = sys.modules.get('tfutil')
    sys.modules['tfutil'] = stub_mod          # pickle verifies that the recorded global resolves while dumping

    class _Pickler(pickle.Pickler):
        def reducer_override(self, o):
            if isinstance(o, Network):
                # object.__new__(tfutil.Network) followed by __setstate__(state): what unpickling the reference's
                # own files does
                return copyreg._reconstructor, (stub_cls, object, None), o.__getstate__()
            return NotImplemented

    try:
        with open(filename, 'wb') as f:
            _Pickler(f, protocol=pickle.HIGHEST_PROTOCOL).dump(obj)
    finally:
        if prev is None:
            del sys.modules['tfutil']
        else:
            sys.modules['tfutil'] = prev
