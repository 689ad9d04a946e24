"""ORACLE / test infrastructure: a tiny *eager* stand-in for the subset of
TensorFlow 1.12 that `/root/reference/{networks,loss}.py` touch, backed by
torch-CPU, so that the reference's own Python can be executed UNMODIFIED in
this container to generate golden vectors (tests/golden/make_golden.py).

It is not TensorFlow: only op *semantics* are restated here (NCHW
cross-correlation conv, REFLECT pad, VALID avg-pool, ...); everything
structural (layer order, gains, scopes, variable names, tf.cond branch
selection) comes from the reference code that runs on top of it.
Nothing in the product imports this package.

tf.cond builds both branches (true_fn first) like graph-mode TF does - so the
variable creation order is the reference's - and returns the selected one.
"""
import builtins
import contextlib
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

__version__ = '1.12.0-shim'


# ------------------------------------------------------------------ dtypes
class DType:
    def __init__(self, name, tdt):
        self.name, self.torch = name, tdt

    def __repr__(self):
        return 'tf.' + self.name

    def __eq__(self, o):
        return isinstance(o, DType) and o.name == self.name

    def __hash__(self):
        return hash(self.name)


float16 = DType('float16', torch.float16)
float32 = DType('float32', torch.float32)
float64 = DType('float64', torch.float64)
int32 = DType('int32', torch.int32)
int64 = DType('int64', torch.int64)
_DT = {d.name: d for d in (float16, float32, float64, int32, int64)}
_FROM_TORCH = {d.torch: d for d in _DT.values()}

# Compute dtype that `float32` maps to (flip to torch.float64 for budget runs).
COMPUTE = {'float32': torch.float32}


def as_dtype(x):
    if isinstance(x, DType):
        return x
    return _DT[np.dtype(x).name if not isinstance(x, str) else x]


def _tdt(d):
    d = as_dtype(d)
    return COMPUTE.get(d.name, d.torch)


# ------------------------------------------------------------------ shapes
class Dimension:
    def __init__(self, v):
        self.value = None if v is None else int(v)

    def __int__(self):
        return self.value

    __index__ = __int__

    def __mul__(self, o):
        return Dimension(self.value * int(o))

    __rmul__ = __mul__

    def __floordiv__(self, o):
        return Dimension(self.value // int(o))

    def __sub__(self, o):
        return Dimension(self.value - int(o))

    def __add__(self, o):
        return Dimension(self.value + int(o))

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return 'Dimension(%s)' % self.value


class TensorShape:
    def __init__(self, dims):
        self.dims = [d if isinstance(d, Dimension) else Dimension(d) for d in dims]

    def __getitem__(self, i):
        if isinstance(i, builtins.slice):
            return TensorShape(self.dims[i])
        return self.dims[i]

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(self.dims)

    def as_list(self):
        return [d.value for d in self.dims]

    def __repr__(self):
        return 'TensorShape(%s)' % self.as_list()


def _ival(v):
    if isinstance(v, Dimension):
        return v.value
    if isinstance(v, Tensor):
        return int(v.t.item())
    if isinstance(v, (np.integer, np.floating)):
        return int(v)
    return v


def _ilist(seq):
    return [_ival(v) for v in seq]


# ------------------------------------------------------------------ tensors
class Tensor:
    __array_priority__ = 1000

    def __init__(self, t, name=None):
        self.t = t
        self.name = name

    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        for have, want in zip(self.t.shape, shape):
            assert want is None or have == _ival(want), (tuple(self.t.shape), shape)
        assert len(shape) == self.t.dim()

    @property
    def dtype(self):
        for name, tdt in COMPUTE.items():
            if self.t.dtype == tdt:
                return _DT[name]
        return _FROM_TORCH[self.t.dtype]

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = tuple(builtins.slice(_ival(i.start), _ival(i.stop), _ival(i.step)) if isinstance(i, builtins.slice) else _ival(i)
                    for i in idx)
        return Tensor(self.t[idx])

    def _bin(self, o, fn, rev=False):
        o = _raw(o, like=self.t)
        return Tensor(fn(o, self.t) if rev else fn(self.t, o))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __pow__(self, o): return Tensor(self.t ** o)
    def __neg__(self): return Tensor(-self.t)
    def __lt__(self, o): return Tensor(self.t < _raw(o, like=self.t))
    def __gt__(self, o): return Tensor(self.t > _raw(o, like=self.t))
    def __le__(self, o): return Tensor(self.t <= _raw(o, like=self.t))
    def __ge__(self, o): return Tensor(self.t >= _raw(o, like=self.t))
    def __bool__(self): return bool(self.t.item())
    def __repr__(self): return 'shim.Tensor(%s, %s)' % (tuple(self.t.shape), self.t.dtype)


class Variable(Tensor):
    def __init__(self, t, name, trainable):
        super().__init__(t, name)
        self.trainable = trainable


def _raw(x, like=None):
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, Dimension):
        x = x.value
    if isinstance(x, torch.Tensor):
        return x
    if like is not None:
        dt = like.dtype if like.dtype.is_floating_point or not isinstance(x, float) else torch.float32
        return torch.as_tensor(np.asarray(x)).to(dt)
    a = np.asarray(x)
    t = torch.as_tensor(a)
    if a.dtype == np.float32:
        t = t.to(COMPUTE['float32'])
    return t


def convert_to_tensor(x, dtype=None):
    t = _raw(x)
    if dtype is not None:
        t = t.to(_tdt(dtype))
    return Tensor(t)


# ------------------------------------------------------------------ variables / scopes
class _Store:
    def __init__(self):
        self.reset()

    def reset(self, values=None, rng=None):
        self.vars = OrderedDict()          # full name -> Variable (creation order)
        self.values = dict(values or {})   # injected values by full name
        self.scope = []
        self.rng = rng or np.random.RandomState(0)
        self.requires_grad = False         # goldens of loss gradients: trainable variables become autograd leaves


STORE = _Store()


def reset_default_graph(values=None, rng=None, requires_grad=False):
    STORE.reset(values, rng)
    STORE.requires_grad = bool(requires_grad)


@contextlib.contextmanager
def variable_scope(name, reuse=None, **_):
    STORE.scope.append(name)
    try:
        yield
    finally:
        STORE.scope.pop()


@contextlib.contextmanager
def name_scope(name=None, *a, **k):
    yield


@contextlib.contextmanager
def control_dependencies(deps):
    yield


@contextlib.contextmanager
def device(name):
    yield


AUTO_REUSE = object()


class _Init:
    def __init__(self, kind, mean=0.0, stddev=1.0):
        self.kind, self.mean, self.stddev = kind, mean, stddev


class initializers:
    @staticmethod
    def random_normal(mean=0.0, stddev=1.0):
        return _Init('normal', mean, stddev)

    @staticmethod
    def zeros():
        return _Init('zeros')


def get_variable(name, shape=None, initializer=None, trainable=True, dtype=None):
    full = '/'.join(STORE.scope + [name])
    if full in STORE.vars:
        return STORE.vars[full]
    if shape is not None:
        shape = _ilist(shape)
    if full in STORE.values:
        val = np.asarray(STORE.values[full], dtype=np.float32)
        if shape is not None:
            assert list(val.shape) == list(shape), (full, val.shape, shape)
    elif isinstance(initializer, _Init):
        if initializer.kind == 'zeros':
            val = np.zeros(shape, np.float32)
        else:
            val = (initializer.mean + initializer.stddev * STORE.rng.randn(*shape)).astype(np.float32)
    else:
        val = np.asarray(initializer, dtype=np.float32)
    t = torch.as_tensor(val).to(COMPUTE['float32']).clone()
    if STORE.requires_grad and trainable:
        t.requires_grad_(True)
    v = Variable(t, full, trainable)
    STORE.vars[full] = v
    return v


# ------------------------------------------------------------------ ops
def constant(value, dtype=None, shape=None, name=None, verify_shape=False):
    a = np.asarray(value, dtype=np.float32 if dtype is None and not isinstance(value, (int, np.integer)) else None)
    t = torch.as_tensor(a)
    if dtype is not None:
        t = t.to(_tdt(dtype))
    elif a.dtype == np.float32:
        t = t.to(COMPUTE['float32'])
    if shape is not None:
        shape = _ilist(shape)
        t = t.reshape(shape) if t.numel() == int(np.prod(shape)) else t.expand(shape).clone()
    return Tensor(t, name)


def cast(x, dtype):
    return Tensor(_raw(x).to(_tdt(dtype)))


def identity(x, name=None):
    return Tensor(_raw(x), name)


def shape(x):
    return [int(d) for d in _raw(x).shape]


def reshape(x, shp):
    return Tensor(_raw(x).reshape(_ilist(shp)))


def tile(x, multiples):
    return Tensor(_raw(x).repeat(*_ilist(multiples)))


def concat(values, axis):
    return Tensor(torch.cat([_raw(v) for v in values], dim=axis))


def split(value, num_or_size_splits, axis=0):
    return [Tensor(t) for t in torch.chunk(_raw(value), int(num_or_size_splits), dim=axis)]


def expand_dims(x, axis):
    return Tensor(_raw(x).unsqueeze(axis))


def transpose(x, perm):
    return Tensor(_raw(x).permute(*perm))


def maximum(a, b):
    a = _raw(a)
    return Tensor(torch.maximum(a, _raw(b, like=a)))


def minimum(a, b):
    if not isinstance(a, (Tensor, torch.Tensor)) and not isinstance(b, (Tensor, torch.Tensor)):
        return min(_ival(a), _ival(b))
    a = _raw(a)
    return Tensor(torch.minimum(a, _raw(b, like=a)))


def clip_by_value(x, lo, hi):
    return Tensor(torch.clamp(_raw(x), lo, hi))


def square(x): return Tensor(_raw(x) ** 2)
def sqrt(x): return Tensor(torch.sqrt(_raw(x)))
def rsqrt(x): return Tensor(torch.rsqrt(_raw(x)))
def exp(x): return Tensor(torch.exp(_raw(x)))
def abs(x): return Tensor(torch.abs(_raw(x)))  # noqa: A001
def floor(x): return Tensor(torch.floor(_raw(x)))
def zeros(shp, dtype=float32): return Tensor(torch.zeros(_ilist(shp), dtype=_tdt(dtype)))
def add_n(xs): return Tensor(sum(_raw(x) for x in xs))


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (axis,)


def reduce_mean(x, axis=None, keepdims=False):
    x = _raw(x)
    return Tensor(x.mean() if axis is None else x.mean(dim=_axes(axis), keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False):
    x = _raw(x)
    return Tensor(x.sum() if axis is None else x.sum(dim=_axes(axis), keepdim=keepdims))


def matmul(a, b, adjoint_a=False):
    a = _raw(a)
    if adjoint_a:
        a = a.transpose(-1, -2)
    return Tensor(a @ _raw(b))


def pad(x, paddings, mode='CONSTANT', constant_values=0):
    x = _raw(x)
    flat = []
    for lo, hi in reversed([list(p) for p in paddings]):
        flat += [_ival(lo), _ival(hi)]
    while len(flat) > 2 and flat[-1] == 0 and flat[-2] == 0:
        flat = flat[:-2]
    if mode.upper() == 'REFLECT':
        return Tensor(F.pad(x, flat, mode='reflect'))
    return Tensor(F.pad(x, flat, mode='constant', value=constant_values))


def reverse(x, axis):
    return Tensor(torch.flip(_raw(x), dims=list(axis)))


def slice(x, begin, size):  # noqa: A001
    x = _raw(x)
    begin = [_ival(b) for b in (_raw(begin).tolist() if isinstance(begin, Tensor) else begin)]
    size = _ilist(size)
    idx = tuple(builtins.slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return Tensor(x[idx])


def cond(pred, true_fn, false_fn):
    t = true_fn()      # graph-mode TF builds both branches, true_fn first
    f = false_fn()
    return t if bool(_raw(pred).item() if isinstance(pred, (Tensor, torch.Tensor)) else pred) else f


def where(c, a, b):
    return Tensor(torch.where(_raw(c), _raw(a), _raw(b)))


def gradients(ys, xs):
    ys = ys if isinstance(ys, (list, tuple)) else [ys]
    y = sum(_raw(v).sum() for v in ys)
    gs = torch.autograd.grad(y, [_raw(x) for x in xs], create_graph=True, allow_unused=True)
    return [None if g is None else Tensor(g) for g in gs]


# Random ops draw from a replaceable hook so that goldens can pin the values.
RANDOM = {'uniform': None, 'normal': None, 'uniform_int': None}


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=float32):  # noqa: A002
    shp = _ilist(shape)
    d = as_dtype(dtype)
    if d in (int32, int64):
        fn = RANDOM['uniform_int'] or (lambda s, lo, hi: STORE.rng.randint(lo, hi, size=s))
        return Tensor(torch.as_tensor(np.asarray(fn(shp, minval, maxval))).to(d.torch))
    fn = RANDOM['uniform'] or (lambda s, lo, hi: STORE.rng.uniform(lo, hi, size=s))
    return Tensor(torch.as_tensor(np.asarray(fn(shp, minval, maxval), dtype=np.float32)).to(_tdt(dtype)))


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32):  # noqa: A002
    shp = _ilist(shape)
    fn = RANDOM['normal'] or (lambda s: STORE.rng.randn(*s))
    return Tensor((mean + stddev * torch.as_tensor(np.asarray(fn(shp), dtype=np.float32))).to(_tdt(dtype)))


class nn:
    @staticmethod
    def conv2d(x, w=None, strides=(1, 1, 1, 1), padding='VALID', data_format='NHWC', filter=None, name=None):  # noqa: A002
        w = w if w is not None else filter
        assert data_format in ('NCHW', 'NHWC') and padding in ('VALID', 'SAME')
        if data_format == 'NHWC':       # tensorflow_vgg's layers: same op on the transposed tensor
            s = _ilist(strides)
            y = nn.conv2d(Tensor(_raw(x).permute(0, 3, 1, 2)), w, (1, 1, s[1], s[2]), padding, 'NCHW')
            return Tensor(y.t.permute(0, 2, 3, 1))
        s = _ilist(strides)
        xr, wr = _raw(x), _raw(w)
        if padding == 'SAME':
            # TF SAME: out = ceil(in / stride), pad_total = max((out - 1) * stride + k - in, 0), the odd pixel goes
            # AFTER (networks.py:148 conv2d_downscale2d: k = 4, stride 2 -> one zero pixel on every side)
            pads = []
            for dim, k, st in ((3, wr.shape[1], s[3]), (2, wr.shape[0], s[2])):
                n_in = xr.shape[dim]
                total = max((-(-n_in // st) - 1) * st + k - n_in, 0)
                pads += [total // 2, total - total // 2]
            xr = F.pad(xr, pads)
        # TF conv2d = cross-correlation with HWIO filters; torch conv2d = cross-correlation with OIHW.
        return Tensor(F.conv2d(xr, wr.permute(3, 2, 0, 1), stride=(s[2], s[3])))

    @staticmethod
    def avg_pool(x, ksize, strides, padding='VALID', data_format='NHWC', name=None):
        k, s = _ilist(ksize), _ilist(strides)
        if data_format == 'NHWC':
            xr = _raw(x)
            # SAME == VALID when the window tiles the map exactly (VGG on power-of-two crops)
            assert padding == 'VALID' or (xr.shape[1] % s[1] == 0 and xr.shape[2] % s[2] == 0 and k[1:3] == s[1:3])
            return Tensor(F.avg_pool2d(xr.permute(0, 3, 1, 2), (k[1], k[2]), (s[1], s[2])).permute(0, 2, 3, 1))
        assert data_format == 'NCHW' and padding == 'VALID'
        return Tensor(F.avg_pool2d(_raw(x), (k[2], k[3]), (s[2], s[3])))

    @staticmethod
    def bias_add(x, b, data_format='NHWC'):
        xr, br = _raw(x), _raw(b)
        return Tensor(xr + (br if data_format == 'NHWC' else br.reshape(1, -1, 1, 1)))

    @staticmethod
    def relu(x):
        return Tensor(torch.relu(_raw(x)))

    @staticmethod
    def tanh(x, name=None):
        return Tensor(torch.tanh(_raw(x)), name)

    @staticmethod
    def conv2d_transpose(x, w, output_shape, strides=(1, 1, 1, 1), padding='SAME', data_format='NHWC'):
        """Gradient of nn.conv2d(SAME) w.r.t. its input (networks.py:101 upscale2d_conv2d): filter [h, w, out, in].
        For stride 2 and the even 4x4 kernel SAME pads one pixel on every side of the forward conv's input, so the
        transpose crops one pixel per side: torch padding = 1."""
        assert data_format == 'NCHW' and padding == 'SAME'
        s = _ilist(strides)
        xr, wr = _raw(x), _raw(w)
        k = wr.shape[0]
        assert s[2] == s[3] == 2 and k == wr.shape[1] and (k - 2) % 2 == 0, 'shim: stride-2 SAME with an even kernel'
        out = F.conv_transpose2d(xr, wr.permute(3, 2, 0, 1), stride=2, padding=(k - 2) // 2)
        want = _ilist(output_shape)
        assert list(out.shape[1:]) == [int(v) for v in want[1:]], (tuple(out.shape), want)
        return Tensor(out)


def trainable_variables():
    return [v for v in STORE.vars.values() if v.trainable]
