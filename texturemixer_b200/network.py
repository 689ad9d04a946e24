"""`Network`: the reference's plugin boundary (tfutil.py:416-750) over device
tensors.  Build functions are selected by dotted name (config.py:77-82 ->
tfutil.py:212-236), called once in *template* mode to discover inputs,
outputs and variables (tfutil.py:458-494) and then in *run* mode for every
evaluation (tfutil.py:505-516).  Variables keep the reference's names and
HWIO/[in,out] shapes, so `__getstate__` produces the reference's version-2
pickle dict (tfutil.py:543-550)."""
import contextlib
import gc
import threading
import os
import importlib
import inspect
from collections import OrderedDict

import numpy as np
import torch

from .runtime import Runtime


# ---------------------------------------------------------------------- name resolution (tfutil.py:212-236)
def import_module(module_or_obj_name):
    parts = module_or_obj_name.split('.')
    parts[0] = {'np': 'numpy', 'tf': 'tensorflow'}.get(parts[0], parts[0])
    for i in range(len(parts), 0, -1):
        name = '.'.join(parts[:i])
        # 'networks.G_res' / 'loss.EG_wgan' resolve to OUR modules: the prefixed candidate is tried FIRST for them, so
        # that a bare `networks` on sys.path (e.g. the reference's TensorFlow file) is never imported, let alone run
        ours_first = parts[0] in ('networks', 'loss', 'network', 'interp', 'train', 'optim', 'misc')
        cands = ('texturemixer_b200.' + name,) if ours_first else (name, 'texturemixer_b200.' + name)
        for cand in cands:
            try:
                module = importlib.import_module(cand)
            except ImportError:
                continue
            return module, '.'.join(parts[i:])
    raise ImportError(module_or_obj_name)


def find_obj_in_module(module, relative_obj_name):
    obj = module
    for part in relative_obj_name.split('.'):
        obj = getattr(obj, part)
    return obj


def import_obj(obj_name):
    module, relative = import_module(obj_name)
    return find_obj_in_module(module, relative)


def call_func_by_name(*args, func=None, **kwargs):
    assert func is not None
    return import_obj(func)(*args, **kwargs)


# ---------------------------------------------------------------------- variables
class Variable:
    """A named parameter living in the owning network's flat device buffer."""
    __slots__ = ('name', 'shape', 'trainable', 'init', 'value', 'size')

    def __init__(self, name, shape, trainable, init):
        self.name, self.shape, self.trainable, self.init = name, tuple(int(s) for s in shape), trainable, init
        self.value = None            # torch view, set by Network._allocate
        self.size = int(np.prod(self.shape)) if self.shape else 1


class T:
    """Tensor handle seen by build functions: a static NCHW shape (None batch
    in template mode) plus, in run mode, the device payload - either an
    external NCHW torch tensor (`nchw`) or an internal activation (`act`)."""
    __slots__ = ('shape', 'ctx', 'nchw', 'act', 'name', 'rgb', 'tanh_done')

    def __init__(self, shape, ctx, nchw=None, act=None, name=None):
        self.shape, self.ctx, self.nchw, self.act, self.name = list(shape), ctx, nchw, act, name
        if ctx is not None and ctx.mode == 'run':
            ctx.handles.append(self)
        self.rgb = None      # images already produced by a fused ToRGB epilogue: (scope, tanh, NCHW tensor)
        self.tanh_done = False   # image head: tanh already applied by the ToRGB kernel (lod == 0)

    def set_shape(self, shape):
        """tf.Tensor.set_shape: fixes the template shape / checks the fed one
        (the reference's implicit shape assert, networks.py:221,325,420-421)."""
        shape = list(shape)
        if self.ctx.mode == 'template':
            self.shape = shape
        else:
            have = self.shape
            if len(have) != len(shape) or any(w is not None and h != w for h, w in zip(have, shape)):
                raise ValueError('%s: input shape %s incompatible with %s' % (self.ctx.net.name, have, shape))


class BuildContext:
    def __init__(self, net, mode, tape=None):
        self.net, self.mode = net, mode
        self.scopes = []
        self.rt = net.rt if mode == 'run' else None
        self.tape = tape        # list of layer records when the forward is run for training (backward.py)
        self.pixelnorm = None   # epsilon of pixel_norm after every activated conv (set by the build function)
        self.handles = []       # every run-mode T, so that release() can drop their device payloads

    def release(self):
        """The recursive closures of the build functions (grow -> grow) are reference CYCLES that capture this
        context and their T handles; left alone they keep every activation and the whole tape of the evaluation
        alive until the cyclic garbage collector happens to run (measured: 16 GB per train step).  Dropping the
        payload references here returns device memory by reference counting as soon as the caller lets go."""
        for t in self.handles:
            t.nchw = t.act = t.rgb = None
        self.handles = []
        self.tape = None

    @contextlib.contextmanager
    def variable_scope(self, name):
        self.scopes.append(name)
        try:
            yield
        finally:
            self.scopes.pop()

    def get_variable(self, name, shape=(), init='normal', trainable=True):
        full = '/'.join(self.scopes + [name])
        v = self.net.vars.get(full)
        if v is None:
            if self.mode != 'template':
                raise KeyError('%s: variable %s does not exist' % (self.net.name, full))
            v = Variable(full, shape, trainable, init)
            self.net.vars[full] = v
        elif tuple(shape) != v.shape:
            raise ValueError('%s: variable %s has shape %s, requested %s' % (self.net.name, full, v.shape, tuple(shape)))
        return v

    def cond(self, pred, true_fn, false_fn):
        """tf.cond as the reference's graph mode sees it: the template builds BOTH
        branches (true_fn first -> same variable creation order, incl. the unused
        lod heads); run mode evaluates only the selected branch."""
        if self.mode == 'template':
            t = true_fn()
            f = false_fn()
            return t if pred else f
        return true_fn() if pred else false_fn()


_CAPTURE_LOCK = threading.Lock()


# ---------------------------------------------------------------------- Network
class Network:
    def __init__(self, name=None, func=None, reuse=False, share_vars_with=None, device=None, seed=None,
                 **static_kwargs):
        """name/func/reuse/**static_kwargs as tfutil.Network.__init__ (tfutil.py:417-435).
        `reuse=True` + `share_vars_with=<Network>` gives the reference's second
        view over the SAME scope (G_fcn over G, run.py:273)."""
        self._init_fields()
        self.name = name
        self.static_kwargs = dict(static_kwargs)
        module, self._build_func_name = import_module(func)
        try:
            self._build_module_src = inspect.getsource(module)
        except (OSError, TypeError):
            self._build_module_src = ''
        self._build_func = find_obj_in_module(module, self._build_func_name)
        self._device = device
        self._init_graph(share_vars_with if reuse else None)
        if not (reuse and share_vars_with is not None):
            self.reset_vars(seed)

    def _init_fields(self):
        self.name = None
        self.scope = None
        self.static_kwargs = dict()
        self.num_inputs = 0
        self.num_outputs = 0
        self.input_shapes = [[]]
        self.output_shapes = [[]]
        self.input_shape = []
        self.output_shape = []
        self.input_names = []
        self.output_names = []
        self.vars = OrderedDict()
        self.trainables = OrderedDict()
        self._build_func = None
        self._build_func_name = None
        self._build_module_src = None
        self._flat = None              # one fp32 device buffer holding every variable
        self._version = 0              # bumped whenever variable values change
        self._prepared = {}            # per-variable tensor-core weight planes (w_hi, w_lo, version)
        self._grad_slots = {}          # name -> (offset, size, shape) inside flat / flat gradient buffers
        self._shared_owner = None      # the network whose variables this view shares (reuse=True), if any
        self._lod_host = 0.0
        self._rt = None
        self._device = None
        self._staging = {}
        self._copy_stream = None
        self._graphs = {}
        self._replicas = {}            # device index -> [copy of this network there, variable version] (run(num_gpus=k))

    @property
    def rt(self):
        if self._rt is None:
            if self._device is not None and str(self._device) == 'cpu':
                raise RuntimeError('%s was created with device="cpu" (variables only); evaluation needs the CUDA '
                                   'runtime - there is no CPU path' % self.name)
            self._rt = Runtime.get(self._device)
        return self._rt

    @property
    def lod(self):
        return float(self._owner()._lod_host)

    def _init_graph(self, share_with=None):
        self.input_names = []
        for param in inspect.signature(self._build_func).parameters.values():
            if param.kind == param.POSITIONAL_OR_KEYWORD and param.default is param.empty:
                self.input_names.append(param.name)
        self.num_inputs = len(self.input_names)
        assert self.num_inputs >= 1
        if self.name is None:
            self.name = self._build_func_name
        self.scope = self.name.replace('/', '_')
        self._lod_host = 0.0

        ctx = BuildContext(self, 'template')
        templates = [T([None], ctx, name=n) for n in self.input_names]
        out = self._build_func(*templates, is_template_graph=True, **self.static_kwargs)
        outs = [out] if isinstance(out, T) else list(out)
        self.output_names = [t.name for t in outs]
        self.num_outputs = len(outs)
        self.input_shapes = [list(t.shape) for t in templates]
        self.output_shapes = [list(t.shape) for t in outs]
        self.input_shape = self.input_shapes[0]
        self.output_shape = self.output_shapes[0]
        template_vars = self.vars
        if share_with is not None:
            for k, v in template_vars.items():
                if k not in share_with.vars or share_with.vars[k].shape != v.shape:
                    raise ValueError('reuse: variable %s missing or mis-shaped in %s' % (k, share_with.name))
            self.vars = share_with.vars
            self._shared_owner = share_with
        else:
            self._shared_owner = None
            self._allocate()
        self.trainables = OrderedDict((k, v) for k, v in self.vars.items() if v.trainable)

    def _var_device(self):
        # device='cpu' is allowed for variable bookkeeping only (host-logic tests,
        # checkpoint conversion); evaluation always requires the CUDA runtime.
        if self._device is not None and str(self._device) == 'cpu':
            return torch.device('cpu')
        return self.rt.device

    _ALIGN = 64   # floats: every variable starts on a 256-byte boundary (vector loads, TMA)

    def _allocate(self):
        """One flat fp32 device buffer for all variables (what the fused optimizer
        and the gradient all-reduce walk); padding between variables stays zero."""
        off = 0
        offsets = []
        for v in self.vars.values():
            offsets.append(off)
            off += (v.size + self._ALIGN - 1) // self._ALIGN * self._ALIGN
        self._flat = torch.zeros(off, dtype=torch.float32, device=self._var_device())
        for v, o in zip(self.vars.values(), offsets):
            v.value = self._flat[o:o + v.size].view(v.shape)

    # -------------------------------------------------------------- variables
    @property
    def flat(self):
        """The one fp32 device buffer holding every variable (256-B aligned slots, zero padding)."""
        return self._owner()._flat

    def mark_variables_changed(self):
        """Call after writing `flat` in place (optimizer step): drops the cached tensor-core weight planes."""
        self._touch()
        if 'lod' in self.vars and self.flat.is_cuda is False:
            self._lod_host = float(self.vars['lod'].value)

    def var_offset(self, name):
        """Element offset of variable `name` inside `flat` (same layout for the flat gradient buffer)."""
        v = self.vars[name].value
        return (v.data_ptr() - self.flat.data_ptr()) // 4

    def grad_view(self, flat_grad, name):
        """The slot of variable `name` inside a flat gradient buffer laid out like `flat`."""
        o = self._owner()
        ent = o._grad_slots.get(name)
        if ent is None:
            v = self.vars[name]
            ent = o._grad_slots[name] = (self.var_offset(name), v.size, v.shape)
        off, size, shape = ent
        return flat_grad[off:off + size].view(shape)

    def _owner(self):
        return self._shared_owner._owner() if self._shared_owner is not None else self

    def _touch(self):
        o = self._owner()
        o._version += 1

    def reset_vars(self, seed=None):
        """Run the initialisers (tfutil.py:497-498): weights ~N(0,1) (networks.py:31),
        biases 0 (networks.py:62), lod 0.  TF's own Philox stream is not
        reproducible here; numpy RandomState(seed) is used instead."""
        rng = np.random.RandomState(seed if seed is not None else np.random.randint(1 << 31))
        for v in self.vars.values():
            if v.init == 'normal':
                host = rng.randn(v.size).astype(np.float32)
            elif v.init == 'zeros':
                host = np.zeros(v.size, np.float32)
            else:
                host = np.full(v.size, np.float32(v.init))
            v.value.copy_(torch.from_numpy(host).reshape(v.shape))
        self._lod_host = float(self.vars['lod'].init) if 'lod' in self.vars and self.vars['lod'].init not in ('normal', 'zeros') else 0.0
        self._touch()

    def find_var(self, var_or_localname):
        return self.vars[var_or_localname] if isinstance(var_or_localname, str) else var_or_localname

    def get_var(self, var_or_localname):
        return self.find_var(var_or_localname).value.detach().cpu().numpy().copy()

    def set_var(self, var_or_localname, new_value):
        v = self.find_var(var_or_localname)
        arr = np.asarray(new_value, dtype=np.float32).reshape(v.shape)
        v.value.copy_(torch.from_numpy(np.ascontiguousarray(arr)).reshape(v.shape))
        if v.name == 'lod':
            self._owner()._lod_host = float(arr)
            self._lod_host = float(arr)
        self._touch()

    def set_lod(self, value):
        """The `tf.assign(net.find_var('lod'), lod_in)` of run.py:310: level of detail of the next evaluations."""
        value = float(np.float32(value))
        o = self._owner()
        if 'lod' in self.vars and (o._lod_host != value or self._lod_host != value):
            self.vars['lod'].value.fill_(value)
            o._lod_host = value
            self._lod_host = value
            self._touch()

    def set_vars(self, name_to_value):
        for k, val in name_to_value.items():
            self.set_var(k, val)

    def copy_vars_from(self, src_net):
        for name in self.vars.keys():
            self.vars[name].value.copy_(src_net.vars[name].value)
        self._lod_host = src_net._owner()._lod_host
        self._touch()

    def copy_trainables_from(self, src_net):
        for name in self.trainables.keys():
            self.vars[name].value.copy_(src_net.vars[name].value)
        self._touch()

    def clone(self, name=None):
        """tfutil.py:579-589: same build function, own copy of the variables."""
        net = object.__new__(Network)
        net._init_fields()
        net.name = name if name is not None else self.name
        net.static_kwargs = dict(self.static_kwargs)
        net._build_module_src = self._build_module_src
        net._build_func_name = self._build_func_name
        net._build_func = self._build_func
        net._device = self._device
        net._init_graph()
        net.copy_vars_from(self)
        return net

    def cached(self, key, make):
        """Per-weight-version cache for derived tensors (e.g. transposed planes for the data gradient)."""
        o = self._owner()
        ent = o._prepared.get(key)
        if ent is None or ent[1] != o._version:
            ent = (make(), o._version)
            o._prepared[key] = ent
        return ent[0]

    def prepared_weights(self, var, wscale, k, cin, cout, up2_phase=False, cin_pad=None, xmerge=False):
        """Cached bf16 hi/lo planes of a conv weight for the tensor-core kernel
        (sub-pixel planes when the conv reads through upscale2d); recomputed
        when any variable of the owning network changed."""
        o = self._owner()
        key = (var.name, float(wscale), bool(up2_phase), cin_pad, bool(xmerge))
        ent = o._prepared.get(key)
        if ent is None or ent[2] != o._version:
            if xmerge:
                hi, lo = self.rt.prepare_weights_xmerge(var.value, wscale, cout)
            else:
                hi, lo = self.rt.prepare_weights(var.value, wscale, k, cin, cout, up2_phase=up2_phase, cin_pad=cin_pad)
            ent = (hi, lo, o._version)
            o._prepared[key] = ent
        return ent[0], ent[1]

    # -------------------------------------------------------------- evaluation
    def get_output_for(self, *in_expr, return_as_list=False, tape=None, **dynamic_kwargs):
        """tfutil.py:505-516 on device tensors: NCHW float32 CUDA tensors in,
        NCHW float32 CUDA tensors out (same order/names as the reference).
        `tape`: a list that receives one record per layer (saved activations) so that
        texturemixer_b200.backward.backward() can differentiate this evaluation - the
        counterpart of tf.gradients over the graph built here (tfutil.py:299)."""
        assert len(in_expr) == self.num_inputs
        all_kwargs = dict(self.static_kwargs)
        all_kwargs.update(dynamic_kwargs)
        ctx = BuildContext(self, 'run', tape=tape)
        ins = []
        for x, name in zip(in_expr, self.input_names):
            if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
                raise TypeError('%s.get_output_for: %s must be a float32 CUDA tensor' % (self.name, name))
            ins.append(T(list(x.shape), ctx, nchw=x.contiguous(), name=name))
        out = self._build_func(*ins, **all_kwargs)
        outs = [out] if isinstance(out, T) else list(out)
        res = [t.nchw for t in outs]
        ctx.release()
        if tape is not None:
            tape.append(dict(kind='outputs', tensors=res))
        if return_as_list:
            return res
        return res[0] if len(res) == 1 else tuple(res)

    def run(self, *in_arrays, return_as_list=False, print_progress=False, minibatch_size=None, num_gpus=1,
            out_mul=1.0, out_add=0.0, out_shrink=1, out_dtype=None, **dynamic_kwargs):
        """tfutil.py:624-680: NumPy in, NumPy out, minibatched.  Host buffers are
        staged through pinned memory.  num_gpus = k > 1 (tfutil.py:644-661): every minibatch is split into k
        contiguous parts which run concurrently on k devices of this process (`_run_multi_gpu`); training jobs use one
        rank per GPU instead (texturemixer_b200.parallel)."""
        assert len(in_arrays) == self.num_inputs
        if num_gpus != 1:
            return self._run_multi_gpu(in_arrays, return_as_list, print_progress, minibatch_size, int(num_gpus),
                                       dict(out_mul=out_mul, out_add=out_add, out_shrink=out_shrink,
                                            out_dtype=out_dtype), dynamic_kwargs)
        num_items = in_arrays[0].shape[0]
        if minibatch_size is None:
            minibatch_size = num_items
        # a global code tiled on the host (np.tile / np.broadcast_to of a [N,C,1,1] array, run.py:375) crosses PCIe as
        # [N,C,1,1] and is tiled on the device by the build function (G_res): stride-0 views are collapsed here
        in_arrays = [_collapse_broadcast(a) for a in in_arrays]
        dev = self.rt.device
        out_arrays = None
        # Software pipeline over minibatches: (1) threaded host copy of minibatch k+1 into pinned staging,
        # (2) its H2D on a copy stream, (3) compute of minibatch k on the current stream, (4) D2H of its result
        # into pinned staging, (5) host copy of minibatch k-1's result - all overlapped; two staging slots.
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy_stream = self._copy_stream
        bounds = [(b, min(b + minibatch_size, num_items)) for b in range(0, num_items, minibatch_size)]
        # inputs that already live in page-locked memory (e.g. torch.empty(..., pin_memory=True).numpy()) are copied
        # to the device straight from the caller's array; pageable ones go through the pinned staging slots
        pinned_src = [_as_pinned(a) for a in in_arrays]
        slot_free = [None, None]          # event: the H2D that last read input slot s has finished

        out_pinned = None                 # results land directly in page-locked arrays that are handed to the caller

        use_graph = not os.environ.get('TMX_NO_GRAPH')
        conv = (out_mul, out_add, out_shrink, out_dtype)
        slot_done = [None, None]          # event: the compute that last read graph slot s has finished
        for k, (mb_begin, mb_end) in enumerate(bounds):
            if print_progress:
                print('\r%d / %d' % (mb_begin, num_items), end='')
            slot = k & 1
            if slot_free[slot] is not None:
                slot_free[slot].synchronize()
            shapes = [(mb_end - mb_begin,) + tuple(src.shape[1:]) for src in in_arrays]
            graph = self._forward_graph(slot, shapes, conv, dynamic_kwargs) if use_graph else None
            mb_in = []
            with torch.cuda.stream(copy_stream):
                if graph is not None and slot_done[slot] is not None:
                    copy_stream.wait_event(slot_done[slot])       # the graph's static inputs are free again
                for i, src in enumerate(in_arrays):
                    if pinned_src[i] is not None:
                        stage = pinned_src[i][mb_begin:mb_end]                # caller's page-locked array: DMA from it
                    else:
                        stage = self._pinned(('in', i, slot), shapes[i], torch.float32)
                        _parallel_copy(stage.numpy(), src[mb_begin:mb_end])   # host copy + cast into pinned memory
                    if graph is not None:
                        graph[1][i].copy_(stage, non_blocking=True)           # H2D straight into the static input
                    else:
                        mb_in.append(stage.to(dev, non_blocking=True))
                ev_in = torch.cuda.Event()
                ev_in.record(copy_stream)
            slot_free[slot] = ev_in
            main.wait_event(ev_in)
            if graph is not None:
                graph[0].replay()                                             # one launch for the whole forward
                mb_out = graph[2]
                slot_done[slot] = torch.cuda.Event()
                slot_done[slot].record(main)
            else:
                for t in mb_in:
                    t.record_stream(main)
                mb_out = self.get_output_for(*mb_in, return_as_list=True, **dynamic_kwargs)
                mb_out = [_convert_output(x, *conv) for x in mb_out]
            if out_pinned is None:
                out_pinned = [torch.empty([num_items] + list(x.shape[1:]), dtype=x.dtype, pin_memory=True)
                              for x in mb_out]
            for dst, x in zip(out_pinned, mb_out):
                dst[mb_begin:mb_end].copy_(x, non_blocking=True)
        main.synchronize()
        out_arrays = [t.numpy() for t in out_pinned]
        if print_progress:
            print('\r%d / %d' % (num_items, num_items))
        if not return_as_list:
            out_arrays = out_arrays[0] if len(out_arrays) == 1 else tuple(out_arrays)
        return out_arrays

    # -------------------------------------------------------------- run(num_gpus=k): in-process device replicas
    def _replica_on(self, index):
        """This network on cuda:`index`: itself, or a cached copy whose variables follow this network's (refreshed
        by one device-to-device copy of the flat buffer whenever the variable version changed)."""
        if index == self.rt.device.index:
            return self
        owner = self._owner()
        ent = self._replicas.get(index)
        if ent is None:
            net = object.__new__(Network)
            net._init_fields()
            net.name = self.name
            net.static_kwargs = dict(self.static_kwargs)
            net._build_module_src = self._build_module_src
            net._build_func_name = self._build_func_name
            net._build_func = self._build_func
            net._device = index
            with torch.cuda.device(index):
                net._init_graph()
            assert net._flat.numel() == owner._flat.numel(), 'replica layout differs'
            ent = self._replicas[index] = [net, -1]
        if ent[1] != owner._version or ent[0]._lod_host != owner._lod_host:
            with torch.cuda.device(index):
                ent[0]._flat.copy_(owner._flat)
                ent[0]._lod_host = owner._lod_host
                ent[0]._touch()
            ent[1] = owner._version
        return ent[0]

    def _run_multi_gpu(self, in_arrays, return_as_list, print_progress, minibatch_size, num_gpus, conv_kwargs,
                       dynamic_kwargs):
        """tfutil.py:644-661: `tf.split(x, num_gpus)` of every minibatch, one part per device, outputs concatenated
        in order.  Device g gets part g of EVERY minibatch (so minibatch-dependent layers - the critic's minibatch
        stddev groups - see the same sub-batches as in the reference); the k devices run concurrently, each through
        the single-device `run` (pinned staging, copy/compute overlap, CUDA-graphed forward) in its own host thread.
        A minibatch that k does not divide is split as evenly as possible (tf.split would refuse it)."""
        import concurrent.futures
        avail = torch.cuda.device_count()
        if num_gpus < 1 or num_gpus > avail:
            raise RuntimeError('Network.run(num_gpus=%d): this process sees %d CUDA device(s)' % (num_gpus, avail))
        base = self.rt.device.index
        devices = [base] + [i for i in range(avail) if i != base][:num_gpus - 1]
        num_items = in_arrays[0].shape[0]
        if minibatch_size is None:
            minibatch_size = num_items
        bounds = [(b, min(b + minibatch_size, num_items)) for b in range(0, num_items, minibatch_size)]
        # item ranges of device g: part g of each minibatch
        parts = [[] for _ in devices]
        for b, e in bounds:
            cuts = np.linspace(b, e, num_gpus + 1).round().astype(np.int64) if (e - b) % num_gpus else \
                np.arange(b, e + 1, (e - b) // num_gpus)
            for g in range(num_gpus):
                if cuts[g + 1] > cuts[g]:
                    parts[g].append((int(cuts[g]), int(cuts[g + 1])))
        per_dev_mb = max(1, -(-minibatch_size // num_gpus))

        def job(g):
            if not parts[g]:
                return None
            with torch.cuda.device(devices[g]):
                net = self._replica_on(devices[g])
                if len(parts[g]) == 1:
                    b, e = parts[g][0]
                    ins = [a[b:e] for a in in_arrays]
                else:
                    ins = [np.concatenate([a[b:e] for b, e in parts[g]], axis=0) for a in in_arrays]
                return net.run(*ins, return_as_list=True, minibatch_size=per_dev_mb, num_gpus=1, **conv_kwargs,
                               **dynamic_kwargs)
        # replicas, their weight copies and their CUDA graphs are made on the calling thread, one after the other: a
        # stream capture does not tolerate another host thread synchronising or allocating meanwhile
        conv = (conv_kwargs['out_mul'], conv_kwargs['out_add'], conv_kwargs['out_shrink'], conv_kwargs['out_dtype'])
        for g, d in enumerate(devices):
            with torch.cuda.device(d):
                net = self._replica_on(d)
                n_g = sum(e - b for b, e in parts[g])
                if n_g and not os.environ.get('TMX_NO_GRAPH'):
                    for k, b in enumerate(range(0, n_g, per_dev_mb)):
                        shapes = [(min(b + per_dev_mb, n_g) - b,) + tuple(_collapse_broadcast(a).shape[1:])
                                  for a in in_arrays]
                        net._forward_graph(k & 1, shapes, conv, dynamic_kwargs)
                    torch.cuda.synchronize(d)
        with concurrent.futures.ThreadPoolExecutor(max_workers=num_gpus) as ex:
            results = list(ex.map(job, range(num_gpus)))
        first = next(r for r in results if r is not None)
        out_arrays = [np.empty((num_items,) + tuple(x.shape[1:]), x.dtype) for x in first]
        for g, res in enumerate(results):
            if res is None:
                continue
            off = 0
            for b, e in parts[g]:
                for dst, src in zip(out_arrays, res):
                    dst[b:e] = src[off:off + (e - b)]
                off += e - b
        if print_progress:
            print('\r%d / %d' % (num_items, num_items))
        if not return_as_list:
            out_arrays = out_arrays[0] if len(out_arrays) == 1 else tuple(out_arrays)
        return out_arrays

    def _forward_graph(self, slot, shapes, conv, dynamic_kwargs):
        """CUDA graph of get_output_for (+ output conversion) for fixed input shapes and the current weight version:
        (graph, static inputs, static outputs).  ~20 kernel launches and as many allocations become one replay, which
        is what lets `run` pipeline small minibatches without being bound by host launch overhead."""
        o = self._owner()
        key = (slot, tuple(shapes), conv, tuple(sorted(dynamic_kwargs.items())))
        ent = self._graphs.get(key)
        if ent is not None and ent[3] == o._version:
            return ent
        dev = self.rt.device
        static_in = [torch.zeros(sh, dtype=torch.float32, device=dev) for sh in shapes]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):       # warm-up outside the capture: weight planes, kernel attributes
            outs = self.get_output_for(*static_in, return_as_list=True, **dynamic_kwargs)
            [_convert_output(x, *conv) for x in outs]
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        # no cyclic garbage collection inside the capture: finalising unrelated CUDA objects (older graphs, events)
        # from another test or caller while the stream is capturing invalidates it.  One capture at a time per
        # process, in thread-local error mode: run(num_gpus=k) builds its replicas' graphs from k host threads while
        # the other devices already copy and compute
        with _CAPTURE_LOCK:
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                # (an explicit capture stream: torch's default one is created once, on whatever device was current)
                with torch.cuda.graph(g, stream=side, capture_error_mode='thread_local'):
                    outs = self.get_output_for(*static_in, return_as_list=True, **dynamic_kwargs)
                    static_out = [_convert_output(x, *conv) for x in outs]
            finally:
                if gc_was_on:
                    gc.enable()
        if len(self._graphs) > 8:
            # evict: a dropped graph's private pool (its static inputs / outputs) may still be read by a replay or a
            # D2H copy in flight - wait for the device before the memory can be handed out again
            torch.cuda.synchronize(dev)
            stale = [k for k, e in self._graphs.items() if e[3] != o._version]
            for k in (stale or list(self._graphs)):
                del self._graphs[k]
        ent = (g, static_in, static_out, o._version)
        self._graphs[key] = ent
        return ent

    def _pinned(self, key, shape, dtype):
        """Page-locked staging buffers, kept per (slot, shape) so that repeated
        `run` calls do not re-register host memory."""
        k = (key, tuple(shape), dtype)
        buf = self._staging.get(k)
        if buf is None:
            buf = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
            self._staging[k] = buf
        return buf

    # -------------------------------------------------------------- EMA / pickling
    def setup_as_moving_average_of(self, src_net, beta=0.99, beta_nontrainable=0.0):
        """tfutil.py:611-621.  Returns a callable update op: var <- lerp(src, var, beta)."""
        same_layout = list(self.vars.keys()) == list(src_net.vars.keys()) and \
            all(a.shape == b.shape for a, b in zip(self.vars.values(), src_net.vars.values()))

        def update_op():
            if same_layout and self.flat.is_cuda and beta_nontrainable == 0.0:
                # one fused kernel over the flat buffers; non-trainables (lod) then follow beta_nontrainable
                import ctypes as C
                from . import _lib
                rt = self.rt
                _lib.check(rt.lib.tmx_ema_update(rt.handle, C.c_void_p(src_net.flat.data_ptr()),
                                                 C.c_void_p(self.flat.data_ptr()), self.flat.numel(), float(beta),
                                                 rt.stream()), 'tmx_ema_update')
                for name, var in self.vars.items():
                    if name not in self.trainables:
                        s = src_net.vars[name].value
                        var.value.copy_(s + (var.value - s) * beta_nontrainable)
            else:
                for name, var in self.vars.items():
                    if name in src_net.vars:
                        cur_beta = beta if name in self.trainables else beta_nontrainable
                        s = src_net.vars[name].value
                        var.value.copy_(s + (var.value - s) * cur_beta)
            self._lod_host = src_net._owner()._lod_host
            self._touch()
        return update_op

    def __getstate__(self):
        return {'version': 2, 'name': self.name, 'static_kwargs': self.static_kwargs,
                'build_module_src': self._build_module_src, 'build_func_name': self._build_func_name,
                'variables': [(k, self.get_var(k)) for k in self.vars.keys()]}

    def __setstate__(self, state):
        """tfutil.py:553-576, except that the pickled module SOURCE is never exec'd
        (SURVEY Appendix D): the build function is resolved by name in this package."""
        self._init_fields()
        assert state['version'] == 2
        self.name = state['name']
        self.static_kwargs = state['static_kwargs']
        self._build_module_src = state['build_module_src']
        self._build_func_name = state['build_func_name']
        self._build_func = import_obj('networks.' + self._build_func_name)
        if not torch.cuda.is_available():
            self._device = 'cpu'     # variables only (checkpoint inspection/conversion); evaluation still needs CUDA
        self._init_graph()
        self.reset_vars(0)
        self.set_vars({name: value for name, value in state['variables'] if name in self.vars})

    def print_layers(self, title=None, hide_layers_with_no_params=False):
        title = title or self.name
        print()
        print('%-32s%-12s%-24s' % (title, 'Params', 'WeightShape'))
        print('%-32s%-12s%-24s' % (('---',) * 3))
        total = 0
        for k, v in self.trainables.items():
            total += v.size
            print('%-32s%-12s%-24s' % (k, v.size, list(v.shape)))
        print('%-32s%-12s' % ('Total', total))
        print()

    def setup_weight_histograms(self, title=None):
        pass  # TensorBoard summaries are outside the hot path


def _collapse_broadcast(a):
    """[N,C,1,1] view of an array whose last two axes are stride-0 broadcasts (np.broadcast_to), else `a`."""
    if isinstance(a, np.ndarray) and a.ndim == 4 and a.shape[2] * a.shape[3] > 1 and a.strides[2] == 0 and a.strides[3] == 0:
        return a[:, :, :1, :1]
    return a


def _as_pinned(a):
    """torch view of a float32 C-contiguous numpy array if it sits in page-locked memory, else None."""
    if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags['C_CONTIGUOUS'] and a.flags['ALIGNED']):
        return None
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')       # read-only arrays: we only read
            t = torch.from_numpy(a)
        return t if t.is_pinned() else None
    except (TypeError, ValueError, RuntimeError):
        return None


_COPY_POOL = None


def _parallel_copy(dst, src, min_bytes=4 << 20, workers=8):
    """dst[...] = src for large host arrays, split over a few threads (numpy releases the GIL while copying):
    a single-threaded memcpy of the 64 MB latent batch would cost more than the whole GPU step."""
    global _COPY_POOL
    n = dst.shape[0]
    if dst.nbytes < min_bytes or n < 2:
        dst[...] = src
        return
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=workers)
    parts = min(workers, n)
    edges = [n * i // parts for i in range(parts + 1)]

    def one(i):
        dst[edges[i]:edges[i + 1]] = src[edges[i]:edges[i + 1]]
    list(_COPY_POOL.map(one, range(parts)))


def _np_dtype(t):
    return {torch.float32: np.float32, torch.uint8: np.uint8, torch.int32: np.int32, torch.float16: np.float16,
            torch.int16: np.int16, torch.int64: np.int64}[t.dtype]


def _convert_output(x, out_mul, out_add, out_shrink, out_dtype):
    """tfutil.py:649-659 output conversion on the device (tmx_convert_output): x * mul + add, avg-pool shrink,
    tf.round (half to even) + saturate_cast for integer dtypes - one kernel, uint8 written directly."""
    dt = None if out_dtype is None else np.dtype(out_dtype)
    if out_mul == 1.0 and out_add == 0.0 and out_shrink == 1 and (dt is None or dt == np.dtype(np.float32)):
        return x
    import ctypes as C
    from . import _lib
    rt = Runtime.get(x.device)
    x = x.contiguous()
    lead, (h, w) = x.shape[:-2], x.shape[-2:]
    is_int = dt is not None and np.issubdtype(dt, np.integer)
    kind = 1 if dt == np.dtype(np.uint8) else (2 if is_int else 0)
    out = torch.empty(tuple(lead) + (h // out_shrink, w // out_shrink),
                      dtype=torch.uint8 if kind == 1 else torch.float32, device=x.device)
    _lib.check(rt.lib.tmx_convert_output(rt.handle, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                         int(np.prod(lead)), h, w, float(out_mul), float(out_add), int(out_shrink),
                                         kind, rt.stream()), 'tmx_convert_output')
    if dt is None or kind == 1 or dt == np.dtype(np.float32):
        return out
    if is_int:                       # other integer types: rounded on the device, narrowed with saturation here
        info = np.iinfo(dt)
        out = out.clamp_(float(info.min), float(info.max))
    return out.to({np.dtype(np.int32): torch.int32, np.dtype(np.int16): torch.int16, np.dtype(np.int64): torch.int64,
                   np.dtype(np.float16): torch.float16, np.dtype(np.float64): torch.float64}[dt])
