"""`Optimizer`: the reference's tfutil.Optimizer (tfutil.py:246-399) over flat
device buffers.  Gradients of a network are written by the hand-written backward
passes into one fp32 buffer laid out like the network's variable buffer
(`Network.flat`); `apply_updates` then does, per optimizer:

  1. mark LOCAL non-finite gradients in a float slot                 (tfutil.py:347-355, made global: SURVEY §8e)
  2. SUM all-reduce across ranks                                     (tfutil.py:326-333)
  3. scale by 1 / total registrations                                (tfutil.py:340-344)
  4. skip the whole update if the summed mark is > 0 on any rank
  5. TF1 Adam, beta powers advancing only on applied steps           (tf.train.AdamOptimizer)

Steps 3-5 are ONE fused kernel per network (`tmx_adam_update`) plus the power update (`tmx_adam_advance`).

Two ways to run step 2:
  * stand-alone (`apply_updates`): one collective per registered network buffer + one for the mark;
  * bucketed (`GradientBucket`): the gradient buffers of ALL optimizers of a training phase are views into one flat
    buffer whose tail holds the marks, so the phase costs ONE all-reduce (reference: one `nccl.all_sum` per
    variable - 96 for the E/G optimizer, 38 per critic).  The trainer calls `mark()` on every optimizer of the
    phase, `bucket.allreduce()`, then `update()`."""
import ctypes as C

import torch

from . import _lib, parallel
from .runtime import Runtime


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class GradientBucket:
    """One flat fp32 device buffer holding the gradient buffers of several networks back to back (each laid out like
    its `Network.flat`, 256-B aligned) and, in its last 64 floats, one non-finite mark per optimizer of the phase."""
    TAIL = 64

    def __init__(self, nets):
        """nets: {name: Network}.  `views[name]` is the flat gradient buffer to hand to the backward passes and to
        `Optimizer.register_gradients`."""
        sizes = {k: n.flat.numel() for k, n in nets.items()}
        dev = next(iter(nets.values())).flat.device
        self.flat = torch.zeros(sum(sizes.values()) + self.TAIL, dtype=torch.float32, device=dev)
        self.views, off = {}, 0
        for k, sz in sizes.items():
            assert off % 64 == 0
            self.views[k] = self.flat[off:off + sz]
            off += sz
        self.marks = self.flat[off:off + self.TAIL]
        self._next_mark = 0
        self.collectives = 0

    def mark_slot(self):
        assert self._next_mark < self.TAIL
        m = self.marks[self._next_mark:self._next_mark + 1]
        self._next_mark += 1
        return m

    def zero_(self):
        """Gradients and marks of the whole phase in one fill."""
        self.flat.zero_()

    def allreduce(self):
        """ONE SUM all-reduce for the phase (no-op in a single-process job)."""
        self.collectives += parallel.allreduce_sum_([self.flat])

    def allreduce_async(self):
        """The same collective, issued without blocking the current stream; returns wait() (parallel.allreduce_sum_async_)."""
        self.collectives += 1 if parallel.world_size() > 1 else 0
        return parallel.allreduce_sum_async_(self.flat)


class Optimizer:
    def __init__(self, name='Train', tf_optimizer='tf.train.AdamOptimizer', learning_rate=0.001,
                 use_loss_scaling=False, beta1=0.9, beta2=0.999, epsilon=1e-8, **kwargs):
        if tf_optimizer != 'tf.train.AdamOptimizer':
            raise NotImplementedError('only tf.train.AdamOptimizer is used by the reference (config.py:84-87)')
        if use_loss_scaling:
            raise NotImplementedError('fp16 loss scaling is off in the reference config (fp32 path)')
        self.name = name
        self.learning_rate = float(learning_rate)
        self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)
        self._nets = []            # [[net, flat_grad, m, v, registrations]]
        self._powers = None
        self._mark = None          # float32 [1]: > 0 = this step is skipped; lives in a GradientBucket's tail if any
        self._own_mark = True
        self.skipped_steps = 0

    # ------------------------------------------------------------------ registration
    def register_gradients(self, net, flat_grad):
        """`flat_grad`: fp32 device buffer shaped like `net.flat` holding d loss / d variables
        (zeros for disconnected variables, tfutil.py:298).  Registering the same network again
        accumulates (tf.add_n of tfutil.py:315)."""
        assert flat_grad.shape == net.flat.shape and flat_grad.dtype == torch.float32
        for ent in self._nets:
            if ent[0] is net:
                if ent[1] is not flat_grad:
                    raise ValueError('one flat gradient buffer per network: accumulate further losses into it')
                ent[4] += 1
                return
        m = torch.zeros_like(net.flat)
        v = torch.zeros_like(net.flat)
        self._nets.append([net, flat_grad, m, v, 1])

    def use_bucket(self, bucket):
        """The non-finite mark of this optimizer travels in `bucket`'s tail (all registered gradient buffers must be
        views of that bucket)."""
        for ent in self._nets:
            g = ent[1]
            lo, hi = bucket.flat.data_ptr(), bucket.flat.data_ptr() + bucket.flat.numel() * 4
            if not (lo <= g.data_ptr() and g.data_ptr() + g.numel() * 4 <= hi):
                raise ValueError('%s: a registered gradient buffer is not a view of the bucket' % self.name)
        self._mark = bucket.mark_slot()
        self._own_mark = False

    def reset_optimizer_state(self):
        """tfutil.py:375-376: zero the Adam slots and restart the beta powers."""
        for _, _, m, v, _ in self._nets:
            m.zero_()
            v.zero_()
        self._powers = None

    # ------------------------------------------------------------------ update
    def _setup(self):
        assert self._nets, 'no gradients registered'
        rt = Runtime.get(self._nets[0][0].flat.device)
        if self._powers is None:
            # (a host list -> device copy: outside any CUDA-graph capture - the trainer's first step runs eagerly)
            self._powers = torch.tensor([self.beta1, self.beta2], dtype=torch.float32, device=rt.device)
        if self._mark is None:
            self._mark = torch.zeros(1, dtype=torch.float32, device=rt.device)
        return rt

    def mark(self):
        """Step 1: flag local non-finite gradients (the slot was zeroed with the bucket, or here when stand-alone)."""
        rt = self._setup()
        if self._own_mark:
            self._mark.zero_()
        st = rt.stream()
        for _, g, _, _, _ in self._nets:
            _lib.check(rt.lib.tmx_nonfinite_mark(rt.handle, _ptr(g), g.numel(), _ptr(self._mark), st),
                       'tmx_nonfinite_mark')

    def update(self, learning_rate=None):
        """Steps 3-5 on already reduced gradients.  Returns the float mark (device, > 0 = the step was skipped on
        every rank: overflow_frequency, tfutil.py:365)."""
        rt = self._setup()
        lr = self.learning_rate if learning_rate is None else float(learning_rate)
        counts = {ent[4] for ent in self._nets}
        if len(counts) != 1:
            raise ValueError('%s: networks registered a different number of times (%s); tfutil.py:340-344 scales '
                             'every gradient by one common 1/num_towers' % (self.name, sorted(counts)))
        total = counts.pop() * parallel.world_size()
        scale = 1.0 / total if total > 1 else 1.0
        st = rt.stream()
        for net, g, m, v, _ in self._nets:
            # every network of this optimizer steps with the SAME beta powers; they advance once, after the last
            _lib.check(rt.lib.tmx_adam_update(rt.handle, _ptr(net.flat), _ptr(g), _ptr(m), _ptr(v), g.numel(), lr,
                                              self.beta1, self.beta2, self.epsilon, scale, _ptr(self._powers), None,
                                              _ptr(self._mark), st), 'tmx_adam_update')
            net.mark_variables_changed()
        _lib.check(rt.lib.tmx_adam_advance(rt.handle, _ptr(self._powers), self.beta1, self.beta2, None,
                                           _ptr(self._mark), st), 'tmx_adam_advance')
        return self._mark

    def apply_updates(self, learning_rate=None):
        """tfutil.py:277-372 stand-alone: mark, all-reduce (one collective per network buffer + the mark), update."""
        if not self._own_mark:
            raise RuntimeError('%s is bucketed: call mark(), bucket.allreduce(), update()' % self.name)
        self.mark()
        parallel.allreduce_sum_([ent[1] for ent in self._nets] + [self._mark])
        return self.update(learning_rate)
