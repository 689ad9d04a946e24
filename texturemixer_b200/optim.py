"""`Optimizer`: the reference's tfutil.Optimizer (tfutil.py:246-399) over flat
device buffers.  Gradients of a network are written by the hand-written backward
passes into one fp32 buffer laid out like the network's variable buffer
(`Network.flat`); `apply_updates` then does, per optimizer:

  1. SUM all-reduce across ranks, one collective per network     (tfutil.py:326-333)
  2. scale by 1 / total registrations                             (tfutil.py:340-344)
  3. skip the whole update if ANY gradient is non-finite           (tfutil.py:347-355)
  4. TF1 Adam, beta powers advancing only on applied steps         (tf.train.AdamOptimizer)

Steps 2-4 are the fused libtmx kernels `tmx_nonfinite_check` / `tmx_adam_step`."""
import ctypes as C

import torch

from . import _lib, parallel
from .runtime import Runtime


def _ptr(t):
    return C.c_void_p(t.data_ptr())


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
        self._nets = []            # [(net, flat_grad, m, v)]
        self._registrations = 0    # per-rank register_gradients calls per network
        self._powers = None
        self._flag = None
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

    def reset_optimizer_state(self):
        """tfutil.py:375-376: zero the Adam slots and restart the beta powers."""
        for _, _, m, v, _ in self._nets:
            m.zero_()
            v.zero_()
        self._powers = None

    # ------------------------------------------------------------------ update
    def apply_updates(self, learning_rate=None):
        assert self._nets, 'no gradients registered'
        lr = self.learning_rate if learning_rate is None else float(learning_rate)
        rt = Runtime.get(self._nets[0][0].flat.device)
        if self._powers is None:
            self._powers = torch.tensor([self.beta1, self.beta2], dtype=torch.float32, device=rt.device)
            self._flag = torch.zeros(1, dtype=torch.int32, device=rt.device)
        parallel.allreduce_sum_([ent[1] for ent in self._nets])
        total = self._nets[0][4] * parallel.world_size()
        scale = 1.0 / total if total > 1 else 1.0
        self._flag.zero_()
        st = rt.stream()
        for _, g, _, _, _ in self._nets:
            _lib.check(rt.lib.tmx_nonfinite_check(rt.handle, _ptr(g), g.numel(), _ptr(self._flag), st),
                       'tmx_nonfinite_check')
        powers_before = self._powers.clone()
        for i, (net, g, m, v, _) in enumerate(self._nets):
            # every network of this optimizer steps with the SAME beta powers; they advance once, after the last
            p = self._powers if i == len(self._nets) - 1 else powers_before.clone()
            _lib.check(rt.lib.tmx_adam_step(rt.handle, _ptr(net.flat), _ptr(g), _ptr(m), _ptr(v), g.numel(), lr,
                                            self.beta1, self.beta2, self.epsilon, scale, _ptr(p), _ptr(self._flag),
                                            st), 'tmx_adam_step')
            net.mark_variables_changed()
        return self._flag        # device int32: 1 = the step was skipped (overflow_frequency, tfutil.py:365)
