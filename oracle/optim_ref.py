"""ORACLE (test infrastructure): numpy restatement of the reference's optimizer
step and weight EMA.

* `tfutil.Optimizer.apply_updates` (tfutil.py:304-372): sum gradients over
  towers, scale by 1/total_grads when total_grads > 1, skip the update when any
  gradient is non-finite.
* `tf.train.AdamOptimizer` as TF 1.12 implements it (the arithmetic lives in
  TensorFlow, not in the reference tree - restated from the TF documentation,
  parity unpinned): m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2;
  lr_t = lr sqrt(1-b2^t)/(1-b1^t); w <- w - lr_t m / (sqrt(v) + eps).
* `Network.setup_as_moving_average_of` (tfutil.py:611-621): var <- lerp(src, var, beta).
All arithmetic in float32 like the TF kernels."""
import numpy as np

f32 = np.float32


class AdamState:
    def __init__(self, n, beta1, beta2):
        self.m = np.zeros(n, f32)
        self.v = np.zeros(n, f32)
        self.p1, self.p2 = f32(beta1), f32(beta2)     # beta powers, advance on applied steps only


def optimizer_step(w, tower_grads, state, lr, beta1=0.0, beta2=0.99, eps=1e-8):
    """w: float32 [n] (updated in place); tower_grads: list of float32 [n], one per tower.
    Returns True when the step was applied, False when skipped (non-finite gradient)."""
    g = tower_grads[0].astype(f32).copy()
    for t in tower_grads[1:]:
        g = g + t.astype(f32)                                   # nccl.all_sum, tfutil.py:326-333
    if len(tower_grads) > 1:
        g = g * f32(1.0 / len(tower_grads))                     # tfutil.py:340-344
    if not np.all(np.isfinite(g)):                              # tfutil.py:347-355
        return False
    b1, b2, e = f32(beta1), f32(beta2), f32(eps)
    lr_t = f32(lr) * np.sqrt(f32(1) - state.p2) / (f32(1) - state.p1)
    state.m[:] = b1 * state.m + (f32(1) - b1) * g
    state.v[:] = b2 * state.v + (f32(1) - b2) * g * g
    w[:] = w - lr_t * state.m / (np.sqrt(state.v) + e)
    state.p1 = state.p1 * b1
    state.p2 = state.p2 * b2
    return True


def ema_update(src, dst, beta):
    """tfutil.py:617: new = lerp(src, cur, beta) = src + (cur - src) * beta."""
    dst[:] = src + (dst - src) * f32(beta)
