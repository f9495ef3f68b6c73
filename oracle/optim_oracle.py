"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's optimizer step (checker for btslpg_adam_step).

What is restated, and where it lives in the reference:
  custom_optimizers.py:47-59   AdamW: per-variable decoupled decay `var <- var - lr * (l1*sign(var) + l2*var)` applied BEFORE
                               the parent's update (l1 and l2 both non-zero / only l1 / else l2 -- the three branches of :49-54)
  tf.keras.optimizers.Adam     third-party TensorFlow (requirements.txt:1 `tensorflow>=2.1.0`, not under /root/reference):
                               the published algorithm of tf.raw_ops.ResourceApplyAdam,
                                   alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t),  t = iterations + 1
                                   m <- m + (g - m)(1 - beta1);  v <- v + (g*g - v)(1 - beta2);  var <- var - alpha*m / (sqrt(v) + epsilon)
  bts_train.py:125-131         start_lr = lr * replicas, end_lr = 0.1 * start_lr unless given;
                               lr(step) = (start - end) * (1 - min(step, total)/total)**0.9 + end, cast to float32
  custom_callbacks.py:46-50    the schedule is evaluated at the 0-based global step at batch begin
  bts_train.py:206             MirroredStrategy averages the replicas' gradients: grad_scale = 1/N on the summed gradient

Parity status: the AdamW decay and the schedule are pinned to the reference's own lines; the Adam update itself is
**parity unpinned** against a TensorFlow run (TensorFlow cannot be installed here) -- it restates the published op.
"""
import numpy as np


def poly_lr(step, lr_start, lr_end, total_steps, power=0.9):
    """bts_train.py:129-131 (float64 arithmetic, float32 result like the tf.cast)."""
    if total_steps <= 0:
        return np.float32(lr_start)
    frac = min(float(step), float(total_steps)) / float(total_steps)
    return np.float32((float(lr_start) - float(lr_end)) * (1.0 - frac) ** power + float(lr_end))


def adamw_step(p, g, m, v, step, lr_start, lr_end=None, total_steps=0, power=0.9, beta1=0.9, beta2=0.999, epsilon=1e-3,
               l1=0.0, l2=0.0, grad_scale=1.0, dtype=np.float64):
    """One update at 0-based global step `step`.  Returns (p, m, v, lr) as new arrays of `dtype`."""
    if lr_end is None:
        lr_end = lr_start * 0.1 if total_steps > 0 else lr_start
    lr = poly_lr(step, lr_start, lr_end, total_steps, power)
    p, g, m, v = (np.asarray(a, dtype=dtype).copy() for a in (p, g, m, v))
    g = g * dtype(grad_scale)
    lr_t = dtype(lr)
    if l1 != 0 or l2 != 0:                                       # custom_optimizers.py:47-59
        if l1 != 0 and l2 != 0:
            decay = dtype(l1) * np.sign(p) + dtype(l2) * p
        elif l1 != 0:
            decay = dtype(l1) * np.sign(p)
        else:
            decay = dtype(l2) * p
        p = p - lr_t * decay
    # Keras holds beta_1, beta_2 and epsilon as tensors of the VARIABLE's dtype (float32): the values that enter the arithmetic
    # are float32(0.9), float32(0.999), float32(1e-3), not the Python doubles
    b1, b2, eps = float(np.float32(beta1)), float(np.float32(beta2)), float(np.float32(epsilon))
    t = step + 1
    alpha = dtype(float(lr) * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    m = m + (g - m) * dtype(1.0 - b1)
    v = v + (g * g - v) * dtype(1.0 - b2)
    p = p - alpha * m / (np.sqrt(v) + dtype(eps))
    return p, m, v, lr


def png16(depth, max_depth):
    """bts_predict.py:140-141, literally (numpy float32 arithmetic and numpy's own uint16 cast)."""
    pred_depth_scaled = np.asarray(depth, np.float32) * 65536 / np.float32(max_depth)
    with np.errstate(invalid="ignore"):
        return pred_depth_scaled.astype(np.uint16)
