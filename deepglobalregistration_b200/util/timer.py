"""Stage timers exposing the attributes the reference's callers read from
``DeepGlobalRegistration.feat_timer`` / ``reg_timer`` (scripts/test_kitti.py:83,92-93 use
``tic()``, ``toc()``, ``.avg``; the reference class is util/timer.py:40-54).

Host wall clock by default; ``Timer(sync=True)`` brackets the interval with a CUDA
synchronisation so that it covers the GPU work enqueued in between (the reference's timers are
only correct where an ``.item()`` happens to synchronise)."""
import math
import time


class RunningStat:
  """Count / mean / variance of a stream of samples (Welford), plus the last value."""

  __slots__ = ('count', 'avg', '_m2', 'sum', 'val')

  def __init__(self):
    self.reset()

  def reset(self):
    self.count, self.avg, self._m2, self.sum, self.val = 0, 0.0, 0.0, 0.0, 0.0

  def update(self, value, n=1):
    # arrays / tensors as the reference's AverageMeter takes them (util/timer.py:25-37): the sample is
    # the mean, its multiplicity the element count
    size = getattr(value, 'size', None)
    if callable(size):            # torch.Tensor
      n, value = int(value.numel()), float(value.float().mean()) if value.numel() else 0.0
    elif size is not None and not isinstance(value, (int, float)):   # numpy array / scalar
      import numpy as np
      arr = np.asarray(value, dtype=np.float64)
      n, value = (int(arr.size), float(arr.mean())) if arr.ndim else (n, float(arr))
    value = float(value)
    n = int(n)
    if n > 0:                     # n equal samples merged in one step (Chan et al. pairwise update)
      total = self.count + n
      delta = value - self.avg
      self.avg += delta * n / total
      self._m2 += delta * delta * self.count * n / total
      self.count = total
    self.sum += value * n
    self.val = value

  @property
  def var(self):
    return self._m2 / self.count if self.count else 0.0

  @property
  def std(self):
    return math.sqrt(self.var)


class Timer(RunningStat):
  __slots__ = ('sync', '_t0', 'diff')

  def __init__(self, sync=False):
    super().__init__()
    self.sync, self._t0, self.diff = sync, None, 0.0

  def _now(self):
    if self.sync:
      import torch
      if torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.perf_counter()

  def tic(self):
    self._t0 = self._now()

  @property
  def start_time(self):          # attribute names of the reference's Timer (util/timer.py:40-54)
    return self._t0

  @property
  def sq_sum(self):
    return self._m2 + self.count * self.avg * self.avg

  def toc(self, average=True):
    if self._t0 is None:
      raise RuntimeError('toc() without tic()')
    self.diff = self._now() - self._t0
    self.update(self.diff)
    return self.avg if average else self.diff


AverageMeter = RunningStat
