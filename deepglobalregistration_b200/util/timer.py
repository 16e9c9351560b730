"""Wall-clock stage timers with the reference's Timer/AverageMeter surface
(util/timer.py:12-54): tic(), toc(average=True), .avg, .sum, .count - read by
scripts/test_kitti.py:83,92-93 through DeepGlobalRegistration.feat_timer / reg_timer."""
import time


class AverageMeter:
  def __init__(self):
    self.reset()

  def reset(self):
    self.val = self.avg = 0.0
    self.sum = self.sq_sum = 0.0
    self.count = 0
    self.var = 0.0

  def update(self, val, n=1):
    self.val = val
    self.sum += val * n
    self.sq_sum += val * val * n
    self.count += n
    self.avg = self.sum / self.count
    self.var = self.sq_sum / self.count - self.avg ** 2


class Timer(AverageMeter):
  def tic(self):
    self.start_time = time.time()

  def toc(self, average=True):
    self.diff = time.time() - self.start_time
    self.update(self.diff)
    return self.avg if average else self.diff
