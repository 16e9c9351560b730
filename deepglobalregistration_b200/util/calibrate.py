"""BatchNorm calibration for seeded random-init checkpoints.

A random-init ResUNet in eval mode with untouched BatchNorm statistics produces
low-contrast features (every ReLU layer adds a common positive mean, so after L2
normalisation all FCGF rows point the same way and feature matching is arbitrary).  Real
checkpoints carry running statistics that whiten every layer.  ``calibrate_batchnorm``
gives a synthetic checkpoint the same property: one operator-by-operator forward pass in
which every MinkowskiBatchNorm first records the mean/variance of its input as its running
statistics.  Used by tests and the benchmark to obtain well-conditioned synthetic pairs;
the calibrated state dict is what both the CUDA path and the CPU oracle then load."""
import torch

from .. import me as ME


@torch.no_grad()
def calibrate_batchnorm(model, sinput):
  hooks = []

  def pre_hook(mod, args):
    x = args[0].F
    mod.bn.running_mean.copy_(x.mean(0))
    mod.bn.running_var.copy_(x.var(0, unbiased=False).clamp_min(1e-6))
    mod.bn.weight.fill_(1.0)
    mod.bn.bias.zero_()

  for m in model.modules():
    if isinstance(m, ME.MinkowskiBatchNorm):
      hooks.append(m.register_forward_pre_hook(pre_hook))
  try:
    model.eval()
    return model(sinput)
  finally:
    for h in hooks:
      h.remove()
