"""ME.MinkowskiFunctional: only ``relu`` is on the hot path (model/resunet.py:602-640,
model/residual_block.py:123,132)."""
from .. import _abi


def relu(x, *a, **k):
  import torch
  if torch.is_grad_enabled() and x.F.requires_grad:      # training path: autograd through torch.relu
    return x._like(torch.relu(x.F))
  return x._like(_abi.affine_act(x.F, relu=True))
