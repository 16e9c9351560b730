"""ORACLE (test infrastructure only) - a CPU ``MinkowskiEngine``-shaped module backed by
oracle/sparse_ops.py, with just the surface the reference's model files touch
(model/resunet.py, model/residual_block.py, model/common.py, model/__init__.py).

Purpose: run the reference's UNMODIFIED ``model/resunet.py`` graph on the CPU and check that
oracle/resunet.py - the hand-written restatement every GPU parity test compares against -
computes the same function (layer order, strides, skip-concatenation order, norm placement,
final normalisation).  That pins the oracle's GRAPH to the reference's own model code; the
sparse operators underneath both sides are oracle/sparse_ops.py, whose MinkowskiEngine semantics
remain unpinned (ME is not installable here).

    import sys; from oracle import me_cpu; me_cpu.install()     # sys.modules['MinkowskiEngine']
"""
import enum
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import sparse_ops as so


class RegionType(enum.Enum):
  HYPER_CUBE = 0
  HYPER_CROSS = 1


class KernelGenerator:
  def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False,
               region_type=RegionType.HYPER_CUBE, dimension=-1, **kw):
    assert region_type == RegionType.HYPER_CUBE and dilation == 1
    self.kernel_size, self.stride, self.is_transpose, self.dimension = kernel_size, stride, is_transpose, dimension


class SparseTensor:
  def __init__(self, features, coordinates=None, coordinate_map_key=None, coordinate_manager=None,
               device=None, coords=None, coords_key=None, coords_manager=None, tensor_stride=1):
    coordinates = coords if coordinates is None else coordinates
    coordinate_map_key = coords_key if coordinate_map_key is None else coordinate_map_key
    coordinate_manager = coords_manager if coordinate_manager is None else coordinate_manager
    self.F = features
    if coordinate_manager is None:
      coordinate_manager = so.CoordinateMaps(coordinates.cpu().numpy() if isinstance(coordinates, torch.Tensor)
                                             else np.asarray(coordinates))
      coordinate_map_key = 1
    self.coordinate_manager, self.coordinate_map_key = coordinate_manager, coordinate_map_key   # key = tensor stride

  coords_man = property(lambda self: self.coordinate_manager)
  coords_key = property(lambda self: self.coordinate_map_key)
  tensor_stride = property(lambda self: self.coordinate_map_key)
  D = property(lambda self: self.coordinate_manager.D)
  C = property(lambda self: torch.from_numpy(self.coordinate_manager.coords_at(self.coordinate_map_key)))

  def _like(self, feats, key=None):
    return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key if key is None else key,
                        coordinate_manager=self.coordinate_manager)

  def __iadd__(self, other):
    assert other.coordinate_map_key == self.coordinate_map_key
    self.F = self.F + other.F
    return self

  def __add__(self, other):
    return self._like(self.F + other.F)


class MinkowskiNetwork(nn.Module):
  def __init__(self, D):
    super().__init__()
    self.D = D


class _ConvBase(nn.Module):
  def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, has_bias=None,
               kernel_generator=None, dimension=-1):
    super().__init__()
    bias = bias if has_bias is None else has_bias
    self.kernel_size, self.stride, self.dimension = kernel_size, stride, dimension
    K = kernel_size ** dimension
    self.kernel = nn.Parameter(torch.zeros(K, in_channels, out_channels) if K > 1 else torch.zeros(in_channels, out_channels))
    self.bias = nn.Parameter(torch.zeros(1, out_channels)) if bias else None


class MinkowskiConvolution(_ConvBase):
  def forward(self, x):
    man, s = x.coordinate_manager, x.coordinate_map_key
    if self.kernel_size == 1:
      assert self.stride == 1
      return x._like(so.linear_forward(x.F, self.kernel.data, None if self.bias is None else self.bias.data))
    if self.stride == 1:
      buckets, out_s = man.same_map(s, self.kernel_size), s
    else:
      assert self.stride == 2
      buckets, out_s = man.down_map(s, self.kernel_size), 2 * s
    n_out = len(man.coords_at(out_s))
    return x._like(so.conv_forward(x.F, self.kernel.data, buckets, n_out,
                                   None if self.bias is None else self.bias.data), out_s)


class MinkowskiConvolutionTranspose(_ConvBase):
  def forward(self, x):
    man, s = x.coordinate_manager, x.coordinate_map_key
    if self.kernel_size == 1:
      assert self.stride == 1
      return x._like(so.linear_forward(x.F, self.kernel.data, None if self.bias is None else self.bias.data))
    assert self.stride == 2 and s % 2 == 0
    out_s = s // 2
    buckets = so.swap_map(man.down_map(out_s, self.kernel_size))
    return x._like(so.conv_forward(x.F, self.kernel.data, buckets, len(man.coords_at(out_s)),
                                   None if self.bias is None else self.bias.data), out_s)


class MinkowskiBatchNorm(nn.Module):
  def __init__(self, num_features, eps=1e-5, momentum=0.1, **kw):
    super().__init__()
    self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum)

  def forward(self, x):
    assert not self.training, 'the oracle stand-in is eval-only'
    return x._like(self.bn(x.F))


def cat(*tensors):
  assert all(t.coordinate_map_key == tensors[0].coordinate_map_key for t in tensors)
  return tensors[0]._like(torch.cat([t.F for t in tensors], 1))


def _relu(x):
  return x._like(torch.relu(x.F))


def sparse_quantize(coordinates, features=None, labels=None, return_index=False, **kw):
  """ME 0.5 form as the reference uses it (core/deep_global_registration.py:152): floor the
  (already divided) coordinates, keep the first row of every voxel, indices ascending.
  -> (unique int coordinates, index)."""
  assert features is None and labels is None
  is_tensor = isinstance(coordinates, torch.Tensor)
  c = np.floor(coordinates.numpy() if is_tensor else np.asarray(coordinates)).astype(np.int32)
  sel, _ = so._first_occurrence(c)
  uniq = torch.from_numpy(c[sel]) if is_tensor else c[sel]
  if return_index:
    return uniq, (torch.from_numpy(sel) if is_tensor else sel)
  return uniq


def batched_coordinates(coords, dtype=torch.int32, device=None):
  return torch.from_numpy(so.batched_coordinates([np.asarray(c) for c in coords])).to(dtype)


def _unsupported(name):
  class _Missing(nn.Module):
    def __init__(self, *a, **k):
      raise NotImplementedError(f'{name} is outside the registration path')
  _Missing.__name__ = name
  return _Missing


def module():
  me = types.ModuleType('MinkowskiEngine')
  me.__oracle_stand_in__ = True
  for obj in (RegionType, KernelGenerator, SparseTensor, MinkowskiNetwork, MinkowskiConvolution,
              MinkowskiConvolutionTranspose, MinkowskiBatchNorm):
    setattr(me, obj.__name__, obj)
  me.cat = cat
  for name in ('MinkowskiSumPooling', 'MinkowskiPoolingTranspose', 'MinkowskiInstanceNorm', 'MinkowskiReLU',
               'MinkowskiELU', 'MinkowskiGlobalPooling', 'MinkowskiBroadcastAddition'):
    setattr(me, name, _unsupported(name))
  mef = types.ModuleType('MinkowskiEngine.MinkowskiFunctional')
  mef.relu = _relu
  utils = types.ModuleType('MinkowskiEngine.utils')
  utils.kaiming_normal_ = lambda *a, **k: None
  utils.sparse_quantize = sparse_quantize
  utils.batched_coordinates = batched_coordinates
  me.MinkowskiFunctional, me.utils = mef, utils
  return me, mef, utils


def install():
  """Register the stand-in as ``MinkowskiEngine``; returns a function that restores sys.modules."""
  names = ('MinkowskiEngine', 'MinkowskiEngine.MinkowskiFunctional', 'MinkowskiEngine.utils')
  saved = {k: sys.modules.get(k) for k in names}
  for k, m in zip(names, module()):
    sys.modules[k] = m

  def restore():
    for k, m in saved.items():
      if m is None:
        sys.modules.pop(k, None)
      else:
        sys.modules[k] = m
  return restore
