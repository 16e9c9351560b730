"""The ResUNet2 family (ResUNetBN2C is the FCGF extractor at D=3 and the inlier network at
D=6) on top of the ME-shaped operator API.

Same constructor signature, attribute names (hence state-dict keys) and forward graph as
the reference's ``ResUNet2`` (model/resunet.py:419-665, blocks from
model/residual_block.py:83-134, norms from model/common.py:11-13), so a checkpoint written
by the reference loads with ``load_state_dict`` unchanged.  The layers are declared from a
table instead of the reference's spelled-out constructor, and ``forward`` has two
executions of the same graph:

  * ``forward``        the operator-by-operator path through the ME-shaped modules;
  * ``forward_fused``  the production path: eval-BatchNorm folded to scale/shift and applied
                       together with the residual add and ReLU in one pass, ME.cat fused
                       into the consuming 1x1 convolution, ReLU+bias+L2-normalise fused
                       into the 1x1 epilogues.
"""
import torch
import torch.nn as nn

from .. import _abi
from .. import me as ME
from ..me import MinkowskiFunctional as MEF


def conv(in_channels, out_channels, kernel_size=3, stride=1, dilation=1, has_bias=False, region_type=0,
         dimension=3):
  """Reference helper (model/residual_block.py:15-44): note it never forwards has_bias or
  dilation, so these convolutions are bias-free."""
  kg = ME.KernelGenerator(kernel_size=kernel_size, stride=stride, dilation=1, dimension=dimension)
  return ME.MinkowskiConvolution(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                 kernel_generator=kg, dimension=dimension)


def conv_tr(in_channels, out_channels, kernel_size, stride=1, dilation=1, has_bias=False,
            region_type=ME.RegionType.HYPER_CUBE, dimension=-1):
  assert dimension > 0, 'Dimension must be a positive integer'
  kg = ME.KernelGenerator(kernel_size, stride, dilation, is_transpose=True, dimension=dimension)
  return ME.MinkowskiConvolutionTranspose(in_channels=in_channels, out_channels=out_channels,
                                          kernel_size=kernel_size, stride=stride, dilation=dilation,
                                          bias=has_bias, kernel_generator=kg, dimension=dimension)


def get_norm(norm_type, num_feats, bn_momentum=0.05, dimension=-1):
  if norm_type == 'BN':
    return ME.MinkowskiBatchNorm(num_feats, momentum=bn_momentum)
  raise ValueError(f'Type {norm_type}, not defined (dgr_b200 builds BN only)')


class BasicBlockBN(nn.Module):
  """conv3-BN-ReLU-conv3-BN, add the input, ReLU (model/residual_block.py:83-134)."""

  def __init__(self, inplanes, planes, bn_momentum=0.1, D=3):
    super().__init__()
    self.conv1 = conv(inplanes, planes, kernel_size=3, stride=1, dimension=D)
    self.norm1 = get_norm('BN', planes, bn_momentum=bn_momentum, dimension=D)
    self.conv2 = conv(planes, planes, kernel_size=3, stride=1, dimension=D)
    self.norm2 = get_norm('BN', planes, bn_momentum=bn_momentum, dimension=D)

  def forward(self, x):
    out = MEF.relu(self.norm1(self.conv1(x)))
    out = self.norm2(self.conv2(out))
    out += x
    return MEF.relu(out)


class ResUNet2(ME.MinkowskiNetwork):
  NORM_TYPE = None
  BLOCK_NORM_TYPE = 'BN'
  CHANNELS = [None, 32, 64, 128, 256]
  TR_CHANNELS = [None, 32, 64, 64, 128]
  REGION_TYPE = ME.RegionType.HYPER_CUBE

  def __init__(self, in_channels=3, out_channels=32, bn_momentum=0.1, conv1_kernel_size=3,
               normalize_feature=False, D=3):
    super().__init__(D)
    C, T = self.CHANNELS, self.TR_CHANNELS
    if self.NORM_TYPE != 'BN':
      raise NotImplementedError('only the BN variants of ResUNet2 are built (released DGR weights use '
                                'ResUNetBN2C)')
    self.normalize_feature = normalize_feature
    self.conv1_kernel_size = conv1_kernel_size
    # encoder: conv{l} (stride 2 from level 2 on) -> norm{l} -> block{l}
    enc_in = [None, in_channels, C[1], C[2], C[3]]
    for l in (1, 2, 3, 4):
      k, s = (conv1_kernel_size, 1) if l == 1 else (3, 2)
      setattr(self, f'conv{l}', conv(enc_in[l], C[l], kernel_size=k, stride=s, dimension=D))
      setattr(self, f'norm{l}', get_norm(self.NORM_TYPE, C[l], bn_momentum=bn_momentum, dimension=D))
      setattr(self, f'block{l}', BasicBlockBN(C[l], C[l], bn_momentum=bn_momentum, D=D))
    # decoder: conv{l}_tr (stride 2) -> norm{l}_tr -> block{l}_tr, then concat with the skip
    dec_in = {4: C[4], 3: C[3] + T[4], 2: C[2] + T[3]}
    for l in (4, 3, 2):
      setattr(self, f'conv{l}_tr', conv_tr(dec_in[l], T[l], kernel_size=3, stride=2, dimension=D))
      setattr(self, f'norm{l}_tr', get_norm(self.NORM_TYPE, T[l], bn_momentum=bn_momentum, dimension=D))
      setattr(self, f'block{l}_tr', BasicBlockBN(T[l], T[l], bn_momentum=bn_momentum, D=D))
    self.conv1_tr = conv(C[1] + T[2], T[1], kernel_size=1, stride=1, dimension=D)
    self.final = ME.MinkowskiConvolution(T[1], out_channels, kernel_size=1, stride=1, dilation=1,
                                         bias=True, dimension=D)

  # ---------------------------------------------------------------------------------------
  def forward(self, x):
    """Operator-by-operator execution, one ME-shaped call per reference line
    (model/resunet.py:598-649)."""
    skips = {}
    out = x
    for l in (1, 2, 3, 4):
      out = getattr(self, f'conv{l}')(out)
      out = getattr(self, f'norm{l}')(out)
      out = getattr(self, f'block{l}')(out)
      skips[l] = out
      out = MEF.relu(out)
    for l in (4, 3, 2):
      out = getattr(self, f'conv{l}_tr')(out)
      out = getattr(self, f'norm{l}_tr')(out)
      out = getattr(self, f'block{l}_tr')(out)
      out = ME.cat(MEF.relu(out), skips[l - 1])
    out = MEF.relu(self.conv1_tr(out))
    out = self.final(out)
    if self.normalize_feature:
      return ME.SparseTensor(_abi.l2_normalize(out.F), coordinate_map_key=out.coordinate_map_key,
                             coordinate_manager=out.coordinate_manager)
    return out

  # ---------------------------------------------------------------------------------------
  def _conv_bn(self, feat, conv_mod, norm_mod, km, out, residual=None, relu=False):
    """sparse conv -> (BN scale/shift [+ residual] [+ ReLU]) in one elementwise pass.
    `out` is a pre-zeroed [n_out, cout] buffer (or None for the output-stationary conv1)."""
    scale, shift = norm_mod.folded()
    if out is None:
      return _abi.spconv_table_fwd(feat, conv_mod.kernel.detach(), km, conv_mod.out_channels, scale, shift)
    ME.sparse_conv(feat, conv_mod, km, out)
    return _abi.affine_act(out, scale=scale, shift=shift, residual=residual, relu=relu, out=out)

  def _uses_table(self, conv_mod, km):
    return km.nbr is not None and conv_mod.in_channels <= 8 and conv_mod.out_channels in (16, 32, 64)

  def _plan(self, man, key):
    """All coordinate maps and kernel maps of the network, built up front: this is where every
    host synchronisation of the forward pass happens (one bucket-offset read per map); the
    convolution phase that follows is launch-only.  Returns the layers in execution order as
    (conv module, norm module, kernel map, role) with the total size of their outputs."""
    s0 = key.stride
    man.prepare([2 * s0, 4 * s0, 8 * s0],
                [(s0, 1, self.conv1.kernel_size), (s0, 1, 3), (s0, 2, 3), (2 * s0, 1, 3), (2 * s0, 2, 3),
                 (4 * s0, 1, 3), (4 * s0, 2, 3), (8 * s0, 1, 3)])
    layers = []
    for l in (1, 2, 3, 4):
      cm = getattr(self, f'conv{l}')
      key_out, km = man.kernel_map(key, cm.stride, cm.kernel_size)
      layers.append((cm, getattr(self, f'norm{l}'), km, 'conv'))
      _, km3 = man.kernel_map(key_out, 1, 3)
      blk = getattr(self, f'block{l}')
      layers += [(blk.conv1, blk.norm1, km3, 'b1'), (blk.conv2, blk.norm2, km3, 'b2')]
      key = key_out
    for l in (4, 3, 2):
      cm = getattr(self, f'conv{l}_tr')
      key_out, km = man.transpose_kernel_map(key, cm.stride, cm.kernel_size)
      layers.append((cm, getattr(self, f'norm{l}_tr'), km, 'conv'))
      _, km3 = man.kernel_map(key_out, 1, 3)
      blk = getattr(self, f'block{l}_tr')
      layers += [(blk.conv1, blk.norm1, km3, 'b1'), (blk.conv2, blk.norm2, km3, 'b2')]
      key = key_out
    total = sum(km.n_out * cm.out_channels for cm, _, km, _ in layers if not self._uses_table(cm, km))
    return layers, total, key

  def forward_fused(self, x):
    """Same graph, fused epilogues.  The ReLUs after each block are idempotent (the block
    already ends in ReLU) and are dropped.  Phase 1 builds every kernel map (all host syncs);
    phase 2 zero-fills ONE slab holding every convolution output and launches the layers
    back to back."""
    man = x.coordinate_manager
    layers, total, key_final = self._plan(man, x.coordinate_map_key)
    slab = _abi.scratch(('conv_out', id(self)), max(total, 1), torch.float32, x.device)
    slab.zero_()
    ofs = 0

    def take(cm, km):
      nonlocal ofs
      if self._uses_table(cm, km):
        return None
      n = km.n_out * cm.out_channels
      buf = slab[ofs:ofs + n].view(km.n_out, cm.out_channels)
      ofs += n
      return buf

    feat, skips, block_in, level = x.F, [], None, 0
    it = iter(layers)
    for stage in range(7):          # 4 encoder levels, then 3 decoder levels
      cm, nm, km, _ = next(it)
      feat = self._conv_bn(feat, cm, nm, km, take(cm, km))
      c1, n1, km3, _ = next(it)
      c2, n2, _, _ = next(it)
      h = self._conv_bn(feat, c1, n1, km3, take(c1, km3), relu=True)
      feat = self._conv_bn(h, c2, n2, km3, take(c2, km3), residual=feat, relu=True)
      if stage < 4:
        skips.append(feat)            # out_s1, out_s2, out_s4, out_s8
      elif stage < 6:
        # decoder levels 4 and 3 feed a 3^D transposed conv: materialise ME.cat(decoder, skip)
        feat = _abi.cat2(feat, skips[2 - (stage - 4)])
    # conv1_tr reads (decoder, skip) directly: ME.cat fused into the 1x1 convolution
    h = _abi.linear_fwd(feat, self.conv1_tr.kernel.detach(), None, b=skips[0], relu=True)
    out = _abi.linear_fwd(h, self.final.kernel.detach(), self.final.bias.detach().reshape(-1).contiguous(),
                          normalize=bool(self.normalize_feature))
    return ME.SparseTensor(out, coordinate_map_key=key_final, coordinate_manager=man)


class ResUNetBN2(ResUNet2):
  NORM_TYPE = 'BN'


class ResUNetBN2B(ResUNet2):
  NORM_TYPE = 'BN'
  TR_CHANNELS = [None, 64, 64, 64, 64]


class ResUNetBN2C(ResUNet2):
  NORM_TYPE = 'BN'
  TR_CHANNELS = [None, 64, 64, 64, 128]


class ResUNetBN2D(ResUNet2):
  NORM_TYPE = 'BN'
  TR_CHANNELS = [None, 64, 64, 128, 128]


class ResUNetBN2E(ResUNet2):
  NORM_TYPE = 'BN'
  CHANNELS = [None, 128, 128, 128, 256]
  TR_CHANNELS = [None, 64, 128, 128, 128]


class ResUNetBN2F(ResUNet2):
  NORM_TYPE = 'BN'
  CHANNELS = [None, 16, 32, 64, 128]
  TR_CHANNELS = [None, 16, 32, 64, 128]
